// dxmc/material/atomhandler.hpp — AtomHandler::toSymbol (R:src/libopendxmc/hdf5wrapper.cpp:429).
#pragma once
#include "../../dxb.h"
#include <string>
namespace dxmc {
struct AtomHandler {
    static std::string toSymbol(std::size_t Z) { return dxb_atom_symbol(static_cast<uint32_t>(Z)); }
    static double atomicWeight(std::size_t Z) { return dxb_atom_weight(static_cast<uint32_t>(Z)); }
    static double standardDensity(std::size_t Z) { return dxb_atom_standard_density(static_cast<uint32_t>(Z)); }
};
}
