// icrp.hpp — ICRP phantom import plan (icrp.cpp) shared with the device remap (import_kernels.cu)
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace dxb {

struct IcrpPlan {
    uint8_t organLut[256];    // organ value in the file -> organ index after arm removal / pruning
    uint8_t materialLut[256]; // organ value in the file -> medium index
    double densityLut[256];   // organ value in the file -> density [g/cm3]
    std::vector<std::string> organNames;
    std::vector<double> organDensity;
    std::vector<uint8_t> organMedium;
    std::vector<std::string> mediaNames;
    std::vector<std::vector<std::pair<uint32_t, double>>> mediaComposition; // (Z, mass %), zero entries kept like the reference
};

bool icrpPlan(const char* organsText, const char* mediaText, bool removeArms, const uint8_t present[256], IcrpPlan& plan);

} // namespace dxb

struct dxb_icrp {
    dxb::IcrpPlan plan;
};
