// dxmc/material/nistmaterials.hpp — NISTMaterials::density / ::Composition
// (R:src/libopendxmc/ctsegmentationpipeline.cpp:74-76,163; R:src/libopendxmc/otherphantomimportpipeline.cpp:87-93).
#pragma once
#include "../../dxb.h"
#include <map>
#include <string>
#include <vector>
namespace dxmc {
class NISTMaterials {
public:
    static std::vector<std::string> listNames()
    {
        std::vector<std::string> n;
        for (int i = 0; i < dxb_nist_count(); ++i)
            n.emplace_back(dxb_nist_name(i));
        return n;
    }
    static double density(const std::string& name)
    {
        const double d = dxb_nist_density(name.c_str());
        return d > 0 ? d : 0.0;
    }
    static std::map<std::size_t, double> Composition(const std::string& name)
    {
        uint32_t Z[32];
        double w[32];
        const int n = dxb_nist_composition(name.c_str(), Z, w, 32);
        std::map<std::size_t, double> c;
        for (int i = 0; i < n && i < 32; ++i)
            c[Z[i]] = w[i];
        return c;
    }
};
}
