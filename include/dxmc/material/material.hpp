// dxmc/material/material.hpp — Material<NSHELLS>: byWeight / byNistName / byChemicalFormula -> optional,
// attenuationValues(E).sum() (R:src/libopendxmc/simulationpipeline.cpp:136; ctsegmentationpipeline.cpp:73-104),
// parseCompoundStr (R:src/libopendxmc/hdf5wrapper.cpp:1100).
#pragma once
#include "../../dxb.h"
#include "atomhandler.hpp" // the reference uses dxmc::AtomHandler after including only material.hpp (R:src/libopendxmc/hdf5wrapper.cpp:429)
#include "nistmaterials.hpp"
#include <cctype>
#include <map>
#include <memory>
#include <optional>
#include <string>
#include <vector>
namespace dxmc {
struct AttenuationValues {
    double photoelectric = 0, incoherent = 0, coherent = 0;
    double sum() const { return photoelectric + incoherent + coherent; }
};

template <int NSHELLS = 5>
class Material {
public:
    static std::optional<Material<NSHELLS>> byWeight(const std::map<std::size_t, double>& weights)
    {
        std::vector<uint32_t> Z;
        std::vector<double> w;
        for (const auto& [z, v] : weights) {
            Z.push_back(static_cast<uint32_t>(z));
            w.push_back(v);
        }
        dxb_material* h = nullptr;
        if (dxb_material_by_weight(&h, static_cast<uint32_t>(Z.size()), Z.data(), w.data()) != DXB_OK)
            return std::nullopt;
        return Material(h);
    }
    static std::optional<Material<NSHELLS>> byNistName(const std::string& name)
    {
        dxb_material* h = nullptr;
        if (dxb_material_by_nist_name(&h, name.c_str()) != DXB_OK)
            return std::nullopt;
        return Material(h);
    }
    static std::optional<Material<NSHELLS>> byChemicalFormula(const std::string& formula)
    {
        dxb_material* h = nullptr;
        if (dxb_material_by_chemical_formula(&h, formula.c_str()) != DXB_OK)
            return std::nullopt;
        return Material(h);
    }
    // "H2O", "C5O2H8", "Ca10(PO4)6(OH)2" are not needed by OpenDXMC: it stores "Z:weight"-free element-count strings;
    // element symbol + optional count, returned as atom counts per Z
    static std::map<std::size_t, double> parseCompoundStr(const std::string& str)
    {
        std::map<std::size_t, double> res;
        std::size_t i = 0;
        while (i < str.size()) {
            if (!std::isupper(static_cast<unsigned char>(str[i]))) {
                ++i;
                continue;
            }
            std::string sym(1, str[i++]);
            while (i < str.size() && std::islower(static_cast<unsigned char>(str[i])))
                sym.push_back(str[i++]);
            std::string num;
            while (i < str.size() && (std::isdigit(static_cast<unsigned char>(str[i])) || str[i] == '.'))
                num.push_back(str[i++]);
            std::size_t Z = 0;
            for (uint32_t z = 1; z <= 92; ++z)
                if (sym == dxb_atom_symbol(z)) {
                    Z = z;
                    break;
                }
            if (Z)
                res[Z] += num.empty() ? 1.0 : std::stod(num);
        }
        return res;
    }
    AttenuationValues attenuationValues(double energy) const
    {
        double v[3] = { 0, 0, 0 };
        dxb_material_attenuation(m_h.get(), energy, v);
        return { v[0], v[1], v[2] };
    }
    double massEnergyTransferAttenuation(double energy) const { return dxb_material_mass_energy_transfer(m_h.get(), energy); }
    double effectiveZ() const { return dxb_material_effective_z(m_h.get()); }
    double formFactor(double x) const { return dxb_material_form_factor(m_h.get(), x); }
    double scatterFactor(double x) const { return dxb_material_scatter_factor(m_h.get(), x); }
    const dxb_material* handle() const { return m_h.get(); }

private:
    explicit Material(dxb_material* h)
        : m_h(h, [](dxb_material* p) { dxb_material_destroy(p); })
    {
    }
    std::shared_ptr<dxb_material> m_h;
};
}
