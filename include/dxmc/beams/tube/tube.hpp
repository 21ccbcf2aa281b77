// dxmc/beams/tube/tube.hpp — dxmc::Tube (R:src/libopendxmc/ctsegmentationpipeline.cpp:66-71,
// R:src/libopendxmc/beamsettingsmodel.cpp:257-346).
#pragma once
#include "../../../dxb.h"
#include <algorithm>
#include <cmath>
#include <map>
#include <vector>
namespace dxmc {
class Tube {
public:
    explicit Tube(double tubeVoltage = 120.0, double anodeAngleDeg = 12.0, double energyResolution = 1.0)
        : m_voltage(std::clamp(tubeVoltage, minVoltage(), maxVoltage()))
        , m_anodeAngleDeg(anodeAngleDeg)
        , m_resolution(energyResolution)
    {
    }
    static constexpr double maxVoltage() { return 150.0; }
    static constexpr double minVoltage() { return 20.0; }
    double voltage() const { return m_voltage; }
    void setVoltage(double v) { m_voltage = std::clamp(v, minVoltage(), maxVoltage()); }
    double anodeAngle() const { return m_anodeAngleDeg * 3.14159265358979323846 / 180.0; }
    double anodeAngleDeg() const { return m_anodeAngleDeg; }
    void setAnodeAngle(double rad) { setAnodeAngleDeg(rad * 180.0 / 3.14159265358979323846); }
    void setAnodeAngleDeg(double deg) { m_anodeAngleDeg = std::clamp(std::abs(deg), 1.0, 89.0); }
    void addFiltrationMaterial(std::size_t Z, double mm)
    {
        if (Z >= 1 && Z <= 92 && (m_filtration.count(Z) || m_filtration.size() < DXB_TUBE_MAX_FILT))
            m_filtration[Z] = std::abs(mm);
    }
    double filtration(std::size_t Z) const
    {
        const auto it = m_filtration.find(Z);
        return it == m_filtration.end() ? 0.0 : it->second;
    }
    const std::map<std::size_t, double>& filtrationMaterials() const { return m_filtration; }
    void setFiltrationMaterials(const std::map<std::size_t, double>& f)
    {
        m_filtration.clear();
        for (const auto& [z, mm] : f)
            addFiltrationMaterial(z, mm);
    }
    void clearFiltrationMaterials() { m_filtration.clear(); }
    void setAlFiltration(double mm) { addFiltrationMaterial(13, mm); }
    void setCuFiltration(double mm) { addFiltrationMaterial(29, mm); }
    void setSnFiltration(double mm) { addFiltrationMaterial(50, mm); }
    double AlFiltration() const { return filtration(13); }
    double CuFiltration() const { return filtration(29); }
    double SnFiltration() const { return filtration(50); }
    double energyResolution() const { return m_resolution; }
    void setEnergyResolution(double r) { m_resolution = std::clamp(r, 0.1, 10.0); }

    std::vector<double> getEnergy() const
    {
        const dxb_tube_desc d = desc();
        const int n = dxb_tube_energies(&d, nullptr, 0);
        std::vector<double> e(static_cast<std::size_t>(n));
        dxb_tube_energies(&d, e.data(), n);
        return e;
    }
    std::vector<double> getSpecter(const std::vector<double>& energies, bool normalize = true) const
    {
        const dxb_tube_desc d = desc();
        std::vector<double> w(energies.size(), 0.0);
        if (!energies.empty())
            dxb_tube_spectrum(&d, energies.data(), static_cast<int>(energies.size()), normalize ? 1 : 0, w.data());
        return w;
    }
    std::vector<std::pair<double, double>> getSpecter(bool normalize = true) const
    {
        const auto e = getEnergy();
        const auto w = getSpecter(e, normalize);
        std::vector<std::pair<double, double>> r(e.size());
        for (std::size_t i = 0; i < e.size(); ++i)
            r[i] = { e[i], w[i] };
        return r;
    }
    double mmAlHalfValueLayer() const
    {
        const dxb_tube_desc d = desc();
        return dxb_tube_al_half_value_layer_mm(&d);
    }
    double meanSpecterEnergy() const
    {
        const dxb_tube_desc d = desc();
        return dxb_tube_mean_energy(&d);
    }
    dxb_tube_desc desc() const
    {
        dxb_tube_desc d {};
        d.voltage_kv = m_voltage;
        d.anode_angle_deg = m_anodeAngleDeg;
        d.energy_resolution_kev = m_resolution;
        d.n_filt = 0;
        for (const auto& [z, mm] : m_filtration) {
            if (d.n_filt >= DXB_TUBE_MAX_FILT)
                break;
            d.filt_Z[d.n_filt] = static_cast<uint32_t>(z);
            d.filt_mm[d.n_filt] = mm;
            ++d.n_filt;
        }
        return d;
    }

private:
    double m_voltage, m_anodeAngleDeg, m_resolution;
    std::map<std::size_t, double> m_filtration;
};
}
