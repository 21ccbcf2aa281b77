// tube.cpp — dxmc::Tube replacement (host side, once per beam).
// Call sites: R:src/libopendxmc/ctsegmentationpipeline.cpp:66-71 (Tube(kVp), setAlFiltration,
// getEnergy, getSpecter) and R:src/libopendxmc/beamsettingsmodel.cpp:257-346 (voltage, anode
// angle, filtration by Z, half value layer, mean energy).
//
// DXMClib uses a semi-analytical tungsten-anode model whose source and constants are not in
// this container.  This is an independent thick-target model:
//   bremsstrahlung  Kramers thin-target yield 1/(T E) integrated along the Thomson-Whiddington
//                   slowing-down path T^2 = T0^2 - rho C x, each depth attenuated by the
//                   tungsten it has to cross towards the take-off direction (anode heel),
//   characteristic  tungsten K lines with yield ~ (U-1)^1.63 above 69.5 kV,
//   filtration      exp(-mu rho t) for each added element, using this library's cross sections.
#include "physics.hpp"

#include <algorithm>
#include <cmath>

namespace dxb {

namespace {

double elementTotal(const Element* el, double e) // cm2/g
{
    const GridPos p = energyPos(std::clamp(e, kEMin, kEMax));
    return (lerpTable(el->photo, p) + lerpTable(el->incoh, p) + lerpTable(el->coh, p)) * kAvogadro / el->A;
}

dxb_tube_desc sanitized(const dxb_tube_desc& in)
{
    dxb_tube_desc t = in;
    t.voltage_kv = std::clamp(t.voltage_kv > 0 ? t.voltage_kv : 120.0, 20.0, 150.0);
    if (!(t.anode_angle_deg > 0))
        t.anode_angle_deg = 12.0;
    t.anode_angle_deg = std::clamp(t.anode_angle_deg, 1.0, 89.0);
    if (!(t.energy_resolution_kev > 0))
        t.energy_resolution_kev = 1.0;
    t.n_filt = std::min<uint32_t>(t.n_filt, DXB_TUBE_MAX_FILT);
    return t;
}

} // namespace

std::vector<double> tubeEnergies(const dxb_tube_desc& tin)
{
    const dxb_tube_desc t = sanitized(tin);
    std::vector<double> e;
    const double step = t.energy_resolution_kev;
    for (double v = step; v <= t.voltage_kv + 1e-9; v += step)
        e.push_back(v);
    return e;
}

std::vector<double> tubeSpectrum(const dxb_tube_desc& tin, const std::vector<double>& energies, bool normalize)
{
    const dxb_tube_desc t = sanitized(tin);
    const double T0 = t.voltage_kv;
    const Element* W = getElement(74);
    const double tanA = std::tan(t.anode_angle_deg * kPi / 180.0);
    constexpr double kTW = 0.65e6; // Thomson-Whiddington constant for W [keV^2 cm^2/g] around 100 kV
    std::vector<double> w(energies.size(), 0.0);

    const double step = energies.size() > 1 ? energies[1] - energies[0] : 1.0;
    double bremsTotal = 0;
    for (size_t i = 0; i < energies.size(); ++i) {
        const double E = energies[i];
        if (E >= T0 || E < kEMin)
            continue;
        const double muW = elementTotal(W, E); // cm2/g
        // N(E) = (2/(rho C E)) * int_E^T0 exp(-mu (T0^2 - T^2)/(C tan a)) dT      (rho cancels)
        constexpr int NT = 64;
        double sum = 0;
        for (int j = 0; j <= NT; ++j) {
            const double T = E + (T0 - E) * j / NT;
            const double depth = (T0 * T0 - T * T) / kTW; // g/cm2 along the electron direction
            const double q = (j == 0 || j == NT) ? 1.0 : ((j & 1) ? 4.0 : 2.0);
            sum += q * std::exp(-muW * depth / tanA);
        }
        w[i] = sum * (T0 - E) / NT / 3.0 / E;
        bremsTotal += w[i] * step;
    }
    // characteristic K lines of tungsten
    const double edgeK = W->edgeK;
    if (T0 > edgeK && bremsTotal > 0) {
        static const double lineE[4] = { 59.318, 57.982, 67.244, 69.067 };
        static const double lineF[4] = { 0.501, 0.291, 0.165, 0.043 };
        // un-filtered brems photons above 10 keV as the yard-stick
        double ref = 0;
        for (size_t i = 0; i < energies.size(); ++i)
            if (energies[i] >= 10.0)
                ref += w[i] * step;
        const double U = T0 / edgeK;
        const double kTotal = 0.105 * std::pow(U - 1.0, 1.63) / std::pow(120.0 / edgeK - 1.0, 1.63) * ref;
        for (int l = 0; l < 4; ++l) {
            // nearest bin
            size_t best = 0;
            double bd = 1e30;
            for (size_t i = 0; i < energies.size(); ++i) {
                const double d = std::fabs(energies[i] - lineE[l]);
                if (d < bd) {
                    bd = d;
                    best = i;
                }
            }
            if (energies[best] < T0) {
                // lines are born at roughly the mean electron depth; attenuate like a T = 0.85 T0 photon source
                const double depth = (T0 * T0 - 0.7225 * T0 * T0) / kTW;
                w[best] += kTotal * lineF[l] * std::exp(-elementTotal(W, lineE[l]) * depth / tanA) / step;
            }
        }
    }
    // added filtration
    for (uint32_t f = 0; f < t.n_filt; ++f) {
        const Element* el = getElement(t.filt_Z[f]);
        if (!el || !(t.filt_mm[f] > 0))
            continue;
        const double massThickness = el->density * t.filt_mm[f] * 0.1; // g/cm2
        for (size_t i = 0; i < energies.size(); ++i)
            if (w[i] > 0)
                w[i] *= std::exp(-elementTotal(el, energies[i]) * massThickness);
    }
    if (normalize) {
        double s = 0;
        for (double v : w)
            s += v;
        if (s > 0)
            for (double& v : w)
                v /= s;
    }
    return w;
}

double tubeMeanEnergy(const dxb_tube_desc& t)
{
    const auto e = tubeEnergies(t);
    const auto w = tubeSpectrum(t, e, true);
    double m = 0;
    for (size_t i = 0; i < e.size(); ++i)
        m += e[i] * w[i];
    return m;
}

double tubeAlHVLmm(const dxb_tube_desc& t)
{
    const auto e = tubeEnergies(t);
    const auto w = tubeSpectrum(t, e, true);
    auto air = Material::byNistName("Air, Dry (near sea level)");
    const Element* al = getElement(13);
    std::vector<double> kerma(e.size()), muAl(e.size());
    double k0 = 0;
    for (size_t i = 0; i < e.size(); ++i) {
        const double en = std::clamp(e[i], kEMin, kEMax);
        kerma[i] = w[i] * e[i] * air->massEnergyTransfer(en);
        muAl[i] = elementTotal(al, en) * al->density; // 1/cm
        k0 += kerma[i];
    }
    if (!(k0 > 0))
        return 0;
    auto transmitted = [&](double cm) {
        double k = 0;
        for (size_t i = 0; i < e.size(); ++i)
            k += kerma[i] * std::exp(-muAl[i] * cm);
        return k / k0;
    };
    double lo = 0, hi = 10.0;
    for (int it = 0; it < 60; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (transmitted(mid) > 0.5)
            lo = mid;
        else
            hi = mid;
    }
    return 0.5 * (lo + hi) * 10.0;
}

} // namespace dxb
