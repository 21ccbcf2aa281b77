// dxmc/beams/pencilbeam.hpp — dxmc::PencilBeam<ENABLETRACKING> (R:src/libopendxmc/beamsettingsmodel.cpp:637-711).
#pragma once
#include "beamtype.hpp"
#include <algorithm>
namespace dxmc {
template <bool ENABLETRACKING = false>
class PencilBeam : public detail::BeamBase {
public:
    PencilBeam(const std::array<double, 3>& pos = { 0, 0, 0 }, const std::array<double, 3>& dir = { 0, 0, 1 }, double energy = 60)
        : detail::BeamBase(DXB_BEAM_PENCIL)
    {
        m_nTubes = 0;
        setPosition(pos);
        setDirection(dir);
        setEnergy(energy);
    }
    std::array<double, 3> position() const { return get3(m_d.position); }
    void setPosition(const std::array<double, 3>& p) { set3(m_d.position, p); }
    std::array<double, 3> direction() const { return get3(m_d.direction); }
    void setDirection(const std::array<double, 3>& d) { set3(m_d.direction, vectormath::normalized(d)); }
    double energy() const { return m_d.energy; }
    void setEnergy(double e) { m_d.energy = std::clamp(e, MIN_ENERGY(), MAX_ENERGY()); }
    double airKerma() const { return m_d.air_kerma; }
    void setAirKerma(double k) { m_d.air_kerma = std::abs(k); }
    void setNumberOfExposures(std::uint64_t n) { m_d.n_exposures = n > 0 ? n : 1; }
};
}
