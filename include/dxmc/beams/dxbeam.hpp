// dxmc/beams/dxbeam.hpp — dxmc::DXBeam<ENABLETRACKING>: ctor (pos, cosines[2], filtration) R:src/libopendxmc/dxmc_specialization.cpp:22;
// accessors R:src/libopendxmc/beamsettingsmodel.cpp:470-611, R:src/libopendxmc/beamactorcontainer.cpp:166-168.
#pragma once
#include "beamtype.hpp"
namespace dxmc {
template <bool ENABLETRACKING = false>
class DXBeam : public detail::BeamBase {
public:
    DXBeam(const std::array<double, 3>& pos = { 0, 0, 0 },
        const std::array<std::array<double, 3>, 2>& dircosines = { { { 1, 0, 0 }, { 0, 1, 0 } } },
        const std::map<std::size_t, double>& filtrationMaterials = {})
        : detail::BeamBase(DXB_BEAM_DX)
    {
        setPosition(pos);
        setDirectionCosines(dircosines);
        m_tube[0].setFiltrationMaterials(filtrationMaterials);
    }
    const std::array<double, 3> position() const { return get3(m_d.position); }
    void setPosition(const std::array<double, 3>& p) { set3(m_d.position, p); }
    std::array<std::array<double, 3>, 2> directionCosines() const { return { { get3(m_d.cosines[0]), get3(m_d.cosines[1]) } }; }
    void setDirectionCosines(const std::array<std::array<double, 3>, 2>& c)
    {
        set3(m_d.cosines[0], vectormath::normalized(c[0]));
        set3(m_d.cosines[1], vectormath::normalized(c[1]));
    }
    void setDirectionCosines(const std::array<double, 3>& x, const std::array<double, 3>& y) { setDirectionCosines({ { x, y } }); }
    const std::array<double, 2> collimationHalfAngles() const { return { m_d.half_angles[0], m_d.half_angles[1] }; }
    void setCollimationHalfAngles(const std::array<double, 2>& a)
    {
        m_d.half_angles[0] = std::abs(a[0]);
        m_d.half_angles[1] = std::abs(a[1]);
    }
    void setCollimationHalfAngles(double x, double y) { setCollimationHalfAngles({ x, y }); }
    std::array<double, 2> collimationHalfAnglesDeg() const { return { m_d.half_angles[0] * RAD_TO_DEG(), m_d.half_angles[1] * RAD_TO_DEG() }; }
    void setCollimationHalfAnglesDeg(const std::array<double, 2>& a) { setCollimationHalfAngles({ a[0] * DEG_TO_RAD(), a[1] * DEG_TO_RAD() }); }
    void setCollimationHalfAnglesDeg(double x, double y) { setCollimationHalfAnglesDeg({ x, y }); }
    double DAPvalue() const { return m_d.dap; }
    void setDAPvalue(double v) { m_d.dap = std::abs(v); }
    void setNumberOfExposures(std::uint64_t n) { m_d.n_exposures = n > 0 ? n : 1; }
    // tube group (R:src/libopendxmc/beamsettingsmodel.cpp:257-346)
    const Tube& tube() const { return m_tube[0]; }
    void setTube(const Tube& t) { m_tube[0] = t; }
    void setTubeVoltage(double kv) { m_tube[0].setVoltage(kv); }
    void setTubeAnodeAngle(double rad) { m_tube[0].setAnodeAngle(rad); }
    void setTubeAnodeAngleDeg(double deg) { m_tube[0].setAnodeAngleDeg(deg); }
    void addTubeFiltrationMaterial(std::size_t Z, double mm) { m_tube[0].addFiltrationMaterial(Z, mm); }
    void removeTubeFiltrationMaterial(std::size_t Z) { m_tube[0].addFiltrationMaterial(Z, 0.0); }
    double tubeFiltration(std::size_t Z) const { return m_tube[0].filtration(Z); }
    void clearTubeFiltrationMaterials() { m_tube[0].clearFiltrationMaterials(); }
    double tubeAlHalfValueLayer() const { return m_tube[0].mmAlHalfValueLayer(); }
    double tubeMeanSpecterEnergy() const { return m_tube[0].meanSpecterEnergy(); }
    void setSourceDetectorDistanceForCalibration(double sdd) { m_d.sdd = std::abs(sdd); }
};
}
