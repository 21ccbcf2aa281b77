// dxmc/beams/cbctbeam.hpp — dxmc::CBCTBeam<ENABLETRACKING>: ctor (isocenter, axis, filtration)
// R:src/libopendxmc/beamsettingsmodel.cpp:730-895; radian accessors R:src/libopendxmc/hdf5wrapper.cpp:502-504,709-715.
#pragma once
#include "beamtype.hpp"
#include <algorithm>
namespace dxmc {
template <bool ENABLETRACKING = false>
class CBCTBeam : public detail::BeamBase {
public:
    CBCTBeam(const std::array<double, 3>& isocenter = { 0, 0, 0 }, const std::array<double, 3>& axis = { 0, 0, 1 },
        const std::map<std::size_t, double>& filtrationMaterials = {})
        : detail::BeamBase(DXB_BEAM_CBCT)
    {
        setIsocenter(isocenter);
        setRotationAxis(axis);
        m_tube[0].setFiltrationMaterials(filtrationMaterials);
    }
    std::array<double, 3> isocenter() const { return get3(m_d.isocenter); }
    void setIsocenter(const std::array<double, 3>& p) { set3(m_d.isocenter, p); }
    std::array<double, 3> rotationAxis() const { return get3(m_d.direction); }
    void setRotationAxis(const std::array<double, 3>& a) { set3(m_d.direction, vectormath::normalized(a)); }
    double sourceDetectorDistance() const { return m_d.sdd; }
    void setSourceDetectorDistance(double d) { m_d.sdd = std::max(std::abs(d), 1.0); }
    double startAngle() const { return m_d.start_angle; }
    void setStartAngle(double a) { m_d.start_angle = a; }
    double startAngleDeg() const { return m_d.start_angle * RAD_TO_DEG(); }
    void setStartAngleDeg(double a) { m_d.start_angle = a * DEG_TO_RAD(); }
    double stopAngle() const { return m_d.stop_angle; }
    void setStopAngle(double a) { m_d.stop_angle = a; }
    double stopAngleDeg() const { return m_d.stop_angle * RAD_TO_DEG(); }
    void setStopAngleDeg(double a) { m_d.stop_angle = a * DEG_TO_RAD(); }
    double stepAngle() const { return m_d.step_angle; }
    void setStepAngle(double a) { m_d.step_angle = std::max(std::abs(a), 0.1 * DEG_TO_RAD()); }
    double stepAngleDeg() const { return m_d.step_angle * RAD_TO_DEG(); }
    void setStepAngleDeg(double a) { setStepAngle(a * DEG_TO_RAD()); }
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
    const Tube& tube() const { return m_tube[0]; }
    void setTube(const Tube& t) { m_tube[0] = t; }
    void setTubeVoltage(double kv) { m_tube[0].setVoltage(kv); }
    void setTubeAnodeAngle(double rad) { m_tube[0].setAnodeAngle(rad); }
    void setTubeAnodeAngleDeg(double deg) { m_tube[0].setAnodeAngleDeg(deg); }
    void addTubeFiltrationMaterial(std::size_t Z, double mm) { m_tube[0].addFiltrationMaterial(Z, mm); }
    double tubeFiltration(std::size_t Z) const { return m_tube[0].filtration(Z); }
    void clearTubeFiltrationMaterials() { m_tube[0].clearFiltrationMaterials(); }
    double tubeAlHalfValueLayer() const { return m_tube[0].mmAlHalfValueLayer(); }
    double tubeMeanSpecterEnergy() const { return m_tube[0].meanSpecterEnergy(); }
};
}
