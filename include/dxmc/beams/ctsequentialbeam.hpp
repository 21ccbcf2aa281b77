// dxmc/beams/ctsequentialbeam.hpp — dxmc::CTSequentialBeam<ENABLETRACKING>, the reference's axial CT beam:
// ctor (start, normal, filtration) and accessors R:src/libopendxmc/beamsettingsmodel.cpp:921-1131.
#pragma once
#include "beamtype.hpp"
#include <algorithm>
namespace dxmc {
namespace detail {
    // what the sequential, spiral and dual-source CT beams share
    class CTBeamBase : public BeamBase {
    public:
        double sourceDetectorDistance() const { return m_d.sdd; }
        void setSourceDetectorDistance(double d) { m_d.sdd = std::max(std::abs(d), 1.0); }
        double collimation() const { return m_d.collimation; }
        void setCollimation(double c) { m_d.collimation = std::max(std::abs(c), 0.01); }
        double startAngle() const { return m_d.start_angle; }
        void setStartAngle(double a) { m_d.start_angle = a; }
        double startAngleDeg() const { return m_d.start_angle * RAD_TO_DEG(); }
        void setStartAngleDeg(double a) { m_d.start_angle = a * DEG_TO_RAD(); }
        double stepAngle() const { return m_d.step_angle; }
        void setStepAngle(double a) { m_d.step_angle = std::max(std::abs(a), 0.1 * DEG_TO_RAD()); }
        double stepAngleDeg() const { return m_d.step_angle * RAD_TO_DEG(); }
        void setStepAngleDeg(double a) { setStepAngle(a * DEG_TO_RAD()); }
        double CTDIdiameter() const { return m_d.ctdi_diameter; }
        void setCTDIdiameter(double d) { m_d.ctdi_diameter = std::max(std::abs(d), 3.0); }
        CTOrganAECFilter& organAECFilter() { return m_organ; }
        const CTOrganAECFilter& organAECFilter() const { return m_organ; }

    protected:
        using BeamBase::BeamBase;
    };
    // single-tube group
    class CTSingleTubeBase : public CTBeamBase {
    public:
        double scanFieldOfView() const { return m_d.fov; }
        void setScanFieldOfView(double f) { m_d.fov = std::max(std::abs(f), 1.0); }
        const BowtieFilter& bowtieFilter() const { return m_bowtie[0]; }
        void setBowtieFilter(const BowtieFilter& f) { m_bowtie[0] = f; }
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

    protected:
        using CTBeamBase::CTBeamBase;
    };
}

template <bool ENABLETRACKING = false>
class CTSequentialBeam : public detail::CTSingleTubeBase {
public:
    CTSequentialBeam(const std::array<double, 3>& start = { 0, 0, 0 }, const std::array<double, 3>& normal = { 0, 0, 1 },
        const std::map<std::size_t, double>& filtrationMaterials = {})
        : detail::CTSingleTubeBase(DXB_BEAM_CT_SEQUENTIAL)
    {
        setPosition(start);
        setScanNormal(normal);
        m_tube[0].setFiltrationMaterials(filtrationMaterials);
    }
    std::array<double, 3> position() const { return get3(m_d.position); }
    void setPosition(const std::array<double, 3>& p) { set3(m_d.position, p); }
    std::array<double, 3> scanNormal() const { return get3(m_d.direction); }
    void setScanNormal(const std::array<double, 3>& n) { set3(m_d.direction, vectormath::normalized(n)); }
    std::uint64_t numberOfSlices() const { return m_d.n_slices; }
    void setNumberOfSlices(std::uint64_t n) { m_d.n_slices = n > 0 ? n : 1; }
    double sliceSpacing() const { return m_d.slice_spacing; }
    void setSliceSpacing(double s) { m_d.slice_spacing = std::abs(s); }
    double CTDIw() const { return m_d.ctdi; }
    void setCTDIw(double v) { m_d.ctdi = std::abs(v); }
};
}
