// dxmc/beams/ctspiralbeam.hpp — dxmc::CTSpiralBeam<ENABLETRACKING>: ctor (start, stop, filtration) and accessors
// R:src/libopendxmc/beamsettingsmodel.cpp:1158-1378; translation via start/stop R:src/libopendxmc/beamactorcontainer.cpp:88-90.
#pragma once
#include "ctsequentialbeam.hpp"
namespace dxmc {
template <bool ENABLETRACKING = false>
class CTSpiralBeam : public detail::CTSingleTubeBase {
public:
    CTSpiralBeam(const std::array<double, 3>& start = { 0, 0, 0 }, const std::array<double, 3>& stop = { 0, 0, 1 },
        const std::map<std::size_t, double>& filtrationMaterials = {})
        : detail::CTSingleTubeBase(DXB_BEAM_CT_SPIRAL)
    {
        setStartStopPosition(start, stop);
        m_tube[0].setFiltrationMaterials(filtrationMaterials);
    }
    std::array<double, 3> startPosition() const { return get3(m_d.start); }
    std::array<double, 3> stopPosition() const { return get3(m_d.stop); }
    void setStartPosition(const std::array<double, 3>& p) { set3(m_d.start, p); }
    void setStopPosition(const std::array<double, 3>& p) { set3(m_d.stop, p); }
    void setStartStopPosition(const std::array<double, 3>& a, const std::array<double, 3>& b)
    {
        setStartPosition(a);
        setStopPosition(b);
    }
    double pitch() const { return m_d.pitch; }
    void setPitch(double p) { m_d.pitch = std::max(std::abs(p), 0.01); }
    double CTDIvol() const { return m_d.ctdi; }
    void setCTDIvol(double v) { m_d.ctdi = std::abs(v); }
    const CTAECFilter& AECFilter() const { return m_aec; }
    void setAECFilter(const CTAECFilter& f) { m_aec = f; }
    void setAECFilterData(const std::array<double, 3>& start, const std::array<double, 3>& stop, const std::vector<double>& w) { m_aec.setData(start, stop, w); }
};
}
