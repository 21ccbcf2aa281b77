// dxmc/beams/filters/ctorganaecfilter.hpp — CTOrganAECFilter (R:src/libopendxmc/beamsettingsmodel.cpp:349-437,
// R:src/libopendxmc/hdf5wrapper.cpp:576-580,896-920).
#pragma once
#include "../../../dxb.h"
#include "../../constants.hpp"
#include <algorithm>
#include <cmath>
namespace dxmc {
class CTOrganAECFilter {
public:
    CTOrganAECFilter()
    {
        m_d.use_filter = 0;
        m_d.compensate_outside = 0;
        m_d.start_angle = 0;
        m_d.stop_angle = PI_VAL();
        m_d.ramp_angle = 20 * DEG_TO_RAD();
        m_d.low_weight = 0.6;
    }
    bool useFilter() const { return m_d.use_filter != 0; }
    void setUseFilter(bool on) { m_d.use_filter = on ? 1 : 0; }
    bool compensateOutside() const { return m_d.compensate_outside != 0; }
    void setCompensateOutside(bool on) { m_d.compensate_outside = on ? 1 : 0; }
    double startAngle() const { return m_d.start_angle; }
    void setStartAngle(double a) { m_d.start_angle = a; }
    double startAngleDeg() const { return m_d.start_angle * RAD_TO_DEG(); }
    void setStartAngleDeg(double a) { m_d.start_angle = a * DEG_TO_RAD(); }
    double stopAngle() const { return m_d.stop_angle; }
    void setStopAngle(double a) { m_d.stop_angle = a; }
    double stopAngleDeg() const { return m_d.stop_angle * RAD_TO_DEG(); }
    void setStopAngleDeg(double a) { m_d.stop_angle = a * DEG_TO_RAD(); }
    double rampAngle() const { return m_d.ramp_angle; }
    void setRampAngle(double a) { m_d.ramp_angle = std::abs(a); }
    double rampAngleDeg() const { return m_d.ramp_angle * RAD_TO_DEG(); }
    void setRampAngleDeg(double a) { m_d.ramp_angle = std::abs(a) * DEG_TO_RAD(); }
    double lowWeight() const { return m_d.low_weight; }
    void setLowWeightFactor(double w) { m_d.low_weight = std::clamp(w, 0.0, 1.0); }
    double maxWeight() const { return dxb_organ_aec_max_weight(&m_d); }
    double operator()(double angle) const { return dxb_organ_aec_weight(&m_d, angle); }
    const dxb_organ_aec& desc() const { return m_d; }

private:
    dxb_organ_aec m_d {};
};
}
