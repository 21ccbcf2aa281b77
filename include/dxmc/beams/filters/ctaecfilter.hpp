// dxmc/beams/filters/ctaecfilter.hpp — CTAECFilter(start, stop, weights), setData, isEmpty, weights, start, stop,
// length, size (R:src/libopendxmc/datacontainer.cpp:37,59; ctimageimportpipeline.cpp:120; ctaecplot.cpp:80-97;
// hdf5wrapper.cpp:448-455).
#pragma once
#include "../../../dxb.h"
#include <array>
#include <cmath>
#include <vector>
namespace dxmc {
class CTAECFilter {
public:
    CTAECFilter() = default;
    CTAECFilter(const std::array<double, 3>& start, const std::array<double, 3>& stop, const std::vector<double>& weights) { setData(start, stop, weights); }
    void setData(const std::array<double, 3>& start, const std::array<double, 3>& stop, const std::vector<double>& weights)
    {
        m_start = start;
        m_stop = stop;
        m_weights = weights;
    }
    bool isEmpty() const { return m_weights.size() < 2; }
    const std::vector<double>& weights() const { return m_weights; }
    const std::array<double, 3>& start() const { return m_start; }
    const std::array<double, 3>& stop() const { return m_stop; }
    std::size_t size() const { return m_weights.size(); }
    double length() const
    {
        double s = 0;
        for (int i = 0; i < 3; ++i)
            s += (m_stop[i] - m_start[i]) * (m_stop[i] - m_start[i]);
        return std::sqrt(s);
    }
    double operator()(const std::array<double, 3>& pos) const
    {
        const dxb_aec d = desc();
        return dxb_aec_weight(&d, pos.data());
    }
    dxb_aec desc() const
    {
        dxb_aec d {};
        d.n = static_cast<uint32_t>(m_weights.size());
        for (int i = 0; i < 3; ++i) {
            d.start[i] = m_start[i];
            d.stop[i] = m_stop[i];
        }
        d.weights = m_weights.data();
        return d;
    }

private:
    std::array<double, 3> m_start { 0, 0, 0 }, m_stop { 0, 0, 0 };
    std::vector<double> m_weights;
};
}
