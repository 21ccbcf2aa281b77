// dxmc/beams/filters/bowtiefilter.hpp — BowtieFilter(vector<pair<angle_rad, weight>>)
// (R:src/libopendxmc/bowtiefilterreader.cpp:74-93; set on beams at R:src/libopendxmc/beamsettingsmodel.cpp:1311).
#pragma once
#include "../../../dxb.h"
#include <utility>
#include <vector>
namespace dxmc {
class BowtieFilter {
public:
    BowtieFilter() = default;
    explicit BowtieFilter(const std::vector<std::pair<double, double>>& angleWeight)
    {
        for (const auto& [a, w] : angleWeight) {
            m_angle.push_back(a);
            m_weight.push_back(w);
        }
    }
    double operator()(double angle) const
    {
        const dxb_bowtie d = desc();
        return dxb_bowtie_weight(&d, angle);
    }
    std::vector<std::pair<double, double>> data() const
    {
        std::vector<std::pair<double, double>> r;
        for (std::size_t i = 0; i < m_angle.size(); ++i)
            r.emplace_back(m_angle[i], m_weight[i]);
        return r;
    }
    dxb_bowtie desc() const
    {
        dxb_bowtie d {};
        d.n = static_cast<uint32_t>(m_angle.size() >= 2 ? m_angle.size() : 0);
        d.angle_rad = m_angle.data();
        d.weight = m_weight.data();
        return d;
    }

private:
    std::vector<double> m_angle, m_weight;
};
}
