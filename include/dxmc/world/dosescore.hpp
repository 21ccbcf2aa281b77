// dxmc/world/dosescore.hpp — what doseScored(i) returns: dose() [mGy], variance() [mGy^2], numberOfEvents()
// (R:src/libopendxmc/simulationpipeline.cpp:178,204,219).
#pragma once
#include <cmath>
#include <cstdint>
namespace dxmc {
class DoseScore {
public:
    DoseScore() = default;
    DoseScore(double dose, double variance, std::uint64_t events)
        : m_dose(dose)
        , m_variance(variance)
        , m_events(events)
    {
    }
    double dose() const { return m_dose; }
    double variance() const { return m_variance; }
    double standardDeviation() const { return std::sqrt(m_variance); }
    double relativeUncertainty() const { return m_dose > 0 ? 1.96 * std::sqrt(m_variance) / m_dose : 0.0; }
    std::uint64_t numberOfEvents() const { return m_events; }

private:
    double m_dose = 0, m_variance = 0;
    std::uint64_t m_events = 0;
};
}
