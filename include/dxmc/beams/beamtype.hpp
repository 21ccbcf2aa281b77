// dxmc/beams/beamtype.hpp — what every beam of the shim offers to dxmc::Transport: numberOfExposures(),
// numberOfParticles(), exposure(i) and (shim-specific) desc(), the C-ABI description of the beam.
// Exposure mirrors what OpenDXMC reads from DXMClib exposures: position(), directionCosines(),
// collimationHalfAngles() (R:src/libopendxmc/beamactorcontainer.cpp:113-195).
#pragma once
#include "../../dxb.h"
#include "../constants.hpp"
#include "../vectormath.hpp"
#include "filters/bowtiefilter.hpp"
#include "filters/ctaecfilter.hpp"
#include "filters/ctorganaecfilter.hpp"
#include "tube/tube.hpp"
#include <array>
#include <concepts>
#include <cstdint>
#include <map>
#include <vector>
namespace dxmc {

class Exposure {
public:
    explicit Exposure(const dxb_exposure& e)
        : m_e(e)
    {
    }
    std::array<double, 3> position() const { return { m_e.position[0], m_e.position[1], m_e.position[2] }; }
    std::array<std::array<double, 3>, 2> directionCosines() const
    {
        return { { { m_e.cosines[0][0], m_e.cosines[0][1], m_e.cosines[0][2] }, { m_e.cosines[1][0], m_e.cosines[1][1], m_e.cosines[1][2] } } };
    }
    std::array<double, 3> direction() const { return { m_e.direction[0], m_e.direction[1], m_e.direction[2] }; }
    std::array<double, 2> collimationHalfAngles() const { return { m_e.half_angles[0], m_e.half_angles[1] }; }
    double weight() const { return m_e.weight; }
    std::uint64_t numberOfParticles() const { return m_e.n_particles; }

private:
    dxb_exposure m_e;
};

template <typename B>
concept BeamType = requires(const B& beam, std::uint64_t i) {
    { beam.numberOfExposures() } -> std::convertible_to<std::uint64_t>;
    { beam.numberOfParticles() } -> std::convertible_to<std::uint64_t>;
    { beam.exposure(i) } -> std::same_as<Exposure>;
    { beam.desc() } -> std::same_as<const dxb_beam_desc&>;
};

namespace detail {
    // state shared by all shim beams: the POD description + the objects whose arrays it points to
    class BeamBase {
    public:
        std::uint64_t numberOfExposures() const { return dxb_beam_number_of_exposures(&desc()); }
        std::uint64_t numberOfParticlesPerExposure() const { return m_d.particles_per_exposure; }
        void setNumberOfParticlesPerExposure(std::uint64_t n) { m_d.particles_per_exposure = n; }
        std::uint64_t numberOfParticles() const { return numberOfExposures() * m_d.particles_per_exposure; }
        Exposure exposure(std::uint64_t i) const
        {
            dxb_exposure e {};
            dxb_beam_exposure(&desc(), i, &e);
            return Exposure(e);
        }
        // the C-ABI view; pointers stay valid until the beam is modified or destroyed
        const dxb_beam_desc& desc() const
        {
            sync();
            return m_d;
        }

    protected:
        explicit BeamBase(int type) { dxb_beam_desc_init(&m_d, type); }
        BeamBase(const BeamBase& o) { *this = o; }
        BeamBase& operator=(const BeamBase& o)
        {
            m_d = o.m_d;
            for (int t = 0; t < 2; ++t) {
                m_tube[t] = o.m_tube[t];
                m_bowtie[t] = o.m_bowtie[t];
            }
            m_aec = o.m_aec;
            m_organ = o.m_organ;
            m_nTubes = o.m_nTubes;
            return *this;
        }
        void sync() const
        {
            for (int t = 0; t < 2; ++t) {
                if (t < m_nTubes) {
                    m_specE[t] = m_tube[t].getEnergy();
                    m_specW[t] = m_tube[t].getSpecter(m_specE[t], true);
                    m_d.spectrum[t].n = static_cast<uint32_t>(m_specE[t].size());
                    m_d.spectrum[t].energy_kev = m_specE[t].data();
                    m_d.spectrum[t].weight = m_specW[t].data();
                } else {
                    m_d.spectrum[t] = dxb_spectrum {};
                }
                m_d.bowtie[t] = m_bowtie[t].desc();
            }
            m_d.aec = m_aec.desc();
            m_d.organ_aec = m_organ.desc();
        }
        static void set3(double (&dst)[3], const std::array<double, 3>& v)
        {
            for (int i = 0; i < 3; ++i)
                dst[i] = v[i];
        }
        static std::array<double, 3> get3(const double (&src)[3]) { return { src[0], src[1], src[2] }; }

        mutable dxb_beam_desc m_d;
        Tube m_tube[2];
        BowtieFilter m_bowtie[2];
        CTAECFilter m_aec;
        CTOrganAECFilter m_organ;
        int m_nTubes = 1;
        mutable std::vector<double> m_specE[2], m_specW[2];
    };
}
}
