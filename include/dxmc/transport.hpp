// dxmc/transport.hpp — dxmc::Transport: Transport(), setNumberOfThreads(int), operator()(world, beam, progress,
// useBeamCalibration) (R:src/libopendxmc/simulationpipeline.cpp:155-165).  The call blocks until every history of the
// beam has run on the GPU(s), the calibration factor is known and the dose score has been updated — like DXMClib's.
#pragma once
#include "../dxb.h"
#include "beams/beamtype.hpp"
#include "transportprogress.hpp"
#include "world/world.hpp"
#include <cstdint>
#include <stdexcept>
#include <string>
namespace dxmc {
class Transport {
public:
    Transport() = default;
    // CPU worker threads have no meaning on the GPU path; kept so that the driver compiles unchanged
    void setNumberOfThreads(std::uint64_t n) { m_threads = n; }
    std::uint64_t numberOfThreads() const { return m_threads; }

    template <typename Item, BeamType B>
    bool operator()(World<Item>& world, const B& beam, TransportProgress* progress = nullptr, bool useBeamCalibration = true) const
    {
        // DXMClib's Transport starts the progress object itself (progress->start(beam.numberOfParticles())):
        // the driver raises the stop flag at the end of every worker run (R:src/libopendxmc/simulationpipeline.cpp:234) and
        // reuses its single m_progress for the next startSimulation(), so nothing else ever clears it.
        if (progress)
            progress->start(dxb_beam_number_of_particles(&beam.desc()));
        const int rc = dxb_run(world.ctx(), &beam.desc(), Item::lowEnergyCorrection(), useBeamCalibration ? 1 : 0,
            progress ? progress->handle() : nullptr);
        world.item().invalidateDose();
        if (rc == DXB_ECANCELLED)
            return false;
        if (rc != DXB_OK)
            throw std::runtime_error(std::string("dxmc::Transport: ") + dxb_last_error(world.ctx()));
        return true;
    }
    template <typename Item, BeamType B>
    static bool run(World<Item>& world, const B& beam, TransportProgress* progress = nullptr, bool useBeamCalibration = true)
    {
        return Transport {}(world, beam, progress, useBeamCalibration);
    }

private:
    std::uint64_t m_threads = 0;
};
}
