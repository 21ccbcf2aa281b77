// dxmc/transportprogress.hpp — dxmc::TransportProgress as OpenDXMC uses it: progress() -> (n, total), message(),
// continueSimulation(), setStopSimulation() (R:src/libopendxmc/simulationpipeline.cpp:37,114-119,139,169,234,261;
// member of the pipeline object at R:src/libopendxmc/simulationpipeline.hpp:65).  Lock-free: polled every 3 s by one
// thread while another runs the transport and a third may request a stop.
#pragma once
#include "../dxb.h"
#include <cstdint>
#include <string>
#include <utility>
namespace dxmc {
class TransportProgress {
public:
    TransportProgress()
        : m_h(dxb_progress_create())
    {
    }
    ~TransportProgress() { dxb_progress_destroy(m_h); }
    TransportProgress(const TransportProgress&) = delete;
    TransportProgress& operator=(const TransportProgress&) = delete;
    void start(std::uint64_t) { dxb_progress_reset(m_h); }
    std::pair<std::uint64_t, std::uint64_t> progress() const
    {
        std::uint64_t d = 0, t = 0;
        dxb_progress_read(m_h, &d, &t);
        // the driver computes (n * 100) / total from its timer (R:src/libopendxmc/simulationpipeline.cpp:114-115): never 0
        return { d, t > 0 ? t : 1 };
    }
    std::string message() const
    {
        char buf[128];
        dxb_progress_message(m_h, buf, sizeof(buf));
        return buf;
    }
    bool continueSimulation() const { return dxb_progress_continue(m_h) != 0; }
    void setStopSimulation() { dxb_progress_stop(m_h); }
    dxb_progress* handle() const { return m_h; }

private:
    dxb_progress* m_h;
};
}
