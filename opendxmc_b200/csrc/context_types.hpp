// context_types.hpp — the state behind a dxb_ctx, shared by context.cu (C ABI: world, beams, run, read-out) and
// exchange.cu (multi-GPU: slab upload, tally exchange, distributed dose score).
#pragma once
#include "kernels.hpp"
#include "physics.hpp"
#include "internal.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <condition_variable>
#include <functional>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

struct dxb_progress {
    std::atomic<uint64_t> done { 0 }, total { 0 };
    std::atomic<int> stop { 0 };
    std::atomic<int64_t> start_ns { 0 };
};

namespace dxb {

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    int device = -1;
    bool owned = true; // false: caller-provided storage (dxb_set_tally_storage), never freed here
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release()
    {
        if (p && owned) {
            int cur = 0;
            cudaGetDevice(&cur);
            if (device >= 0)
                cudaSetDevice(device);
            cudaFree(p);
            cudaSetDevice(cur);
        }
        p = nullptr;
        n = 0;
        owned = true;
    }
    void adopt(T* ptr, size_t count, int dev)
    {
        release();
        p = ptr;
        n = count;
        device = dev;
        owned = false;
    }
    cudaError_t alloc(size_t count, int dev)
    {
        if (p && n == count && device == dev)
            return cudaSuccess;
        release();
        device = dev;
        if (count == 0)
            return cudaSuccess;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T));
        if (e == cudaSuccess)
            n = count;
        return e;
    }
    template <typename H>
    cudaError_t upload(const std::vector<H>& h, int dev, cudaStream_t s)
    {
        static_assert(sizeof(H) == sizeof(T), "size mismatch");
        cudaError_t e = alloc(h.size(), dev);
        if (e != cudaSuccess || h.empty())
            return e;
        return cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s);
    }
};

// One voxel world resident on one device: grid + the tables of its materials.
struct World {
    int device = 0;
    uint64_t dim[3] = { 0, 0, 0 };
    double spacing[3] = { 1, 1, 1 };
    double center[3] = { 0, 0, 0 };
    size_t nvox = 0;
    int n_mat = 0;
    DevBuf<unsigned int> voxels;
    DevBuf<unsigned long long> tally;  // 4 words / voxel (buffer 0)
    DevBuf<unsigned long long> tally1; // buffer 1: exchanging contexts double-buffer the tallies (exchange.cu)
    int cur = 0;                       // buffer the next / last transport scores into
    unsigned long long* tallyCur() const { return cur ? tally1.p : tally.p; }
    DevBuf<float4> att;
    DevBuf<float> tot, etr, majorant, ffcdf, sf;
    DevBuf<ShellDev> shells;
    DevBuf<int> nshells;
    DevBuf<float> restJ0;
    DevBuf<unsigned int> maxDensityBits; // [257]: per-material max density bits, [256] = largest material index seen
    DevBuf<double> stageDensity;         // staging for the caller's f64 density / u8 material (kept between set_grid calls)
    DevBuf<unsigned char> stageMaterial;
    bool hasGrid = false, hasTables = false;
    // slab-local majorants (pool kernel, LM builds)
    std::vector<float> hostTot;          // [n_mat * kDevNE] the f32 total-attenuation table as uploaded
    DevBuf<unsigned int> slabMax;        // [lmSlabs * 256] per slab and material: largest density (24-bit float bits)
    DevBuf<float> lmRatio;            // [lmSlabs * kLmBands]
    std::vector<float> lmHost;           // host copy of lmRatio: local / global majorant, in (0, 1]
    int lmShift = 0, lmSlabs = 0;
    bool lmUseful = false;               // the table predicts a gain (mean ratio below the threshold)
    // dense box (pool kernel, DB builds): bounding box of the voxels that are not thin, outside ratios
    bool dbBuilt = false, dbUseful = false;
    int dbBox[6] = { 0, 0, 0, 0, 0, 0 };          // first voxel index x y z, one past the last x y z
    float dbFaces[6] = { 0, 0, 0, 0, 0, 0 };      // the same as coordinates [cm]: low x y z, high x y z
    float dbRatio[16] = { 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1 };
    DevBuf<unsigned int> dbScratch;               // [256 thin limits][256 outside maxima][6 box words]
    // brick pre-filter (pool kernel, quad step): upper bounds of mu / mu_max per brick and energy octave
    DevBuf<unsigned char> brickBound;    // [brickN[0] * brickN[1] * brickN[2] * 8]
    int brickShift = 0, brickN[3] = { 0, 0, 0 };

    GridDev gridDev() const
    {
        GridDev g;
        g.nx = static_cast<int>(dim[0]);
        g.ny = static_cast<int>(dim[1]);
        g.nz = static_cast<int>(dim[2]);
        const double hx = 0.5 * dim[0] * spacing[0], hy = 0.5 * dim[1] * spacing[1], hz = 0.5 * dim[2] * spacing[2];
        g.x0 = static_cast<float>(center[0] - hx);
        g.y0 = static_cast<float>(center[1] - hy);
        g.z0 = static_cast<float>(center[2] - hz);
        g.x1 = static_cast<float>(center[0] + hx);
        g.y1 = static_cast<float>(center[1] + hy);
        g.z1 = static_cast<float>(center[2] + hz);
        g.inv_dx = static_cast<float>(1.0 / spacing[0]);
        g.inv_dy = static_cast<float>(1.0 / spacing[1]);
        g.inv_dz = static_cast<float>(1.0 / spacing[2]);
        g.offx = -g.x0 * g.inv_dx;
        g.offy = -g.y0 * g.inv_dy;
        g.offz = -g.z0 * g.inv_dz;
        g.voxels = voxels.p;
        g.tally = tallyCur();
        return g;
    }
    TablesDev tablesDev() const
    {
        TablesDev t;
        t.n_mat = n_mat;
        t.att = att.p;
        t.tot = tot.p;
        t.etr = etr.p;
        t.majorant = majorant.p;
        t.ffcdf = ffcdf.p;
        t.sf = sf.p;
        t.shells = shells.p;
        t.n_shells = nshells.p;
        t.rest_j0 = restJ0.p;
        return t;
    }
};

struct DeviceState {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool ownStream = true;
    World world;                 // the patient grid
    std::unique_ptr<World> ctdi; // nested calibration phantom (device 0 only)
    double ctdiDiameter = 0;
    DevBuf<double> dose, variance;
    DevBuf<unsigned long long> events;
    DevBuf<ExposureDev> exposures;
    DevBuf<float> specProb[2], bowAngle[2], bowWeight[2];
    DevBuf<unsigned short> specAlias[2];
    uint64_t uploadedBeam = 0;           // hash of the beam whose arrays the buffers above hold (0: unknown)
    DevBuf<unsigned long long> counters; // [0] work cursor, [8..14] stats
    cudaEvent_t evStart = nullptr, evTransport = nullptr, evEnd = nullptr;
    // nested CTDI run: hole id per phantom voxel (-1: none) and the five integer kerma sums
    DevBuf<signed char> ctdiHole;
    DevBuf<unsigned long long> holeSums;
    // ---- tally exchange (exchange.cu): this device is participant `part` of `dxb_ctx::parts`
    int part = 0;
    size_t vb = 0, ve = 0;             // voxel slab of the dose score this device owns
    cudaStream_t xstream = nullptr;    // exchange stream: peer pulls (copy engines), slab reduce -> dose, tally clear
    cudaEvent_t evBufReady[2] = { nullptr, nullptr };      // tally buffer b is clean again
    cudaEvent_t evPullsDone[2] = { nullptr, nullptr };     // this device has pulled its slab of every peer's buffer b
    cudaEvent_t evTransportDone[2] = { nullptr, nullptr }; // the transport kernels scoring into buffer b have finished
    cudaEvent_t evTimer[2] = { nullptr, nullptr };         // dxb_timer_begin / dxb_timer_end
    cudaEvent_t evX[4] = { nullptr, nullptr, nullptr, nullptr }; // exchange stream: before pulls, after pulls, after reduce, before the lazy clear
    bool pullsPending[2] = { false, false };
    bool needsClear[2] = { false, false };
    DevBuf<unsigned long long> staging; // (parts - 1) slabs of 4 x u64 per voxel
    std::vector<const unsigned long long*> peerTally[2]; // [buffer][participant]: device pointers valid on THIS device
    // one process per GPU: {sequence number, five hole sums} of this rank's shard of the nested calibration run, readable
    // by the peers through CUDA IPC
    DevBuf<unsigned long long> mailbox;
    std::vector<const unsigned long long*> peerMailbox;
};

// a beam expanded for the device: exposures, alias tables, bowtie knots (context.cu: prepareBeam)
struct PreparedBeam {
    std::vector<ExposureDev> exposures;
    uint64_t ppe = 0, nTotal = 0;
    std::vector<float> prob[2], bowA[2], bowW[2];
    std::vector<unsigned short> alias[2];
    int specN[2] = { 1, 1 };
    float specE0[2] = { 0, 0 }, specStep[2] = { 1, 1 };
    double maxWeight = 1.0;
    uint64_t hash = 0; // beamHash of the descriptor it was prepared from (0: not cached)
};

// CT calibration phantom (host side), cached per diameter
struct CtdiPhantom {
    double diameter = 0;
    uint64_t dim[3];
    double spacing[3];
    std::vector<double> density;
    std::vector<uint8_t> material; // 0 air, 1 PMMA, 2 measurement air
    std::vector<signed char> hole; // per voxel: -1 or hole id 0..4 (central 10 cm only)
    size_t holeCount[5] = { 0, 0, 0, 0, 0 };
    std::vector<std::shared_ptr<Material>> mats;
};

struct Options {
    uint64_t batch = 1ull << 27; // local histories per launch
    int threads = 256;
    int blocksPerSm = 0;         // 0: occupancy query
    int tableInSmem = 1;
    int slots = 0;               // 0 = one photon per lane in registers (transport.cu, default: fastest);
                                 // 2/3/4/6 = lane-multiplexed photons in shared memory (transport_mux.cu)
    int refillThreshold = -1;    // warp phase machine; -1 = the default of the selected kernel
    int interactThreshold = -1;
    int rayleighThreshold = -1;
    int poolSlots = 16;          // > 0: block-pooled kernel (transport_pool.cu, default) with this many slots per lane class;
                                 // 0: `slots` selects the register kernel (0) or the lane-multiplexed one
    int smemPadKb = 0;           // experiment: extra dynamic shared memory per block (shrinks L1)
    int stepPairs = 0;           // pool / mux kernels: step pairs per step phase (0: kernel default, pool 2, mux 1)
    int poolThreads = 256;       // pool kernel: threads per block (the block shares one photon pool)
    int poolMinBlocks = 0;       // pool kernel: 5 / 6 select the 48 / 40-register builds (more resident warps), else 64 registers
    int stepQuad = 1;            // pool kernel, step_pairs == 2: issue the four gathers of both pairs at once
    int diag = 0;                // pool kernel: count phase executions / claimed lanes (slower; printed to stderr)
    int brickFilter = 0;         // pool kernel, quad step: skip the gathers of certainly-virtual collisions (bit-identical results;
                                 // 4.5x fewer gathers on C2 but no faster: off by default)
    int brickVoxels = 16;        // brick edge in voxels (a power of two)
    int denseBox = -1;           // pool kernel: dense-box tracking; -1 auto (on when the box is a small enough part of the grid), 0 off, 1 on
    double denseTheta = 0.02;    // a voxel is thin when its attenuation stays below this fraction of the majorant at every energy
    int localMajorant = -1;      // pool kernel: slab-local majorants; -1 auto (on when the table predicts a gain), 0 off, 1 on
    double slabCm = 8.0;         // target slab thickness [cm] (rounded to a power-of-two number of voxel layers; profiles/r02_sweep.txt)
    int serviceWarps = 4;        // pool kernel: warps per block preferring interaction / Rayleigh / refill phases
    int interactBias = -999;     // -999: kernel default (mux 16: interaction phase when waiting lanes + bias >= stepping lanes;
                                 // pool 24: stepper warps keep stepping while at least this many lanes can claim a photon)
};

} // namespace dxb

namespace dxb {

// One resident host thread per additional device of a context (the calling thread serves device 0): the multi-GPU paths
// issue their per-device driver calls from these, bound to their device once, instead of spawning threads per call.
class DeviceWorkers {
public:
    explicit DeviceWorkers(const std::vector<int>& devices) // devices[k]: CUDA ordinal served by worker k (device index k + 1)
    {
        for (size_t k = 0; k < devices.size(); ++k)
            threads.emplace_back([this, k, dev = devices[k]]() { loop(k + 1, dev); });
    }
    ~DeviceWorkers()
    {
        {
            std::lock_guard<std::mutex> lock(m);
            stop = true;
        }
        cvWork.notify_all();
        for (auto& t : threads)
            t.join();
    }
    size_t size() const { return threads.size(); }
    void post(const std::function<void(size_t)>* f)
    {
        {
            std::lock_guard<std::mutex> lock(m);
            task = f;
            remaining = threads.size();
            ++generation;
        }
        cvWork.notify_all();
    }
    void wait()
    {
        std::unique_lock<std::mutex> lock(m);
        cvDone.wait(lock, [this]() { return remaining == 0; });
        task = nullptr;
    }

private:
    void loop(size_t index, int device)
    {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(size_t)>* f = nullptr;
            {
                std::unique_lock<std::mutex> lock(m);
                cvWork.wait(lock, [&]() { return stop || generation != seen; });
                if (stop)
                    return;
                seen = generation;
                f = task;
            }
            cudaSetDevice(device);
            (*f)(index);
            {
                std::lock_guard<std::mutex> lock(m);
                --remaining;
            }
            cvDone.notify_one();
        }
    }
    std::mutex m;
    std::condition_variable cvWork, cvDone;
    uint64_t generation = 0;
    size_t remaining = 0;
    bool stop = false;
    const std::function<void(size_t)>* task = nullptr;
    std::vector<std::thread> threads;
};

} // namespace dxb

struct dxb_ctx {
    std::vector<std::unique_ptr<dxb::DeviceState>> devs;
    std::unique_ptr<dxb::DeviceWorkers> workers; // created by the first overDevices() of a multi-device context
    std::vector<std::shared_ptr<dxb::Material>> materials;
    uint64_t seed = 0x0DDC0FFEEull;   // base Philox key (dxb_set_seed)
    uint64_t beamCounter = 0;         // dxb_run_transport calls since the last dxb_set_seed
    uint64_t beamKey = 0x0DDC0FFEEull; // key of the last beam: seed + counter * DXB_BEAM_KEY_STRIDE
    uint64_t runKey = 0x0DDC0FFEEull;  // key the next kernel launch uses (beam key, or its calibration stream)
    uint64_t rank = 0, world = 1;
    uint64_t calibHistories = 36000000ull;
    dxb::Options opt;
    std::string error;
    dxb_run_stats stats {};
    float scaleE = 16777216.0f, scaleE2 = 65536.0f; // 2^24, 2^16 fixed-point quanta per keV, keV^2
    int smCount = 148;
    bool tallyValid = false;
    std::unique_ptr<dxb::PreparedBeam> lastBeam; // the last main beam, expanded (reused when the same beam runs again)
    std::shared_ptr<void> scene;            // the save file dxb_load_scene read (an h5mini::File), written back by dxb_save_dose
    std::unique_ptr<dxb::CtdiPhantom> ctdi; // host copy of the calibration phantom, built once per diameter
    // ---- tally exchange (exchange.cu).  In-process: one participant per device of this context; one process per GPU:
    // one participant per rank, peers mapped through CUDA IPC (dxb_exchange_export / dxb_exchange_import).
    bool exchanging = false; // double-buffered tallies, pipelined exchange, distributed dose score
    bool ipc = false;        // participants are other processes (host barriers by the caller instead of events)
    int parts = 1;           // participants
    bool exchanged = false;  // the last beam's tallies have been handed to the exchange (no longer readable)
    std::vector<void*> ipcOpened;
    struct {                 // parameters of the exchange dxb_finish_beam asked for (read by every device's mgExchangeOnDevice)
        bool active = false;
        int buffer = 0;
        double factor = 0;
        float scaleE = 1, scaleE2 = 1;
    } pending;
    bool exchangeTimed = false;
    std::chrono::steady_clock::time_point lastReturn {}; // DXB_TRACE_HOST=1: host time between two dxb_run_transport calls
    uint64_t mailSeq = 0;    // calibrated beams so far (the mailbox sequence number)
    double exchangeMs[4] = { 0, 0, 0, 0 }; // last flush: pulls, reduce, clear (CUDA events on the exchange stream of device 0)
};

namespace dxb {

// DXB_TRACE_HOST=1: dxb_run_transport prints its host-side phase times to stderr (where does the time between kernels go)
inline bool traceHost()
{
    static const bool on = [] { const char* e = std::getenv("DXB_TRACE_HOST"); return e && e[0] == '1'; }();
    return on;
}

// (the multi-device paths report errors from one host thread per device)
inline std::mutex& errorMutex()
{
    static std::mutex m;
    return m;
}
inline int fail(dxb_ctx* c, int code, const std::string& msg)
{
    if (c) {
        std::lock_guard<std::mutex> lock(errorMutex());
        c->error = msg;
    }
    return code;
}

#define CUDA_TRY(ctx, expr)                                                                        \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            cudaGetLastError();                                                                    \
            return fail(ctx, DXB_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));       \
        }                                                                                          \
    } while (0)

// f(device index) on one host thread per device (each bound to its device; the caller's thread serves device 0): host-to-
// device and device-to-host copies of pageable caller memory are staged by the driver and block the calling thread, so N
// links need N threads - and N launch sequences issued side by side start the N GPUs together.  Not re-entrant.
template <typename F>
inline int overDevices(dxb_ctx* c, F f)
{
    const size_t n = c->devs.size();
    std::vector<int> rc(n, DXB_OK);
    if (n > 1 && (!c->workers || c->workers->size() != n - 1)) {
        std::vector<int> devices;
        for (size_t i = 1; i < n; ++i)
            devices.push_back(c->devs[i]->device);
        c->workers = std::make_unique<DeviceWorkers>(devices);
    }
    const std::function<void(size_t)> task = [&](size_t i) { rc[i] = f(i); };
    if (n > 1)
        c->workers->post(&task);
    if (cudaSetDevice(c->devs[0]->device) != cudaSuccess)
        rc[0] = fail(c, DXB_ECUDA, "cudaSetDevice failed");
    else
        rc[0] = f(0);
    if (n > 1)
        c->workers->wait();
    cudaSetDevice(c->devs[0]->device);
    for (int r : rc)
        if (r != DXB_OK)
            return r;
    return DXB_OK;
}

// ---- context.cu helpers used by exchange.cu
int finishGrid(dxb_ctx* c, World& w, cudaStream_t s);
int uploadGrid(dxb_ctx* c, World& w, const uint64_t dim[3], const double spacing[3], const double* density, const uint8_t* material,
    cudaStream_t s, size_t begin, size_t end, bool finish);

// ---- exchange.cu: several GPUs behind one context, or one context per process with CUDA IPC peers
inline void slabOf(size_t n, int part, int parts, size_t& b, size_t& e)
{
    b = n * static_cast<size_t>(part) / static_cast<size_t>(parts);
    e = n * (static_cast<size_t>(part) + 1) / static_cast<size_t>(parts);
}
int mgInit(dxb_ctx* c);                                   // peer access, exchange streams and events
void mgDestroy(dxb_ctx* c);
int mgSetGrid(dxb_ctx* c, const uint64_t dim[3], const double spacing[3], const double* density, const uint8_t* material);
int mgPrepareExchange(dxb_ctx* c);                        // second tally buffer, staging, slabs (after the grid is known)
int mgEnqueueExchange(dxb_ctx* c, double factor);         // pulls + slab reduce -> dose + clear, asynchronous
int mgExchangeOnDevice(dxb_ctx* c, DeviceState& d);       // one device's share of the noted exchange
int mgEnqueuePending(dxb_ctx* c);                         // the noted exchange on every device (side by side on the device threads)
int mgFlush(dxb_ctx* c);                                  // every enqueued exchange has completed
int mgGetDose(dxb_ctx* c, size_t begin, size_t end, double* dose, double* variance, uint64_t* events);
int mgGatherDose(dxb_ctx* c);                             // device 0 receives every slab of the dose score
int mgSumTallies(dxb_ctx* c, DevBuf<unsigned long long>& out); // device 0: sum over the devices' current tally buffers
// one process per GPU: publishes this rank's five calibration hole sums and adds those of all peers (waits for them)
int mgShareHoleSums(dxb_ctx* c, const unsigned long long* devSums, unsigned long long total[5]);

} // namespace dxb
