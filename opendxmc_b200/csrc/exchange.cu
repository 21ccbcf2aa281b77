// exchange.cu — several B200s behind ONE dxb_ctx (the reference is one process looping over beams,
// R:src/libopendxmc/simulationpipeline.cpp:161-167), and the same exchange between one-process-per-GPU contexts
// whose tally buffers are mapped into each other through CUDA IPC.
//
// Histories shard over the participants (dxb_shard_*), every participant holds a full replica of the packed voxel grid
// and the tables, and the per-beam fixed-point tallies have to be summed once per beam.  Here that one
// exchange step is taken OFF the critical path:
//   * the tallies are double-buffered: beam i scores into buffer i & 1;
//   * after the transport of beam i, participant r PULLS its 1/N voxel slab of every peer's buffer with the copy
//     engines (cudaMemcpyPeerAsync over NVLink 5 / NVSwitch: no SM is involved, so the pulls run underneath the transport
//     kernels of beam i + 1, which occupy every SM);
//   * one streaming kernel then adds the N slabs (64-bit integers: order independent, bitwise identical to one GPU),
//     converts energy to dose for the slab and accumulates it into r's part of the dose score; the buffer is cleared
//     lazily, at the head of the next exchange, when every peer is known to have pulled from it (mgExchangeOnDevice).
// In-process, the per-device driver calls are issued side by side by the context's resident device threads (overDevices).
// The dose score stays distributed (slab r on participant r) until it is read out: every device copies its own slab to
// the caller's arrays over its own PCIe link, concurrently.  The grid upload is sharded the same way: device r uploads
// and packs slab r of the caller's arrays, the packed 4-byte slabs are all-gathered by peer copies.
#include "context_types.hpp"

#include <chrono>
#include <thread>

namespace dxb {

namespace {

int g_exchangeBlocks = 148 * 8;

// sum of the local slab and `n_staged` pulled slabs -> dose score of voxels [begin, end)
//   DoseScore::addScoredEnergy (recalled): dose += E k / (rho V);  var += var_E (k/(rho V))^2, var_E = sum E^2 - (sum E)^2 / n
__global__ void reduceSlabsToDoseKernel(const unsigned long long* __restrict__ local, const unsigned long long* __restrict__ staging,
    int n_staged, size_t slab_stride_voxels, const unsigned int* __restrict__ voxels, double* __restrict__ dose,
    double* __restrict__ variance, unsigned long long* __restrict__ events, size_t begin, size_t end, double inv_scale_e,
    double inv_scale_e2, double factor, double voxel_volume)
{
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = begin + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < end; i += stride) {
        const ulonglong2* q = reinterpret_cast<const ulonglong2*>(local + i * 4);
        ulonglong2 a = __ldcs(q), b = __ldcs(q + 1);
        unsigned long long se = a.x, se2 = a.y, sn = b.x;
        const size_t j = i - begin;
        for (int k = 0; k < n_staged; ++k) {
            q = reinterpret_cast<const ulonglong2*>(staging + (static_cast<size_t>(k) * slab_stride_voxels + j) * 4);
            a = __ldcs(q);
            b = __ldcs(q + 1);
            se += a.x;
            se2 += a.y;
            sn += b.x;
        }
        if (sn == 0)
            continue;
        const double rho = static_cast<double>(voxelDensity(voxels[i]));
        if (!(rho > 0.0))
            continue;
        const double e = static_cast<double>(se) * inv_scale_e;
        const double e2 = static_cast<double>(se2) * inv_scale_e2;
        const double nn = static_cast<double>(sn);
        const double varE = fmax(0.0, e2 - e * e / nn);
        const double f = factor / (rho * voxel_volume);
        dose[i] += e * f;
        variance[i] += varE * f * f;
        events[i] += sn;
    }
}

#define CUDA_TRY_T(ctx, expr) CUDA_TRY(ctx, expr)

// mailbox = {sequence number, five sums}: the sums are visible system-wide before the sequence number announces them
__global__ void publishMailboxKernel(unsigned long long* __restrict__ mailbox, const unsigned long long* __restrict__ sums, unsigned long long seq)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        for (int k = 0; k < 5; ++k)
            mailbox[1 + k] = sums[k];
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long*>(mailbox) = seq;
    }
}

size_t maxSlab(size_t n, int parts) { return (n + static_cast<size_t>(parts) - 1) / static_cast<size_t>(parts); }

} // namespace

int mgInit(dxb_ctx* c)
{
    const size_t n = c->devs.size();
    g_exchangeBlocks = c->smCount * 8;
    for (size_t i = 0; i < n; ++i) {
        DeviceState& d = *c->devs[i];
        CUDA_TRY(c, cudaSetDevice(d.device));
        for (size_t j = 0; j < n; ++j) {
            if (i == j)
                continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, d.device, c->devs[j]->device);
            if (!can)
                return fail(c, DXB_ECUDA, "devices " + std::to_string(d.device) + " and " + std::to_string(c->devs[j]->device) + " cannot access each other's memory");
            const cudaError_t e = cudaDeviceEnablePeerAccess(c->devs[j]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return fail(c, DXB_ECUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            cudaGetLastError();
        }
        d.part = static_cast<int>(i);
        CUDA_TRY(c, cudaStreamCreateWithFlags(&d.xstream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            CUDA_TRY(c, cudaEventCreateWithFlags(&d.evBufReady[b], cudaEventDisableTiming));
            CUDA_TRY(c, cudaEventCreateWithFlags(&d.evPullsDone[b], cudaEventDisableTiming));
            CUDA_TRY(c, cudaEventCreateWithFlags(&d.evTransportDone[b], cudaEventDisableTiming));
            CUDA_TRY(c, cudaEventCreate(&d.evTimer[b]));
        }
        for (int k = 0; k < 4; ++k)
            CUDA_TRY(c, cudaEventCreate(&d.evX[k]));
    }
    c->parts = static_cast<int>(n);
    c->exchanging = n > 1;
    CUDA_TRY(c, cudaSetDevice(c->devs[0]->device));
    return DXB_OK;
}

void mgDestroy(dxb_ctx* c)
{
    for (auto& d : c->devs) {
        cudaSetDevice(d->device);
        if (d->xstream) {
            cudaStreamSynchronize(d->xstream);
            cudaStreamDestroy(d->xstream);
        }
        for (int b = 0; b < 2; ++b) {
            for (cudaEvent_t e : { d->evBufReady[b], d->evPullsDone[b], d->evTransportDone[b], d->evTimer[b] })
                if (e)
                    cudaEventDestroy(e);
        }
        for (int k = 0; k < 4; ++k)
            if (d->evX[k])
                cudaEventDestroy(d->evX[k]);
    }
    if (!c->devs.empty())
        cudaSetDevice(c->devs[0]->device);
    for (void* p : c->ipcOpened)
        cudaIpcCloseMemHandle(p);
    c->ipcOpened.clear();
}

// AAVoxelGrid::setData + World::build over the devices of one context: device r uploads and packs slab r of the caller's
// arrays (9 B/voxel over ITS PCIe link, 1/N of the grid), the packed slabs are all-gathered with peer copies, the
// per-material density maxima are max-reduced, every device builds the majorant.
int mgSetGrid(dxb_ctx* c, const uint64_t dim[3], const double spacing[3], const double* density, const uint8_t* material)
{
    const size_t n = static_cast<size_t>(dim[0]) * dim[1] * dim[2];
    const int parts = static_cast<int>(c->devs.size());
    int rc = mgFlush(c);
    if (rc != DXB_OK)
        return rc;
    rc = overDevices(c, [&](size_t i) -> int {
        DeviceState& d = *c->devs[i];
        size_t b, e;
        slabOf(n, static_cast<int>(i), parts, b, e);
        int r = uploadGrid(c, d.world, dim, spacing, density, material, d.stream, b, e, false);
        if (r != DXB_OK)
            return r;
        CUDA_TRY_T(c, d.dose.alloc(n, d.device));
        CUDA_TRY_T(c, d.variance.alloc(n, d.device));
        CUDA_TRY_T(c, d.events.alloc(n, d.device));
        CUDA_TRY_T(c, cudaMemsetAsync(d.dose.p, 0, n * sizeof(double), d.stream));
        CUDA_TRY_T(c, cudaMemsetAsync(d.variance.p, 0, n * sizeof(double), d.stream));
        CUDA_TRY_T(c, cudaMemsetAsync(d.events.p, 0, n * sizeof(unsigned long long), d.stream));
        CUDA_TRY_T(c, cudaStreamSynchronize(d.stream));
        return DXB_OK;
    });
    if (rc != DXB_OK)
        return rc;
    // density maxima / largest material index: element-wise maximum over the devices (non-negative floats order like their bits)
    std::vector<unsigned int> mx(257, 0u), tmp(257);
    for (auto& d : c->devs) {
        CUDA_TRY(c, cudaSetDevice(d->device));
        CUDA_TRY(c, cudaMemcpy(tmp.data(), d->world.maxDensityBits.p, 257 * sizeof(unsigned int), cudaMemcpyDeviceToHost));
        for (int k = 0; k < 257; ++k)
            mx[k] = std::max(mx[k], tmp[k]);
    }
    // all-gather of the packed slabs: every device pulls the slabs of its peers (staggered, so that no device serves two readers)
    for (int i = 0; i < parts; ++i) {
        DeviceState& d = *c->devs[i];
        CUDA_TRY(c, cudaSetDevice(d.device));
        CUDA_TRY(c, cudaMemcpyAsync(d.world.maxDensityBits.p, mx.data(), 257 * sizeof(unsigned int), cudaMemcpyHostToDevice, d.stream));
        for (int k = 1; k < parts; ++k) {
            const int p = (i + k) % parts;
            size_t b, e;
            slabOf(n, p, parts, b, e);
            if (e > b)
                CUDA_TRY(c, cudaMemcpyPeerAsync(d.world.voxels.p + b, d.device, c->devs[p]->world.voxels.p + b, c->devs[p]->device,
                                (e - b) * sizeof(unsigned int), d.stream));
        }
    }
    rc = overDevices(c, [&](size_t i) -> int { // majorant, material-index check; synchronises the device's stream
        return finishGrid(c, c->devs[i]->world, c->devs[i]->stream);
    });
    if (rc != DXB_OK)
        return rc;
    CUDA_TRY(c, cudaSetDevice(c->devs[0]->device));
    return mgPrepareExchange(c);
}

int mgPrepareExchange(dxb_ctx* c)
{
    const size_t n = c->devs[0]->world.nvox;
    const int parts = c->parts;
    const size_t ms = maxSlab(n, parts);
    for (auto& dp : c->devs) {
        DeviceState& d = *dp;
        World& w = d.world;
        CUDA_TRY(c, cudaSetDevice(d.device));
        CUDA_TRY(c, w.tally1.alloc(n * 4, d.device));
        CUDA_TRY(c, d.staging.alloc(static_cast<size_t>(parts - 1) * ms * 4, d.device));
        slabOf(n, d.part, parts, d.vb, d.ve);
        CUDA_TRY(c, cudaMemsetAsync(w.tally.p, 0, n * 4 * sizeof(unsigned long long), d.xstream));
        CUDA_TRY(c, cudaMemsetAsync(w.tally1.p, 0, n * 4 * sizeof(unsigned long long), d.xstream));
        for (int b = 0; b < 2; ++b) {
            CUDA_TRY(c, cudaEventRecord(d.evBufReady[b], d.xstream));
            d.pullsPending[b] = false;
            d.needsClear[b] = false;
        }
        w.cur = 0;
    }
    c->pending.active = false;
    for (auto& dp : c->devs) { // (all devices clear concurrently)
        CUDA_TRY(c, cudaSetDevice(dp->device));
        CUDA_TRY(c, cudaStreamSynchronize(dp->xstream));
    }
    if (!c->ipc) {
        for (auto& dp : c->devs)
            for (int b = 0; b < 2; ++b) {
                dp->peerTally[b].assign(parts, nullptr);
                for (int p = 0; p < parts; ++p)
                    dp->peerTally[b][p] = b ? c->devs[p]->world.tally1.p : c->devs[p]->world.tally.p;
            }
    }
    c->exchanged = false;
    CUDA_TRY(c, cudaSetDevice(c->devs[0]->device));
    return DXB_OK;
}

// The exchange of the beam that was just finished, enqueued at once on every participant's exchange stream (in-process: by
// the per-device host threads side by side, ~15 driver calls each); it runs underneath the next beam's transport kernels.
// One process per GPU: the caller has just put a barrier between the ranks' transport and this call.
int mgEnqueueExchange(dxb_ctx* c, double factor)
{
    c->pending.active = true;
    c->pending.buffer = c->devs[0]->world.cur;
    c->pending.factor = factor;
    c->pending.scaleE = c->scaleE;
    c->pending.scaleE2 = c->scaleE2;
    for (auto& dp : c->devs)
        dp->world.cur = c->pending.buffer ^ 1;
    c->exchanged = true;
    return mgEnqueuePending(c);
}

// One device's share of the noted exchange, enqueued on its exchange stream: clear the OTHER buffer (lazily - see below),
// pull this device's voxel slab of buffer b from every peer, reduce the slabs to dose.  No cross-device event is involved:
//   * buffer b is complete on every participant - dxb_run_transport returned after synchronising the transport streams of
//     all devices of the context, and with one process per GPU the caller separates it from dxb_finish_beam by a barrier;
//   * buffer b^1 (the beam before) has been pulled by every peer - each participant waits for its own pulls at the end of
//     dxb_run_transport (they ran under that beam's kernels), before the same barrier / before this call.
int mgExchangeOnDevice(dxb_ctx* c, DeviceState& d)
{
    const int parts = c->parts;
    const int b = c->pending.buffer;
    World& w = d.world;
    const size_t n = w.nvox;
    const size_t ms = maxSlab(n, parts);
    const double vol = w.spacing[0] * w.spacing[1] * w.spacing[2];
    cudaStream_t xs = d.xstream;
    CUDA_TRY(c, cudaSetDevice(d.device));
    CUDA_TRY(c, cudaStreamWaitEvent(xs, d.evTransportDone[b], 0)); // (free: the host has already seen it)
    CUDA_TRY(c, cudaEventRecord(d.evX[3], xs));
    if (d.needsClear[b ^ 1]) {
        unsigned long long* other = (b ^ 1) ? w.tally1.p : w.tally.p;
        CUDA_TRY(c, cudaMemsetAsync(other, 0, n * 4 * sizeof(unsigned long long), xs));
        CUDA_TRY(c, cudaEventRecord(d.evBufReady[b ^ 1], xs));
        d.needsClear[b ^ 1] = false;
    }
    CUDA_TRY(c, cudaEventRecord(d.evX[0], xs));
    const size_t bytes = (d.ve - d.vb) * 4 * sizeof(unsigned long long);
    for (int k = 1; k < parts && bytes > 0; ++k) {
        const int p = (d.part + k) % parts; // staggered: at any moment every participant serves one reader
        const unsigned long long* src = d.peerTally[b][p] + d.vb * 4;
        unsigned long long* dst = d.staging.p + static_cast<size_t>(k - 1) * ms * 4;
        CUDA_TRY(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, xs));
    }
    CUDA_TRY(c, cudaEventRecord(d.evPullsDone[b], xs));
    CUDA_TRY(c, cudaEventRecord(d.evX[1], xs));
    d.pullsPending[b] = true;
    const unsigned long long* local = b ? w.tally1.p : w.tally.p;
    if (d.ve > d.vb)
        reduceSlabsToDoseKernel<<<g_exchangeBlocks, 256, 0, xs>>>(local, d.staging.p, parts - 1, ms, w.voxels.p, d.dose.p, d.variance.p,
            d.events.p, d.vb, d.ve, 1.0 / c->pending.scaleE, 1.0 / c->pending.scaleE2, c->pending.factor, vol);
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaEventRecord(d.evX[2], xs));
    d.needsClear[b] = true; // cleared by the next exchange
    return DXB_OK;
}

int mgEnqueuePending(dxb_ctx* c)
{
    if (!c->pending.active)
        return DXB_OK;
    const int rc = c->devs.size() > 1 ? overDevices(c, [&](size_t i) -> int { return mgExchangeOnDevice(c, *c->devs[i]); })
                                      : mgExchangeOnDevice(c, *c->devs[0]);
    if (rc != DXB_OK)
        return rc;
    c->pending.active = false;
    c->exchangeTimed = true;
    CUDA_TRY(c, cudaSetDevice(c->devs[0]->device));
    return DXB_OK;
}

int mgFlush(dxb_ctx* c)
{
    if (!c->exchanging)
        return DXB_OK;
    const int prc = mgEnqueuePending(c);
    if (prc != DXB_OK)
        return prc;
    for (auto& dp : c->devs) {
        if (!dp->xstream)
            continue;
        CUDA_TRY(c, cudaSetDevice(dp->device));
        CUDA_TRY(c, cudaStreamSynchronize(dp->xstream));
        dp->pullsPending[0] = dp->pullsPending[1] = false;
    }
    if (c->exchangeTimed) {
        DeviceState& d0 = *c->devs[0];
        float ms = 0;
        if (cudaEventElapsedTime(&ms, d0.evX[0], d0.evX[1]) == cudaSuccess)
            c->exchangeMs[0] = ms;
        if (cudaEventElapsedTime(&ms, d0.evX[1], d0.evX[2]) == cudaSuccess)
            c->exchangeMs[1] = ms;
        if (cudaEventElapsedTime(&ms, d0.evX[3], d0.evX[0]) == cudaSuccess)
            c->exchangeMs[2] = ms; // the lazy clear of the other buffer (0 when there was nothing to clear)
        cudaGetLastError();
        c->exchangeTimed = false;
    }
    CUDA_TRY(c, cudaSetDevice(c->devs[0]->device));
    return DXB_OK;
}

// doseScored(i) for voxels [begin, end): every device copies the part of ITS slab that falls into the range
int mgGetDose(dxb_ctx* c, size_t begin, size_t end, double* dose, double* variance, uint64_t* events)
{
    int rc = mgFlush(c);
    if (rc != DXB_OK)
        return rc;
    return overDevices(c, [&](size_t i) -> int {
        DeviceState& d = *c->devs[i];
        const size_t b = std::max(begin, d.vb), e = std::min(end, d.ve);
        if (e <= b)
            return DXB_OK;
        const size_t m = e - b;
        if (dose)
            CUDA_TRY_T(c, cudaMemcpyAsync(dose + b, d.dose.p + b, m * sizeof(double), cudaMemcpyDeviceToHost, d.stream));
        if (variance)
            CUDA_TRY_T(c, cudaMemcpyAsync(variance + b, d.variance.p + b, m * sizeof(double), cudaMemcpyDeviceToHost, d.stream));
        if (events)
            CUDA_TRY_T(c, cudaMemcpyAsync(events + b, d.events.p + b, m * sizeof(uint64_t), cudaMemcpyDeviceToHost, d.stream));
        CUDA_TRY_T(c, cudaStreamSynchronize(d.stream));
        return DXB_OK;
    });
}

// the whole dose score on device 0 (post-processing, per-organ dose: O(N) kernels that read all of it)
int mgGatherDose(dxb_ctx* c)
{
    int rc = mgFlush(c);
    if (rc != DXB_OK || c->ipc || c->devs.size() < 2)
        return rc;
    DeviceState& d0 = *c->devs[0];
    CUDA_TRY(c, cudaSetDevice(d0.device));
    for (size_t i = 1; i < c->devs.size(); ++i) {
        DeviceState& p = *c->devs[i];
        const size_t m = p.ve - p.vb;
        if (m == 0)
            continue;
        CUDA_TRY(c, cudaMemcpyPeerAsync(d0.dose.p + p.vb, d0.device, p.dose.p + p.vb, p.device, m * sizeof(double), d0.stream));
        CUDA_TRY(c, cudaMemcpyPeerAsync(d0.variance.p + p.vb, d0.device, p.variance.p + p.vb, p.device, m * sizeof(double), d0.stream));
        CUDA_TRY(c, cudaMemcpyPeerAsync(d0.events.p + p.vb, d0.device, p.events.p + p.vb, p.device, m * sizeof(unsigned long long), d0.stream));
    }
    CUDA_TRY(c, cudaStreamSynchronize(d0.stream));
    return DXB_OK;
}

// the last beam's tallies summed over the devices of the context, on device 0 (dxb_get_energy_scored before the exchange)
int mgSumTallies(dxb_ctx* c, DevBuf<unsigned long long>& out)
{
    DeviceState& d0 = *c->devs[0];
    const size_t words = d0.world.nvox * 4;
    CUDA_TRY(c, cudaSetDevice(d0.device));
    CUDA_TRY(c, out.alloc(words, d0.device));
    CUDA_TRY(c, cudaMemcpyAsync(out.p, d0.world.tallyCur(), words * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, d0.stream));
    std::vector<const unsigned long long*> peers;
    for (size_t i = 1; i < c->devs.size(); ++i)
        peers.push_back(c->devs[i]->world.tallyCur());
    DevBuf<const unsigned long long*> dPeers;
    CUDA_TRY(c, dPeers.upload(peers, d0.device, d0.stream));
    launchPeerReduce(out.p, dPeers.p, static_cast<int>(peers.size()), words, d0.stream);
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaStreamSynchronize(d0.stream));
    return DXB_OK;
}

int mgShareHoleSums(dxb_ctx* c, const unsigned long long* devSums, unsigned long long total[5])
{
    DeviceState& d = *c->devs[0];
    if (!c->ipc || !d.mailbox.p || d.peerMailbox.size() != static_cast<size_t>(c->parts))
        return fail(c, DXB_ESTATE, "calibration exchange: no mailboxes (dxb_exchange_import)");
    CUDA_TRY(c, cudaSetDevice(d.device));
    const unsigned long long seq = ++c->mailSeq;
    publishMailboxKernel<<<1, 32, 0, d.stream>>>(d.mailbox.p, devSums, seq);
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaStreamSynchronize(d.stream));
    for (int k = 0; k < 5; ++k)
        total[k] = 0;
    const auto t0 = std::chrono::steady_clock::now();
    for (int p = 0; p < c->parts; ++p) {
        unsigned long long h[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
        for (;;) {
            CUDA_TRY(c, cudaMemcpy(h, d.peerMailbox[p], sizeof(h), cudaMemcpyDeviceToHost));
            if (h[0] == seq) {
                // the sums were made visible before the number: read them (again) now that the number is known
                CUDA_TRY(c, cudaMemcpy(h, d.peerMailbox[p], sizeof(h), cudaMemcpyDeviceToHost));
                break;
            }
            if (h[0] > seq)
                return fail(c, DXB_ESTATE, "calibration exchange: rank " + std::to_string(p) + " is ahead (the ranks must finish the same beams with the same flags)");
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 60.0)
                return fail(c, DXB_ESTATE, "calibration exchange: timed out waiting for rank " + std::to_string(p));
        }
        for (int k = 0; k < 5; ++k)
            total[k] += h[1 + k];
    }
    return DXB_OK;
}

} // namespace dxb

using namespace dxb;

// ============================================================================ C ABI: one process per GPU over CUDA IPC
extern "C" {

int dxb_exchange_export(dxb_ctx* c, void* handles)
{
    if (!c || !handles)
        return DXB_EINVAL;
    if (c->devs.size() != 1)
        return fail(c, DXB_ESTATE, "exchange_export: one device per context (one process per GPU)");
    DeviceState& d = *c->devs[0];
    World& w = d.world;
    if (w.nvox == 0 || !w.tally.p)
        return fail(c, DXB_ESTATE, "exchange_export: set the grid first");
    if (!w.tally.owned)
        return fail(c, DXB_ESTATE, "exchange_export: the tally buffer is caller-provided storage (dxb_set_tally_storage)");
    CUDA_TRY(c, cudaSetDevice(d.device));
    if (!d.xstream) {
        const int rc = mgInit(c);
        if (rc != DXB_OK)
            return rc;
    }
    CUDA_TRY(c, w.tally1.alloc(w.nvox * 4, d.device));
    // the mailbox gets an allocation of its own (2 MiB: not carved out of a block shared with other small buffers)
    CUDA_TRY(c, d.mailbox.alloc((2u << 20) / sizeof(unsigned long long), d.device));
    CUDA_TRY(c, cudaMemset(d.mailbox.p, 0, 64));
    static_assert(3 * sizeof(cudaIpcMemHandle_t) == DXB_EXCHANGE_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h[3];
    CUDA_TRY(c, cudaIpcGetMemHandle(&h[0], w.tally.p));
    CUDA_TRY(c, cudaIpcGetMemHandle(&h[1], w.tally1.p));
    CUDA_TRY(c, cudaIpcGetMemHandle(&h[2], d.mailbox.p));
    std::memcpy(handles, h, sizeof(h));
    return DXB_OK;
}

int dxb_exchange_import(dxb_ctx* c, uint64_t rank, uint64_t world, const void* all_handles)
{
    if (!c || !all_handles || world == 0 || rank >= world || world > 64)
        return fail(c, DXB_EINVAL, "exchange_import: need rank < world <= 64 and the gathered handles");
    if (c->devs.size() != 1 || !c->devs[0]->world.tally1.p)
        return fail(c, DXB_ESTATE, "exchange_import: call dxb_exchange_export first");
    DeviceState& d = *c->devs[0];
    World& w = d.world;
    CUDA_TRY(c, cudaSetDevice(d.device));
    for (void* p : c->ipcOpened)
        cudaIpcCloseMemHandle(p);
    c->ipcOpened.clear();
    const auto* h = static_cast<const cudaIpcMemHandle_t*>(all_handles);
    for (int b = 0; b < 2; ++b)
        d.peerTally[b].assign(world, nullptr);
    d.peerMailbox.assign(world, nullptr);
    for (uint64_t p = 0; p < world; ++p) {
        for (int b = 0; b < 3; ++b) {
            const unsigned long long* ptr = b == 0 ? w.tally.p : (b == 1 ? w.tally1.p : d.mailbox.p);
            if (p != rank) {
                void* opened = nullptr;
                CUDA_TRY(c, cudaIpcOpenMemHandle(&opened, h[p * 3 + b], cudaIpcMemLazyEnablePeerAccess));
                c->ipcOpened.push_back(opened);
                ptr = static_cast<const unsigned long long*>(opened);
            }
            if (b < 2)
                d.peerTally[b][p] = ptr;
            else
                d.peerMailbox[p] = ptr;
        }
    }
    c->mailSeq = 0;
    CUDA_TRY(c, cudaMemset(d.mailbox.p, 0, 64));
    c->ipc = true;
    c->parts = static_cast<int>(world);
    c->exchanging = world > 1;
    d.part = static_cast<int>(rank);
    c->rank = rank;
    c->world = world;
    return mgPrepareExchange(c);
}

int dxb_exchange_close(dxb_ctx* c)
{
    if (!c || c->devs.empty())
        return DXB_EINVAL;
    if (!c->ipc)
        return DXB_OK;
    const int rc = mgFlush(c);
    CUDA_TRY(c, cudaSetDevice(c->devs[0]->device));
    for (void* p : c->ipcOpened)
        cudaIpcCloseMemHandle(p);
    c->ipcOpened.clear();
    c->ipc = false;
    c->exchanging = false;
    c->parts = 1;
    c->devs[0]->part = 0;
    c->devs[0]->world.cur = 0;
    c->tallyValid = false;
    return rc;
}

int dxb_flush(dxb_ctx* c)
{
    if (!c || c->devs.empty())
        return DXB_EINVAL;
    for (auto& d : c->devs) {
        CUDA_TRY(c, cudaSetDevice(d->device));
        CUDA_TRY(c, cudaStreamSynchronize(d->stream));
    }
    return mgFlush(c);
}

int dxb_timer_begin(dxb_ctx* c)
{
    if (!c || c->devs.empty())
        return DXB_EINVAL;
    int rc = dxb_flush(c);
    if (rc != DXB_OK)
        return rc;
    for (auto& d : c->devs) {
        CUDA_TRY(c, cudaSetDevice(d->device));
        if (!d->evTimer[0]) {
            CUDA_TRY(c, cudaEventCreate(&d->evTimer[0]));
            CUDA_TRY(c, cudaEventCreate(&d->evTimer[1]));
        }
        CUDA_TRY(c, cudaEventRecord(d->evTimer[0], d->stream));
    }
    CUDA_TRY(c, cudaSetDevice(c->devs[0]->device));
    return DXB_OK;
}

int dxb_timer_end(dxb_ctx* c, double* ms_max)
{
    if (!c || c->devs.empty() || !ms_max)
        return DXB_EINVAL;
    // an exchange that is still only noted (in-process) belongs to the timed work
    const int prc = mgEnqueuePending(c);
    if (prc != DXB_OK)
        return prc;
    double mx = 0;
    for (auto& d : c->devs) {
        if (!d->evTimer[0])
            return fail(c, DXB_ESTATE, "timer_end without timer_begin");
        CUDA_TRY(c, cudaSetDevice(d->device));
        // the exchange stream's work is part of the job: the end mark follows it
        if (d->xstream) {
            CUDA_TRY(c, cudaEventRecord(d->evTimer[1], d->xstream));
            CUDA_TRY(c, cudaStreamWaitEvent(d->stream, d->evTimer[1], 0));
        }
        CUDA_TRY(c, cudaEventRecord(d->evTimer[1], d->stream));
        CUDA_TRY(c, cudaEventSynchronize(d->evTimer[1]));
        float ms = 0;
        CUDA_TRY(c, cudaEventElapsedTime(&ms, d->evTimer[0], d->evTimer[1]));
        mx = std::max(mx, static_cast<double>(ms));
    }
    CUDA_TRY(c, cudaSetDevice(c->devs[0]->device));
    *ms_max = mx;
    return DXB_OK;
}

int dxb_exchange_times(const dxb_ctx* c, double out_ms[3])
{
    if (!c || !out_ms)
        return DXB_EINVAL;
    for (int k = 0; k < 3; ++k)
        out_ms[k] = c->exchangeMs[k];
    return DXB_OK;
}

} // extern "C"
