// transport.cu — the photon-history kernels for sm_100a.
//
// One launch = a batch of histories of one beam.  Histories are handed to warps from one global cursor
// (256-history pieces), so concurrently resident histories belong to the same / adjacent exposures and the
// primary-beam slab of the voxel grid stays in the 126 MB L2.
//
// Replaces (recalled DXMClib): Transport::runWorker -> exposure.sampleParticle
// -> World::transport -> AAVoxelGrid::woodcockTransport -> interactions::interact -> EnergyScore.
// The algorithm and its random-number protocol are restated independently in oracle/oracle.cpp.
//
// No tensor cores: nothing here is a dense contraction.  The hot loop is one dependent 8-byte voxel gather
// per tentative step (HBM/L2 sector bound) + FP32/INT issue (Philox, log, lerp); see DESIGN.md §Kernels.
//
// Random-number protocol (mirrored draw for draw by the oracle): stream(history) = Philox4x32-10 with
// key = seed and counter = (history lo, history hi, block, 0).  A history consumes whole BLOCKS of four
// 24-bit uniforms, one block per event, so that every lane of a warp that executes an event generates its
// block at the same time (no divergent generator refills):
//   blocks 0-1  source sampling (angles, spectrum alias lookup)
//   then, from block 2 on, one block per
//     step pair        w0,w1 = length / acceptance of tentative step A, w2,w3 = the same for step B
//                      (B is only used if A was a virtual collision; otherwise w2,w3 are discarded)
//     interaction try  w0 = channel choice (first try of an interaction only), w1,w2 = Compton candidate /
//                      rejection test, w3 = azimuth
//     Rayleigh try     w0 = form-factor CDF target (Thomson: rejection variable), w1 = rejection test (Thomson:
//                      polar angle), w2 = azimuth
//     roulette         w0
#include "transport_common.cuh"

namespace dxb {

namespace {

// ------------------------------------------------------------------ the history kernel
//
// Warp-level phase machine.  Every lane owns at most one photon and is in one of these states:
//   STEP      tentative Woodcock steps (cheap, executed by most lanes every iteration),
//   WAIT_NEW  a tentative collision was accepted as real; channel not chosen yet,
//   WAIT_C    Compton chosen, previous candidate rejected; waits for the next interaction phase,
//   WAIT_R    Rayleigh chosen; waits for a Rayleigh phase,
//   DEAD      no photon; waiting for a new history.
// Each iteration the warp picks ONE phase by vote: the (expensive) interaction code only runs once
// `interact_threshold` lanes wait for it (or nobody can step); every execution of it performs exactly one
// sampling try per lane, rejected lanes simply stay in their WAIT state, so rejection loops never run with a
// handful of lanes.  Dead lanes are only refilled once `refill_threshold` of them are dead.  Source sampling is
// warp-cooperative: all 32 lanes sample one history each, photons that hit the grid are compacted into a
// per-warp shared-memory buffer, and dead lanes pop from it.  (profiles/r01_kernel_history.md has the ncu
// numbers that drove this structure: 5.4 -> 12.8 -> ... active lanes per instruction.)
struct Photon {
    float px, py, pz;
    float dx, dy, dz;
    float E, w;
};

// state values double as vote increments: one __reduce_add_sync gives all lane counts (dead: bits 0-7,
// waiting for an interaction try: bits 8-15, waiting for a Rayleigh try: bits 16-23; bit 30 only tags WAIT_C)
constexpr int kStStep = 0, kStDead = 1, kStWaitNew = 0x100, kStWaitC = 0x100 | (1 << 30), kStWaitR = 0x10000;
constexpr int kWarpBufFloats = kSourceBufWords * 32; // px py pz dx dy dz E w histOffset epos.i epos.f muMax
static_assert(kSourceBufWords == 12, "source buffer layout");

template <int MODE, bool CALIB, bool SMEM_TABLE>
__global__ void __launch_bounds__(256, 4) transportKernel(const __grid_constant__ RunParams P)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nWarps = blockDim.x >> 5;
    // layout: see transportSmemBytes (device_types.cuh)
    unsigned long long* __restrict__ spool = reinterpret_cast<unsigned long long*>(s_raw) + warp * 4; // poolNext, poolEnd, bufBase
    float* __restrict__ s_dyn = reinterpret_cast<float*>(s_raw + nWarps * 32);
    float* __restrict__ sbuf = s_dyn + warp * kWarpBufFloats; // SoA: word f of entry k at sbuf[f * 32 + k]
    unsigned int* __restrict__ scnt = reinterpret_cast<unsigned int*>(s_dyn + nWarps * kWarpBufFloats) + threadIdx.x; // stride blockDim.x
    float* __restrict__ s_tot = s_dyn + nWarps * kWarpBufFloats + blockDim.x * kLaneCounters;
    for (int k = 0; k < kLaneCounters; ++k)
        scnt[k * blockDim.x] = 0u;
    if (lane == 0) {
        spool[0] = 0ull;
        spool[1] = 0ull;
        spool[2] = 0ull;
    }
    __syncwarp();
    if (SMEM_TABLE) {
        const int n = P.tab.n_mat * kDevNE;
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            s_tot[i] = P.tab.tot[i];
        __syncthreads();
    }
    const float* __restrict__ totTable = SMEM_TABLE ? s_tot : P.tab.tot;
    const GridDev& G = P.grid;
    // the warp's pool of local history indices (carved from the global cursor in 256-history pieces) lives in
    // shared memory (spool); only the refill phase touches it
    bool drained = false;
    int bufCount = 0;                 // entries in the warp's source buffer

    Photon ph;
    unsigned int hlo = 0, hhi = 0, blk = 2; // Philox counter of this lane's history
    int status = kStDead;
    TabPos epos;
    // majorant of the photon's energy, pre-multiplied: u * mu_max = k * muMaxU24 (k = 2^24 u), step = log2(1 - u) * stepScale
    float muMaxU24 = kU24, stepScale = -kLn2;
    unsigned int voxel = 0;
    int mat = 0;
    epos.i = 0;
    epos.f = 0.0f;
    ph.px = ph.py = ph.pz = ph.dx = ph.dy = ph.dz = ph.E = ph.w = 0.0f;

    unsigned int nSteps = 0; // the other per-lane counters live in shared memory (scnt)

    // after an accepted scatter: cut-off, roulette, majorant / exit distance refresh
    auto finishScatter = [&](float& edep, bool energyChanged) {
        bool alive = true;
        if (ph.E < kMinEnergy) {
            edep += ph.E * ph.w;
            ph.E = 0.0f;
            alive = false;
        } else if (ph.w < kRouletteThreshold) {
            const PhiloxBlock rb = philox4x32_10(P.round_key, hlo, hhi, blk++);
            if (rb.u(0) < kRouletteKill)
                alive = false;
            else
                ph.w *= 1.0f / (1.0f - kRouletteKill);
        }
        if (alive) {
            if (energyChanged) {
                epos = energyPos(ph.E);
                const float muMax = lerp(__ldg(P.tab.majorant + epos.i), __ldg(P.tab.majorant + epos.i + 1), epos.f);
                muMaxU24 = muMax * kU24;
                stepScale = -kLn2 * __fdividef(1.0f, muMax);
            }
            status = kStStep;
        } else {
            status = kStDead;
        }
    };

    for (;;) {
        const unsigned int votes = __reduce_add_sync(0xffffffffu, static_cast<unsigned int>(status) & 0x00ffffffu);
        const int nDead = votes & 0xff, nInt = (votes >> 8) & 0xff, nRay = (votes >> 16) & 0xff, nStep = 32 - nDead - nInt - nRay;
        const bool canRefill = !(drained && bufCount == 0);
        int phase; // 0 step, 1 interact, 2 refill, 3 rayleigh
        if (canRefill && nDead >= P.refill_threshold)
            phase = 2;
        else if (nInt >= P.interact_threshold)
            phase = 1;
        else if (nRay >= P.rayleigh_threshold)
            phase = 3;
        else if (nStep > 0)
            phase = 0;
        else if (nInt > 0)
            phase = 1;
        else if (nRay > 0)
            phase = 3;
        else if (canRefill)
            phase = 2;
        else
            break;

        if (phase == 0) {
            // ------------------------------------------------------------ a pair of tentative Woodcock steps
            float kermaA = 0.0f, kermaB = 0.0f;
            unsigned int voxA = 0, voxB = 0;
            if (status == kStStep) {
                const PhiloxBlock rb = philox4x32_10(P.round_key, hlo, hhi, blk++);
                const float sA = __log2f(fmaf(rb.k(0), -kU24, 1.0f)) * stepScale;
                const float sB = __log2f(fmaf(rb.k(2), -kU24, 1.0f)) * stepScale;
                const float ax = fmaf(ph.dx, sA, ph.px), ay = fmaf(ph.dy, sA, ph.py), az = fmaf(ph.dz, sA, ph.pz);
                const float bx = fmaf(ph.dx, sB, ax), by = fmaf(ph.dy, sB, ay), bz = fmaf(ph.dz, sB, az);
                // a step whose end point is outside the grid ends the history (B is only reached through A)
                const bool inA = voxelIndex(G, ax, ay, az, voxA);
                const bool inB = voxelIndex(G, bx, by, bz, voxB) && inA;
                // both gathers are issued before either is used: B is speculative (wasted if A turns out real)
                unsigned int cellA = 0u, cellB = 0u;
                if (inA)
                    cellA = loadVoxel(G.voxels + voxA);
                if (inB)
                    cellB = loadVoxel(G.voxels + voxB);
                if (!inA) {
                    status = kStDead; // left the grid
                } else {
                    ++nSteps;
                    const int matA = voxelMaterial(cellA);
                    const float* tt = totTable + matA * kDevNE + epos.i;
                    const float muA = voxelDensity(cellA) * lerp(tt[0], tt[1], epos.f);
                    if (CALIB && matA == P.score_material) {
                        // collision estimator of air kerma: every tentative collision carries 1/mu_max of track length
                        const float* et = P.tab.etr + matA * kDevNE + epos.i;
                        kermaA = ph.w * ph.E * lerp(__ldg(et), __ldg(et + 1), epos.f) * (stepScale * -kInvLn2);
                    }
                    if (rb.k(1) * muMaxU24 < muA) {
                        ph.px = ax;
                        ph.py = ay;
                        ph.pz = az;
                        voxel = voxA;
                        mat = matA;
                        status = kStWaitNew;
                    } else if (!inB) {
                        status = kStDead;
                    } else {
                        ++nSteps;
                        const int matB = voxelMaterial(cellB);
                        const float* tb = totTable + matB * kDevNE + epos.i;
                        const float muB = voxelDensity(cellB) * lerp(tb[0], tb[1], epos.f);
                        if (CALIB && matB == P.score_material) {
                            const float* et = P.tab.etr + matB * kDevNE + epos.i;
                            kermaB = ph.w * ph.E * lerp(__ldg(et), __ldg(et + 1), epos.f) * (stepScale * -kInvLn2);
                        }
                        ph.px = bx;
                        ph.py = by;
                        ph.pz = bz;
                        if (rb.k(3) * muMaxU24 < muB) {
                            voxel = voxB;
                            mat = matB;
                            status = kStWaitNew;
                        }
                    }
                }
            }
            if (CALIB) {
                unsigned int mScore = __ballot_sync(0xffffffffu, kermaA > 0.0f);
                if (kermaA > 0.0f)
                    scoreEnergy(mScore, G.tally, voxA, kermaA, P.tally_scale_e, P.tally_scale_e2);
                mScore = __ballot_sync(0xffffffffu, kermaB > 0.0f);
                if (kermaB > 0.0f)
                    scoreEnergy(mScore, G.tally, voxB, kermaB, P.tally_scale_e, P.tally_scale_e2);
            }
        } else if (phase == 1) {
            // ------------------------------------------------------------ one interaction try per waiting lane
            float edep = 0.0f;
            if (status == kStWaitNew || status == kStWaitC) {
                const PhiloxBlock rb = philox4x32_10(P.round_key, hlo, hhi, blk++);
                bool compton = status == kStWaitC;
                if (status == kStWaitNew) {
                    scnt[0] += 1u;
                    const float4 a = __ldg(P.tab.att + mat * kDevNE + epos.i);
                    const float4 b = __ldg(P.tab.att + mat * kDevNE + epos.i + 1);
                    const float aPhoto = lerp(a.x, b.x, epos.f);
                    const float aIncoh = lerp(a.y, b.y, epos.f);
                    const float aTot = lerp(a.w, b.w, epos.f);
                    const float r2 = rb.u(0) * aTot;
                    if (r2 < aPhoto) {
                        const float ef = MODE >= 2 ? photoFluorescence(P.tab, mat, ph.E, rb.u(1), rb.u(2)) : 0.0f;
                        if (MODE >= 2 && ef > 0.0f) {
                            // fluorescence photon: isotropic, one extra block for its direction
                            const PhiloxBlock rf = philox4x32_10(P.round_key, hlo, hhi, blk++);
                            isotropic(rf.u(0), rf.u(1), ph.dx, ph.dy, ph.dz);
                            edep = (ph.E - ef) * ph.w;
                            ph.E = ef;
                            finishScatter(edep, true);
                        } else {
                            edep = ph.E * ph.w;
                            ph.E = 0.0f;
                            status = kStDead;
                        }
                    } else if (r2 < aPhoto + aIncoh) {
                        compton = true;
                    } else {
                        status = kStWaitR;
                    }
                }
                if (compton) {
                    float e, cosT;
                    bool ok = comptonTry<MODE>(P.tab, mat, ph.E, rb.u(1), rb.u(2), e, cosT);
                    if (MODE >= 2 && ok) {
                        // impulse approximation: shell + Doppler broadening from one extra block
                        const PhiloxBlock ri = philox4x32_10(P.round_key, hlo, hhi, blk++);
                        ok = dopplerBroaden(P.tab, mat, ph.E, e, cosT, ri.u(0), ri.u(1), e);
                    }
                    if (ok) {
                        deflect(ph.dx, ph.dy, ph.dz, cosT, kTwoPi * rb.u(3));
                        const float E0 = ph.E;
                        ph.E = E0 * e;
                        edep = (E0 - ph.E) * ph.w;
                        finishScatter(edep, true);
                    } else {
                        status = kStWaitC;
                    }
                }
                if (CALIB)
                    edep = 0.0f;
            }
            if (!CALIB) {
                const unsigned int mScore = __ballot_sync(0xffffffffu, edep > 0.0f);
                if (edep > 0.0f) {
                    scnt[blockDim.x] += 1u;
                    scoreEnergy(mScore, G.tally, voxel, edep, P.tally_scale_e, P.tally_scale_e2);
                }
            }
        } else if (phase == 3) {
            // ------------------------------------------------------------ one Rayleigh try per waiting lane
            if (status == kStWaitR) {
                const PhiloxBlock rb = philox4x32_10(P.round_key, hlo, hhi, blk++);
                float cosT;
                if (rayleighTry<MODE>(P.tab, mat, ph.E, rb.u(0), rb.u(1), cosT)) {
                    deflect(ph.dx, ph.dy, ph.dz, cosT, kTwoPi * rb.u(2));
                    float edep = 0.0f; // Rayleigh deposits nothing (E >= cut-off here)
                    finishScatter(edep, false);
                }
            }
        } else {
            // ------------------------------------------------------------ refill
            const unsigned int laneLt = (1u << lane) - 1u;
            if (bufCount == 0) {
                // warp-cooperative source sampling of the next (up to) 32 histories
                unsigned long long poolNext = spool[0], poolEnd = spool[1];
                if (poolNext == poolEnd && !drained) {
                    constexpr unsigned long long kPiece = 256;
                    unsigned long long base = 0;
                    if (lane == 0)
                        base = atomicAdd(P.work_counter, kPiece);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    const unsigned long long start = P.local_begin + base;
                    if (start >= P.local_end) {
                        drained = true;
                    } else {
                        poolNext = start;
                        poolEnd = min(start + kPiece, P.local_end);
                    }
                }
                const unsigned long long avail = poolEnd - poolNext;
                const int nb = static_cast<int>(min(avail, 32ull));
                // local index -> global history id (65536-history blocks dealt round-robin over ranks); a piece never
                // straddles a shard block (256 divides 65536), so the 32 ids are consecutive
                const unsigned long long sblk = poolNext / kShardBlock;
                const unsigned long long bufBase = (sblk * P.world + P.rank) * kShardBlock + (poolNext % kShardBlock);
                const unsigned long long h = bufBase + lane;
                bool hit = false;
                Photon q;
                TabPos qpos;
                float qmu = 1.0f;
                qpos.i = 0;
                qpos.f = 0.0f;
                q.px = q.py = q.pz = q.dx = q.dy = q.dz = q.E = q.w = 0.0f;
                if (lane < nb && h < P.n_total) {
                    SourceSample ss;
                    hit = sampleSource(P, h, ss);
                    q.px = ss.px, q.py = ss.py, q.pz = ss.pz;
                    q.dx = ss.dx, q.dy = ss.dy, q.dz = ss.dz;
                    q.E = ss.E;
                    q.w = ss.w;
                    scnt[2 * blockDim.x] += 1u;
                    {
                        const unsigned long long em = (static_cast<unsigned long long>(scnt[4 * blockDim.x]) << 32 | scnt[3 * blockDim.x])
                            + static_cast<unsigned long long>(__float2ll_rn(ss.E * ss.w * 65536.0f));
                        scnt[3 * blockDim.x] = static_cast<unsigned int>(em);
                        scnt[4 * blockDim.x] = static_cast<unsigned int>(em >> 32);
                    }
                    if (hit) {
                        qpos = energyPos(ss.E);
                        qmu = lerp(__ldg(P.tab.majorant + qpos.i), __ldg(P.tab.majorant + qpos.i + 1), qpos.f);
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    spool[0] = poolNext + nb;
                    spool[1] = poolEnd;
                    spool[2] = bufBase;
                }
                const unsigned int mHit = __ballot_sync(0xffffffffu, hit);
                if (hit) {
                    const int k = __popc(mHit & laneLt);
                    sbuf[0 * 32 + k] = q.px;
                    sbuf[1 * 32 + k] = q.py;
                    sbuf[2 * 32 + k] = q.pz;
                    sbuf[3 * 32 + k] = q.dx;
                    sbuf[4 * 32 + k] = q.dy;
                    sbuf[5 * 32 + k] = q.dz;
                    sbuf[6 * 32 + k] = q.E;
                    sbuf[7 * 32 + k] = q.w;
                    sbuf[8 * 32 + k] = __int_as_float(lane);
                    sbuf[9 * 32 + k] = __int_as_float(qpos.i);
                    sbuf[10 * 32 + k] = qpos.f;
                    sbuf[11 * 32 + k] = qmu;
                }
                bufCount = __popc(mHit);
                __syncwarp();
            }
            // dead lanes pop from the top of the buffer
            const unsigned int mDead = __ballot_sync(0xffffffffu, status == kStDead);
            if (status == kStDead) {
                const int r = __popc(mDead & laneLt);
                if (r < bufCount) {
                    const int k = bufCount - 1 - r;
                    ph.px = sbuf[0 * 32 + k];
                    ph.py = sbuf[1 * 32 + k];
                    ph.pz = sbuf[2 * 32 + k];
                    ph.dx = sbuf[3 * 32 + k];
                    ph.dy = sbuf[4 * 32 + k];
                    ph.dz = sbuf[5 * 32 + k];
                    ph.E = sbuf[6 * 32 + k];
                    ph.w = sbuf[7 * 32 + k];
                    const unsigned long long h = spool[2] + static_cast<unsigned int>(__float_as_int(sbuf[8 * 32 + k]));
                    epos.i = __float_as_int(sbuf[9 * 32 + k]);
                    epos.f = sbuf[10 * 32 + k];
                    const float muMax = sbuf[11 * 32 + k];
                    muMaxU24 = muMax * kU24;
                    stepScale = -kLn2 * __fdividef(1.0f, muMax);
                    hlo = static_cast<unsigned int>(h);
                    hhi = static_cast<unsigned int>(h >> 32);
                    blk = 2u; // blocks 0-1 belong to the source
                    status = kStStep;
                }
            }
            bufCount -= min(nDead, bufCount);
            __syncwarp();
        }
    }

    // ---------------- statistics
    unsigned long long v[5] = { nSteps, scnt[0], scnt[blockDim.x], static_cast<unsigned long long>(scnt[4 * blockDim.x]) << 32 | scnt[3 * blockDim.x],
        scnt[2 * blockDim.x] };
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        unsigned long long x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0 && x)
            atomicAdd(P.stats + k, x);
    }
}

// ------------------------------------------------------------------ grid preparation kernels
__global__ void packVoxelsKernel(const double* __restrict__ density, const unsigned char* __restrict__ material,
    unsigned int* __restrict__ out, size_t n, unsigned int* __restrict__ maxDensityBits /* [257]: [256] = largest material index */)
{
    __shared__ unsigned int s_max[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        s_max[i] = 0u;
    __syncthreads();
    unsigned int mmax = 0u;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        float rho = static_cast<float>(density[i]);
        if (!(rho > 0.0f))
            rho = 0.0f;
        const unsigned int m = material[i];
        const unsigned int q = quantizeDensityBits(__float_as_uint(rho));
        out[i] = q | m;
        mmax = max(mmax, m);
        atomicMax(&s_max[m], q); // non-negative floats order like their bit patterns
    }
    for (int o = 16; o > 0; o >>= 1)
        mmax = max(mmax, __shfl_xor_sync(0xffffffffu, mmax, o));
    if ((threadIdx.x & 31) == 0 && mmax)
        atomicMax(&maxDensityBits[256], mmax);
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (s_max[i])
            atomicMax(&maxDensityBits[i], s_max[i]);
}

// slab-local majorants: per slab of 2^shift voxel layers and per material the largest (24-bit) density that occurs.
// Block (slab, part) scans one part of the slab's voxels; maxima meet in shared memory, then in out[slab * 256 + material].
__global__ void slabMaxKernel(const unsigned int* __restrict__ voxels, size_t layerSize, int nz, int shift, int parts,
    unsigned int* __restrict__ out)
{
    __shared__ unsigned int s_max[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        s_max[i] = 0u;
    __syncthreads();
    const int slab = blockIdx.x / parts, part = blockIdx.x % parts;
    const int z0 = slab << shift, z1 = min(nz, (slab + 1) << shift);
    const size_t begin = static_cast<size_t>(z0) * layerSize, end = static_cast<size_t>(z1) * layerSize;
    for (size_t i = begin + static_cast<size_t>(part) * blockDim.x + threadIdx.x; i < end; i += static_cast<size_t>(parts) * blockDim.x) {
        const unsigned int v = voxels[i];
        atomicMax(&s_max[v & 0xFFu], v & 0xFFFFFF00u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (s_max[i])
            atomicMax(&out[slab * 256 + i], s_max[i]);
}

// Dense box (transport_pool.cu, DB builds): the bounding box, in voxel indices, of every voxel that is not thin - a voxel of
// material m is thin when its 24-bit density does not exceed thin[m].  box = {min x, min y, min z, max x, max y, max z}
// (initialised to {nx, ny, nz, -1, -1, -1} by the caller).  Rows along x are scanned by warps: coalesced, and the row's y / z
// are warp-uniform.
__global__ void denseBoxKernel(const unsigned int* __restrict__ voxels, int nx, int ny, int nz, const unsigned int* __restrict__ thin,
    int* __restrict__ box)
{
    __shared__ unsigned int s_thin[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        s_thin[i] = thin[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const size_t warp = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const size_t warps = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 5;
    const size_t rows = static_cast<size_t>(ny) * nz;
    int lo[3] = { nx, ny, nz }, hi[3] = { -1, -1, -1 };
    for (size_t row = warp; row < rows; row += warps) {
        const int y = static_cast<int>(row % ny), z = static_cast<int>(row / ny);
        const unsigned int* r = voxels + row * nx;
        int x0 = nx, x1 = -1;
        for (int x = lane; x < nx; x += 32) {
            const unsigned int v = r[x];
            if ((v & 0xFFFFFF00u) > s_thin[v & 0xFFu]) {
                x0 = min(x0, x);
                x1 = max(x1, x);
            }
        }
        if (x1 >= 0) {
            lo[0] = min(lo[0], x0);
            hi[0] = max(hi[0], x1);
            lo[1] = min(lo[1], y);
            hi[1] = max(hi[1], y);
            lo[2] = min(lo[2], z);
            hi[2] = max(hi[2], z);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = min(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = max(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if (lane == 0 && hi[a] >= 0) {
            atomicMin(box + a, lo[a]);
            atomicMax(box + 3 + a, hi[a]);
        }
    }
}

// ... and per material the largest 24-bit density among the voxels OUTSIDE that box: out[material]
__global__ void outsideMaxKernel(const unsigned int* __restrict__ voxels, int nx, int ny, int nz, const int* __restrict__ box,
    unsigned int* __restrict__ out)
{
    __shared__ unsigned int s_max[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        s_max[i] = 0u;
    __syncthreads();
    const int bx0 = box[0], by0 = box[1], bz0 = box[2], bx1 = box[3], by1 = box[4], bz1 = box[5];
    const int lane = threadIdx.x & 31;
    const size_t warp = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const size_t warps = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 5;
    const size_t rows = static_cast<size_t>(ny) * nz;
    for (size_t row = warp; row < rows; row += warps) {
        const int y = static_cast<int>(row % ny), z = static_cast<int>(row / ny);
        const bool rowInside = y >= by0 && y <= by1 && z >= bz0 && z <= bz1;
        const unsigned int* r = voxels + row * nx;
        for (int x = lane; x < nx; x += 32) {
            if (rowInside && x >= bx0 && x <= bx1)
                continue;
            const unsigned int v = r[x];
            atomicMax(&s_max[v & 0xFFu], v & 0xFFFFFF00u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (s_max[i])
            atomicMax(&out[i], s_max[i]);
}

// Brick pre-filter table (transport_pool.cu, quad step): one block per brick of 2^shift voxels per edge.  The block finds the
// largest (24-bit) density per material inside the brick, then for each of the 8 energy octaves the largest ratio
//   max_m rho_max(m) * tot_m(node) / majorant(node)   over the octave's nodes (both ends included: the kernels interpolate
// numerator and denominator linearly between nodes, and a ratio of two linear functions is monotone in between),
// and stores q with (q + 1) / 256 strictly above it.
__global__ void brickBoundKernel(const unsigned int* __restrict__ voxels, int nx, int ny, int nz, int shift, int nbx, int nby,
    const float* __restrict__ tot, const float* __restrict__ majorant, int n_mat, unsigned char* __restrict__ out)
{
    __shared__ unsigned int s_max[256];
    __shared__ unsigned int s_r[8];
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        s_max[i] = 0u;
    if (threadIdx.x < 8)
        s_r[threadIdx.x] = 0u;
    __syncthreads();
    const int b = blockIdx.x;
    const int bx = b % nbx, by = (b / nbx) % nby, bz = b / (nbx * nby);
    const int edge = 1 << shift, cells = edge * edge * edge;
    for (int t = threadIdx.x; t < cells; t += blockDim.x) {
        const int i = (bx << shift) + (t & (edge - 1)), j = (by << shift) + ((t >> shift) & (edge - 1)), k = (bz << shift) + (t >> (2 * shift));
        if (i < nx && j < ny && k < nz) {
            const unsigned int v = voxels[(static_cast<size_t>(k) * ny + j) * nx + i];
            atomicMax(&s_max[v & 0xFFu], v & 0xFFFFFF00u);
        }
    }
    __syncthreads();
    for (int band = 0; band < 8; ++band) {
        const int node = band * 64 + threadIdx.x;
        if (threadIdx.x <= 64 && node < kDevNE) {
            float mu = 0.0f;
            for (int m = 0; m < n_mat; ++m)
                if (s_max[m])
                    mu = fmaxf(mu, __uint_as_float(s_max[m]) * tot[m * kDevNE + node]);
            const float r = mu / majorant[node];
            atomicMax(&s_r[band], __float_as_uint(fmaxf(r, 0.0f))); // non-negative floats order like their bit patterns
        }
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        const float r = __uint_as_float(s_r[threadIdx.x]);
        const int q = min(255, static_cast<int>(floorf(r * 256.0f * (1.0f + 9.5367431640625e-7f))));
        out[static_cast<size_t>(b) * 8 + threadIdx.x] = static_cast<unsigned char>(max(q, 0));
    }
}

__global__ void majorantKernel(const float* __restrict__ tot, const unsigned int* __restrict__ maxDensityBits,
    int n_mat, float* __restrict__ majorant)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kDevNE)
        return;
    float m = 0.0f;
    for (int k = 0; k < n_mat; ++k)
        m = fmaxf(m, __uint_as_float(maxDensityBits[k]) * tot[k * kDevNE + i]);
    majorant[i] = fmaxf(m, 1e-12f);
}

// per-beam energy tallies -> accumulated dose score (DoseScore::addScoredEnergy, recalled):
//   dose += E k / (rho V);  var += var_E (k/(rho V))^2 with var_E = sum E^2 - (sum E)^2 / n
__global__ void energyToDoseKernel(const unsigned long long* __restrict__ tally, const unsigned int* __restrict__ voxels,
    double* __restrict__ dose, double* __restrict__ variance, unsigned long long* __restrict__ events, size_t n,
    double inv_scale_e, double inv_scale_e2, double factor, double voxel_volume)
{
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const ulonglong4 t = reinterpret_cast<const ulonglong4*>(tally)[i];
        if (t.z == 0)
            continue;
        const double rho = static_cast<double>(voxelDensity(voxels[i]));
        if (!(rho > 0.0))
            continue;
        const double e = static_cast<double>(t.x) * inv_scale_e;
        const double e2 = static_cast<double>(t.y) * inv_scale_e2;
        const double nn = static_cast<double>(t.z);
        const double varE = fmax(0.0, e2 - e * e / nn);
        const double f = factor / (rho * voxel_volume);
        dose[i] += e * f;
        variance[i] += varE * f * f;
        events[i] += t.z;
    }
}

// Multi-process finish, fused with the exchange step: this rank converts voxels [begin, end) and
// reads, per voxel, the SUM over all ranks of the fixed-point tallies
//   MULTICAST  through an NVSwitch multicast address: one multimem.ld_reduce per word, the switch adds the ranks'
//              words in flight (NVLS), so the slab costs one pass at link speed and no intermediate buffer;
//   otherwise  by P2P loads from the peer-mapped tally buffers of the other ranks (NVLink) + the local one.
// Integer sums: identical to the single-GPU tallies whatever the order.
__device__ __forceinline__ void addToDoseScore(unsigned long long se, unsigned long long se2, unsigned long long sn, unsigned int cell,
    double* __restrict__ dose, double* __restrict__ variance, unsigned long long* __restrict__ events, size_t i, double inv_scale_e,
    double inv_scale_e2, double factor, double voxel_volume)
{
    if (sn == 0)
        return;
    const double rho = static_cast<double>(voxelDensity(cell));
    if (!(rho > 0.0))
        return;
    const double e = static_cast<double>(se) * inv_scale_e;
    const double e2 = static_cast<double>(se2) * inv_scale_e2;
    const double nn = static_cast<double>(sn);
    const double varE = fmax(0.0, e2 - e * e / nn);
    const double f = factor / (rho * voxel_volume);
    dose[i] += e * f;
    variance[i] += varE * f * f;
    events[i] += sn;
}

// P2P variant: every thread owns one voxel and adds the 32-byte records of the peers (two 16-byte loads each,
// a warp reads 1 KB contiguous per peer).
__global__ void fusedPullToDoseKernel(const unsigned long long* __restrict__ tally, const unsigned long long* const* __restrict__ peers,
    int n_peers, int first_peer, const unsigned int* __restrict__ voxels, double* __restrict__ dose, double* __restrict__ variance,
    unsigned long long* __restrict__ events, size_t begin, size_t end, double inv_scale_e, double inv_scale_e2, double factor,
    double voxel_volume)
{
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = begin + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < end; i += stride) {
        const ulonglong4 t = reinterpret_cast<const ulonglong4*>(tally)[i];
        unsigned long long se = t.x, se2 = t.y, sn = t.z;
        for (int k = 0; k < n_peers; ++k) {
            // ranks start with different peers so that the links are used evenly
            int p = first_peer + k;
            if (p >= n_peers)
                p -= n_peers;
            const ulonglong2* q = reinterpret_cast<const ulonglong2*>(peers[p] + i * 4);
            const ulonglong2 v0 = __ldcg(q), v1 = __ldcg(q + 1);
            se += v0.x;
            se2 += v0.y;
            sn += v1.x;
        }
        addToDoseScore(se, se2, sn, voxels[i], dose, variance, events, i, inv_scale_e, inv_scale_e2, factor, voxel_volume);
    }
}

// Multicast variant: a warp reads 32 voxels = 128 consecutive 64-bit words of the multicast address with four
// fully coalesced multimem.ld_reduce instructions (lane l reads words l, l+32, l+64, l+96; the switch returns the
// sum over ranks), then shuffles the three words of voxel L to lane L.  (Skipping the pad word of each record
// was measured slower: 4.24 vs 3.82 ms per 1.25 GB slab — the gaps break the 256-byte requests.)
__global__ void fusedMulticastToDoseKernel(const unsigned long long* __restrict__ mc, const unsigned int* __restrict__ voxels,
    double* __restrict__ dose, double* __restrict__ variance, unsigned long long* __restrict__ events, size_t begin, size_t end,
    double inv_scale_e, double inv_scale_e2, double factor, double voxel_volume)
{
    const int lane = threadIdx.x & 31;
    const size_t warpsTotal = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 5;
    const size_t gw = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    for (size_t v0 = begin + gw * 32; v0 < end; v0 += warpsTotal * 32) {
        const unsigned long long* base = mc + v0 * 4;
        unsigned long long r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int word = 32 * j + lane;
            r[j] = 0ull;
            if (v0 + (word >> 2) < end)
                asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u64 %0, [%1];" : "=l"(r[j]) : "l"(base + word) : "memory");
        }
        const int src = (4 * lane) & 31, grp = lane >> 3; // voxel `lane` lives in load `grp`, lanes src .. src+2
        unsigned long long comp[3];
#pragma unroll
        for (int cidx = 0; cidx < 3; ++cidx) {
            const unsigned long long t0 = __shfl_sync(0xffffffffu, r[0], src + cidx), t1 = __shfl_sync(0xffffffffu, r[1], src + cidx);
            const unsigned long long t2 = __shfl_sync(0xffffffffu, r[2], src + cidx), t3 = __shfl_sync(0xffffffffu, r[3], src + cidx);
            comp[cidx] = grp == 0 ? t0 : (grp == 1 ? t1 : (grp == 2 ? t2 : t3));
        }
        const size_t i = v0 + lane;
        if (i < end)
            addToDoseScore(comp[0], comp[1], comp[2], voxels[i], dose, variance, events, i, inv_scale_e, inv_scale_e2, factor, voxel_volume);
    }
}

__global__ void tallyToEnergyKernel(const unsigned long long* __restrict__ tally, double* __restrict__ e,
    double* __restrict__ e2, unsigned long long* __restrict__ cnt, size_t n, double inv_scale_e, double inv_scale_e2)
{
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const ulonglong4 t = reinterpret_cast<const ulonglong4*>(tally)[i];
        if (e)
            e[i] = static_cast<double>(t.x) * inv_scale_e;
        if (e2)
            e2[i] = static_cast<double>(t.y) * inv_scale_e2;
        if (cnt)
            cnt[i] = t.z;
    }
}

// in-process multi-GPU: device 0 pulls the peers' tallies over NVLink peer memory and adds them
__global__ void peerReduceKernel(unsigned long long* __restrict__ dst, const unsigned long long* const* __restrict__ peers,
    int n_peers, size_t n_words)
{
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    const size_t n4 = n_words / 4;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
        ulonglong4 acc = reinterpret_cast<ulonglong4*>(dst)[i];
        for (int p = 0; p < n_peers; ++p) {
            const ulonglong4 v = reinterpret_cast<const ulonglong4*>(peers[p])[i];
            acc.x += v.x;
            acc.y += v.y;
            acc.z += v.z;
            acc.w += v.w;
        }
        reinterpret_cast<ulonglong4*>(dst)[i] = acc;
    }
}

// CT calibration: integer sums of the kerma tally (word 0) over the voxels of each of the five CTDI measurement holes
__global__ void holeSumKernel(const unsigned long long* __restrict__ tally, const signed char* __restrict__ hole, size_t n,
    unsigned long long* __restrict__ sums /*[5]*/)
{
    __shared__ unsigned long long s_sum[5];
    if (threadIdx.x < 5)
        s_sum[threadIdx.x] = 0ull;
    __syncthreads();
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int h = hole[i];
        if (h >= 0) {
            const unsigned long long v = tally[i * 4];
            if (v)
                atomicAdd(&s_sum[h], v);
        }
    }
    __syncthreads();
    if (threadIdx.x < 5 && s_sum[threadIdx.x])
        atomicAdd(&sums[threadIdx.x], s_sum[threadIdx.x]);
}

// reference post-processing (R:src/libopendxmc/simulationpipeline.cpp:180-185,206-211,221-229)
__global__ void postprocessKernel(const double* __restrict__ in, const unsigned int* __restrict__ voxels, double* __restrict__ out,
    size_t n, int maskAir, double scale)
{
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        double v = in[i];
        if (maskAir && voxelMaterial(voxels[i]) == 0)
            v = 0.0;
        out[i] = v * scale;
    }
}

__global__ void u64ToDoubleKernel(const unsigned long long* __restrict__ in, const unsigned int* __restrict__ voxels,
    double* __restrict__ out, size_t n, int maskAir)
{
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        double v = static_cast<double>(in[i]);
        if (maskAir && voxelMaterial(voxels[i]) == 0)
            v = 0.0;
        out[i] = v;
    }
}

// block max reduction for the uGy decision (max(dose) < 1)
__global__ void maxKernel(const double* __restrict__ in, const unsigned int* __restrict__ voxels, size_t n, int maskAir,
    unsigned long long* __restrict__ outBits)
{
    double m = 0.0;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        double v = in[i];
        if (maskAir && voxelMaterial(voxels[i]) == 0)
            v = 0.0;
        m = fmax(m, v);
    }
    for (int o = 16; o > 0; o >>= 1)
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0)
        atomicMax(outBits, static_cast<unsigned long long>(__double_as_longlong(m))); // non-negative doubles
}

// per-organ mass-weighted dose (R:src/libopendxmc/dosetablepipeline.cpp:60-84):
//   dose_o = sum(dose rho V) / sum(rho V) over voxels of organ o
__global__ void organDoseKernel(const double* __restrict__ dose, const double* __restrict__ variance,
    const unsigned int* __restrict__ voxels, const unsigned char* __restrict__ organ, size_t n, double voxel_volume,
    double* __restrict__ energy /*[256]*/, double* __restrict__ mass /*[256]*/, unsigned long long* __restrict__ count /*[256]*/,
    double* __restrict__ varEnergy /*[256]*/)
{
    __shared__ double s_e[256], s_m[256], s_v[256];
    __shared__ unsigned long long s_c[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        s_e[i] = 0.0;
        s_m[i] = 0.0;
        s_v[i] = 0.0;
        s_c[i] = 0ull;
    }
    __syncthreads();
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned int o = organ[i];
        const double m = static_cast<double>(voxelDensity(voxels[i])) * voxel_volume;
        atomicAdd(&s_e[o], dose[i] * m);
        atomicAdd(&s_m[o], m);
        atomicAdd(&s_v[o], variance[i] * m * m);
        atomicAdd(&s_c[o], 1ull);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        if (s_c[i]) {
            atomicAdd(&energy[i], s_e[i]);
            atomicAdd(&mass[i], s_m[i]);
            atomicAdd(&varEnergy[i], s_v[i]);
            atomicAdd(&count[i], s_c[i]);
        }
    }
}

// device-side table lookups with the kernel's own float code (parity test "lookups within 1e-6")
__global__ void attenuationProbeKernel(TablesDev tab, int mat, const float* __restrict__ energy, int n, float* __restrict__ out4)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const TabPos p = energyPos(energy[i]);
    const float4 a = tab.att[mat * kDevNE + p.i];
    const float4 b = tab.att[mat * kDevNE + p.i + 1];
    out4[4 * i + 0] = lerp(a.x, b.x, p.f);
    out4[4 * i + 1] = lerp(a.y, b.y, p.f);
    out4[4 * i + 2] = lerp(a.z, b.z, p.f);
    out4[4 * i + 3] = lerp(tab.tot[mat * kDevNE + p.i], tab.tot[mat * kDevNE + p.i + 1], p.f);
}

__global__ void majorantProbeKernel(const float* __restrict__ majorant, const float* __restrict__ energy, int n, float* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const TabPos p = energyPos(energy[i]);
    out[i] = lerp(majorant[p.i], majorant[p.i + 1], p.f);
}

// CT segmentation, R:src/libopendxmc/ctsegmentationpipeline.cpp:136-156:
// material = first i with HU < sep[i] (else the last material); density from the HU/attenuation relation
__global__ void segmentKernel(const double* __restrict__ hu, size_t n, const double* __restrict__ sep, int n_sep,
    const double* __restrict__ matAtt /*[n_sep+1] spectrum-weighted mass attenuation*/, double waterAttDens, double airAttDens,
    unsigned char* __restrict__ material, double* __restrict__ density)
{
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double h = hu[i];
        int m = n_sep;
        for (int t = n_sep - 1; t >= 0; --t)
            if (h < sep[t])
                m = t;
        material[i] = static_cast<unsigned char>(m);
        const double dens = ((waterAttDens - airAttDens) * h / 1000.0 + waterAttDens) / matAtt[m];
        density[i] = fmax(dens, 0.0);
    }
}

} // namespace

// ---------------------------------------------------------------------- host-callable launchers
template <int MODE, bool CALIB, bool SMEM>
static cudaError_t launchT(const RunParams& p, const LaunchConfig& cfg, cudaStream_t stream)
{
    auto kern = transportKernel<MODE, CALIB, SMEM>;
    if (cfg.smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(cfg.smem));
        if (e != cudaSuccess)
            return e;
    }
    kern<<<cfg.blocks, cfg.threads, cfg.smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launchTransport(const RunParams& p, int mode, bool calib, const LaunchConfig& cfg, cudaStream_t stream)
{
#define DXB_DISPATCH(M, C)                                     \
    if (cfg.table_in_smem)                                     \
        return launchT<M, C, true>(p, cfg, stream);            \
    else                                                       \
        return launchT<M, C, false>(p, cfg, stream);
    if (!calib) {
        if (mode == 0) {
            DXB_DISPATCH(0, false)
        } else if (mode == 1) {
            DXB_DISPATCH(1, false)
        } else {
            DXB_DISPATCH(2, false)
        }
    } else {
        if (mode == 0) {
            DXB_DISPATCH(0, true)
        } else if (mode == 1) {
            DXB_DISPATCH(1, true)
        } else {
            DXB_DISPATCH(2, true)
        }
    }
#undef DXB_DISPATCH
}

int transportOccupancy(int mode, bool calib, bool smemTable, int threads, size_t smem)
{
    int nb = 0;
    cudaError_t e;
#define DXB_OCC(M, C, S)                                                                                            \
    {                                                                                                               \
        auto k = transportKernel<M, C, S>;                                                                          \
        if (smem > 48 * 1024)                                                                                       \
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));           \
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, threads, smem);                                   \
    }
    if (!calib) {
        if (mode == 0) {
            if (smemTable) DXB_OCC(0, false, true) else DXB_OCC(0, false, false)
        } else if (mode == 1) {
            if (smemTable) DXB_OCC(1, false, true) else DXB_OCC(1, false, false)
        } else {
            if (smemTable) DXB_OCC(2, false, true) else DXB_OCC(2, false, false)
        }
    } else {
        if (mode == 0) {
            if (smemTable) DXB_OCC(0, true, true) else DXB_OCC(0, true, false)
        } else if (mode == 1) {
            if (smemTable) DXB_OCC(1, true, true) else DXB_OCC(1, true, false)
        } else {
            if (smemTable) DXB_OCC(2, true, true) else DXB_OCC(2, true, false)
        }
    }
#undef DXB_OCC
    return e == cudaSuccess ? nb : 0;
}

// grid size of the HBM-streaming kernels: 8 blocks of 256 threads per SM of the device the context runs on
static int g_streamBlocks = 148 * 8;
void setLaunchSmCount(int sms) { g_streamBlocks = (sms > 0 ? sms : 148) * 8; }

void launchPackVoxels(const double* density, const unsigned char* material, unsigned int* out, size_t n, unsigned int* maxBits, cudaStream_t s)
{
    packVoxelsKernel<<<g_streamBlocks, 256, 0, s>>>(density, material, out, n, maxBits);
}
void launchSlabMax(const unsigned int* voxels, size_t layerSize, int nz, int shift, int nslabs, unsigned int* out, cudaStream_t s)
{
    const int parts = max(1, g_streamBlocks / max(nslabs, 1));
    slabMaxKernel<<<nslabs * parts, 256, 0, s>>>(voxels, layerSize, nz, shift, parts, out);
}
void launchDenseBox(const unsigned int* voxels, int nx, int ny, int nz, const unsigned int* thin, int* box, cudaStream_t s)
{
    denseBoxKernel<<<g_streamBlocks, 256, 0, s>>>(voxels, nx, ny, nz, thin, box);
}
void launchOutsideMax(const unsigned int* voxels, int nx, int ny, int nz, const int* box, unsigned int* out, cudaStream_t s)
{
    outsideMaxKernel<<<g_streamBlocks, 256, 0, s>>>(voxels, nx, ny, nz, box, out);
}
void launchBrickBound(const unsigned int* voxels, int nx, int ny, int nz, int shift, int nbx, int nby, int nbz, const float* tot,
    const float* majorant, int n_mat, unsigned char* out, cudaStream_t s)
{
    brickBoundKernel<<<nbx * nby * nbz, 128, 0, s>>>(voxels, nx, ny, nz, shift, nbx, nby, tot, majorant, n_mat, out);
}
void launchMajorant(const float* tot, const unsigned int* maxBits, int n_mat, float* majorant, cudaStream_t s)
{
    majorantKernel<<<(kDevNE + 127) / 128, 128, 0, s>>>(tot, maxBits, n_mat, majorant);
}
void launchEnergyToDose(const unsigned long long* tally, const unsigned int* voxels, double* dose, double* variance,
    unsigned long long* events, size_t n, double inv_e, double inv_e2, double factor, double vol, cudaStream_t s)
{
    energyToDoseKernel<<<g_streamBlocks, 256, 0, s>>>(tally, voxels, dose, variance, events, n, inv_e, inv_e2, factor, vol);
}
void launchFusedReduceToDose(const unsigned long long* tally, bool multicast, const unsigned long long* const* peers, int n_peers,
    int first_peer, const unsigned int* voxels, double* dose, double* variance, unsigned long long* events, size_t begin, size_t end,
    double inv_e, double inv_e2, double factor, double vol, cudaStream_t s)
{
    if (multicast)
        fusedMulticastToDoseKernel<<<g_streamBlocks, 256, 0, s>>>(tally, voxels, dose, variance, events, begin, end, inv_e, inv_e2, factor, vol);
    else
        fusedPullToDoseKernel<<<g_streamBlocks, 256, 0, s>>>(tally, peers, n_peers, first_peer, voxels, dose, variance, events, begin, end, inv_e,
            inv_e2, factor, vol);
}
void launchTallyToEnergy(const unsigned long long* tally, double* e, double* e2, unsigned long long* cnt, size_t n,
    double inv_e, double inv_e2, cudaStream_t s)
{
    tallyToEnergyKernel<<<g_streamBlocks, 256, 0, s>>>(tally, e, e2, cnt, n, inv_e, inv_e2);
}
void launchPeerReduce(unsigned long long* dst, const unsigned long long* const* peers, int n_peers, size_t n_words, cudaStream_t s)
{
    peerReduceKernel<<<g_streamBlocks, 256, 0, s>>>(dst, peers, n_peers, n_words);
}
void launchHoleSums(const unsigned long long* tally, const signed char* hole, size_t n, unsigned long long* sums, cudaStream_t s)
{
    holeSumKernel<<<g_streamBlocks / 4, 256, 0, s>>>(tally, hole, n, sums);
}
void launchPostprocess(const double* in, const unsigned int* voxels, double* out, size_t n, int maskAir, double scale, cudaStream_t s)
{
    postprocessKernel<<<g_streamBlocks, 256, 0, s>>>(in, voxels, out, n, maskAir, scale);
}
void launchU64ToDouble(const unsigned long long* in, const unsigned int* voxels, double* out, size_t n, int maskAir, cudaStream_t s)
{
    u64ToDoubleKernel<<<g_streamBlocks, 256, 0, s>>>(in, voxels, out, n, maskAir);
}
void launchMax(const double* in, const unsigned int* voxels, size_t n, int maskAir, unsigned long long* outBits, cudaStream_t s)
{
    maxKernel<<<g_streamBlocks / 2, 256, 0, s>>>(in, voxels, n, maskAir, outBits);
}
void launchOrganDose(const double* dose, const double* variance, const unsigned int* voxels, const unsigned char* organ, size_t n,
    double vol, double* energy, double* mass, unsigned long long* count, double* varEnergy, cudaStream_t s)
{
    organDoseKernel<<<g_streamBlocks / 2, 256, 0, s>>>(dose, variance, voxels, organ, n, vol, energy, mass, count, varEnergy);
}
void launchAttenuationProbe(const TablesDev& tab, int mat, const float* energy, int n, float* out4, cudaStream_t s)
{
    attenuationProbeKernel<<<(n + 127) / 128, 128, 0, s>>>(tab, mat, energy, n, out4);
}
void launchMajorantProbe(const float* majorant, const float* energy, int n, float* out, cudaStream_t s)
{
    majorantProbeKernel<<<(n + 127) / 128, 128, 0, s>>>(majorant, energy, n, out);
}
void launchSegment(const double* hu, size_t n, const double* sep, int n_sep, const double* matAtt, double waterAttDens,
    double airAttDens, unsigned char* material, double* density, cudaStream_t s)
{
    segmentKernel<<<g_streamBlocks, 256, 0, s>>>(hu, n, sep, n_sep, matAtt, waterAttDens, airAttDens, material, density);
}

} // namespace dxb
