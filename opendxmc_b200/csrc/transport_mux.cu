// transport_mux.cu — the lane-multiplexed photon-history kernel for sm_100a (the default transport kernel).
//
// Same physics, same random-number protocol (header of transport.cu) and therefore the same results, history
// for history, as the one-photon-per-lane kernel in transport.cu; what changes is where photons live.
//
// Every lane owns M photon *slots* in shared memory (10 words each, lane-private columns: word f of slot j of
// lane l sits at [(f*M + j)*32 + l], so every access of a warp is bank-conflict free) and a status word with
// one bit per slot and state.  Each iteration the warp votes for ONE phase (step / interaction try / Rayleigh
// try / refill) and every lane picks, among ITS OWN slots, a photon that is in that phase.  With M = 4 a lane
// almost always has one, so phases run with ~28-32 active lanes instead of the 18 of the register kernel
// (profiles/r01_transport_v5_summary.txt), at the price of loading / storing the photon state around each phase.
// Nothing persistent lives in registers, so occupancy is set by shared memory, not by the register file.
//
// Replaces (recalled DXMClib): Transport::runWorker -> exposure.sampleParticle
// -> World::transport -> AAVoxelGrid::woodcockTransport -> interactions::interact -> EnergyScore.
#include "transport_common.cuh"

namespace dxb {

namespace {

constexpr unsigned int kMetaBlkMask = 0xFFFFFu; // Philox block index of the history (20 bits)
constexpr int kMetaMatShift = 20;               // material of the pending interaction (8 bits)
constexpr unsigned int kMetaRetry = 1u << 28;   // Compton already chosen, previous candidate rejected

// slot words
enum : int { kWPx = 0, kWPy, kWPz, kWDx, kWDy, kWDz, kWE, kWW, kWHlo, kWMeta };
static_assert(kWMeta + 1 == kSlotWords, "slot layout");
// byte of the status word that holds the slot mask of a state
enum : int { kPhStep = 0, kPhInt = 1, kPhRay = 2, kPhDead = 3, kPhNone = 4 };

template <int M>
struct MuxBounds {
    static constexpr int kMinBlocks = M <= 3 ? 4 : (M <= 4 ? 3 : 2);
};

template <int MODE, bool CALIB, bool SMEM_TABLE, int M>
__global__ void __launch_bounds__(256, MuxBounds<M>::kMinBlocks) transportKernelMux(const __grid_constant__ RunParams P)
{
    static_assert(M >= 1 && M <= 7, "status bytes hold at most 7 slot bits");
    extern __shared__ __align__(16) unsigned char s_raw[];
    constexpr unsigned int kFull = 0xffffffffu;
    constexpr int kStride = M * 32; // distance between two words of one slot
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nWarps = blockDim.x >> 5;
    // layout: muxSmemBytes (device_types.cuh)
    float* __restrict__ s_f = reinterpret_cast<float*>(s_raw);
    const int nTab = SMEM_TABLE ? P.tab.n_mat * kDevNE : 0;
    float* __restrict__ s_tot = s_f;
    float* __restrict__ s_maj = s_f + nTab;
    float* __restrict__ slots = s_maj + kDevNE + warp * (kSlotWords * kStride) + lane;
    float* __restrict__ sbuf = s_maj + kDevNE + nWarps * (kSlotWords * kStride) + warp * (kMuxBufWords * 32);
    for (int i = threadIdx.x; i < nTab; i += blockDim.x)
        s_tot[i] = P.tab.tot[i];
    for (int i = threadIdx.x; i < kDevNE; i += blockDim.x)
        s_maj[i] = P.tab.majorant[i];
    __syncthreads();
    const float* __restrict__ totTable = SMEM_TABLE ? s_tot : P.tab.tot;
    const GridDev& G = P.grid;

    // warp-uniform bookkeeping
    unsigned long long poolNext = 0, poolEnd = 0, bufBase = 0;
    bool drained = false;
    int bufCount = 0;
    // per-lane
    unsigned int st = ((1u << M) - 1u) << (8 * kPhDead); // every slot dead
    unsigned int nSteps = 0, nInteractions = 0, nDeposits = 0, nHistories = 0;
    unsigned long long emitted = 0;

    for (;;) {
        // lanes that own at least one slot in each state: bit 7 of a status byte (< 0x80) + 0x7f is set iff it is non-zero
        const unsigned int votes = __reduce_add_sync(kFull, ((st + 0x7f7f7f7fu) >> 7) & 0x01010101u);
        const int nStep = votes & 0xff, nInt = (votes >> 8) & 0xff, nRay = (votes >> 16) & 0xff, nDead = votes >> 24;
        const bool canRefill = !(drained && bufCount == 0);
        int phase;
        if (canRefill && nDead >= P.refill_threshold)
            phase = kPhDead;
        else if (nRay >= P.rayleigh_threshold)
            phase = kPhRay;
        else if (nInt > 0 && (nInt + P.interact_bias >= nStep || nInt >= P.interact_threshold))
            phase = kPhInt;
        else if (nStep > 0)
            phase = kPhStep;
        else if (nInt > 0)
            phase = kPhInt;
        else if (nRay > 0)
            phase = kPhRay;
        else if (canRefill && nDead > 0)
            phase = kPhDead;
        else
            break;

        if (phase == kPhStep) {
            // ------------------------------------------------------------ pairs of tentative Woodcock steps
            const unsigned int m = st & 0xffu;
            const bool active = m != 0u;
            const int j = active ? __ffs(m) - 1 : 0;
            float* __restrict__ sp = slots + j * 32;
            float px = 0.f, py = 0.f, pz = 0.f, dx = 0.f, dy = 0.f, dz = 0.f, E = 0.f, w = 0.f;
            unsigned int hlo = 0, hhi = 0, blk = 0;
            TabPos epos;
            epos.i = 0;
            epos.f = 0.f;
            float muMaxU24 = kU24, stepScale = -kLn2;
            if (active) {
                px = sp[kWPx * kStride];
                py = sp[kWPy * kStride];
                pz = sp[kWPz * kStride];
                dx = sp[kWDx * kStride];
                dy = sp[kWDy * kStride];
                dz = sp[kWDz * kStride];
                E = sp[kWE * kStride];
                if (CALIB)
                    w = sp[kWW * kStride];
                hlo = __float_as_uint(sp[kWHlo * kStride]);
                blk = __float_as_uint(sp[kWMeta * kStride]) & kMetaBlkMask;
                hhi = P.hbase_hi + (hlo < P.hbase_lo ? 1u : 0u);
                epos = energyPos(E);
                const float muMax = lerp(s_maj[epos.i], s_maj[epos.i + 1], epos.f);
                muMaxU24 = muMax * kU24;
                stepScale = -kLn2 * __fdividef(1.0f, muMax);
            }
            bool stepping = active;
            int newPhase = kPhStep;
            int mat = 0;
            for (int it = 0; it < P.step_pairs; ++it) {
                float kermaA = 0.0f, kermaB = 0.0f;
                unsigned int voxA = 0, voxB = 0;
                if (stepping) {
                    const PhiloxBlock rb = philox4x32_10(P.round_key, hlo, hhi, blk++);
                    const float sA = __log2f(fmaf(rb.k(0), -kU24, 1.0f)) * stepScale;
                    const float sB = __log2f(fmaf(rb.k(2), -kU24, 1.0f)) * stepScale;
                    const float ax = fmaf(dx, sA, px), ay = fmaf(dy, sA, py), az = fmaf(dz, sA, pz);
                    const float bx = fmaf(dx, sB, ax), by = fmaf(dy, sB, ay), bz = fmaf(dz, sB, az);
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
                        newPhase = kPhDead; // left the grid
                        stepping = false;
                    } else {
                        ++nSteps;
                        const int matA = voxelMaterial(cellA);
                        const float* tt = totTable + matA * kDevNE + epos.i;
                        const float muA = voxelDensity(cellA) * lerp(tt[0], tt[1], epos.f);
                        if (CALIB && matA == P.score_material) {
                            // collision estimator of air kerma: every tentative collision carries 1/mu_max of track length
                            const float* et = P.tab.etr + matA * kDevNE + epos.i;
                            kermaA = w * E * lerp(__ldg(et), __ldg(et + 1), epos.f) * (stepScale * -kInvLn2);
                        }
                        if (rb.k(1) * muMaxU24 < muA) {
                            px = ax;
                            py = ay;
                            pz = az;
                            mat = matA;
                            newPhase = kPhInt;
                            stepping = false;
                        } else if (!inB) {
                            newPhase = kPhDead;
                            stepping = false;
                        } else {
                            ++nSteps;
                            const int matB = voxelMaterial(cellB);
                            const float* tb = totTable + matB * kDevNE + epos.i;
                            const float muB = voxelDensity(cellB) * lerp(tb[0], tb[1], epos.f);
                            if (CALIB && matB == P.score_material) {
                                const float* et = P.tab.etr + matB * kDevNE + epos.i;
                                kermaB = w * E * lerp(__ldg(et), __ldg(et + 1), epos.f) * (stepScale * -kInvLn2);
                            }
                            px = bx;
                            py = by;
                            pz = bz;
                            if (rb.k(3) * muMaxU24 < muB) {
                                mat = matB;
                                newPhase = kPhInt;
                                stepping = false;
                            }
                        }
                    }
                }
                if (CALIB) {
                    unsigned int mScore = __ballot_sync(kFull, kermaA > 0.0f);
                    if (kermaA > 0.0f)
                        scoreEnergy(mScore, G.tally, voxA, kermaA, P.tally_scale_e, P.tally_scale_e2);
                    mScore = __ballot_sync(kFull, kermaB > 0.0f);
                    if (kermaB > 0.0f)
                        scoreEnergy(mScore, G.tally, voxB, kermaB, P.tally_scale_e, P.tally_scale_e2);
                }
                if (!__any_sync(kFull, stepping))
                    break;
            }
            if (active) {
                if (newPhase != kPhDead) {
                    sp[kWPx * kStride] = px;
                    sp[kWPy * kStride] = py;
                    sp[kWPz * kStride] = pz;
                    sp[kWMeta * kStride] = __uint_as_float(blk | (static_cast<unsigned int>(mat) << kMetaMatShift));
                }
                st ^= (1u << j) ^ (1u << (8 * newPhase + j));
            }
        } else if (phase == kPhInt || phase == kPhRay) {
            // ------------------------------------------------------------ one sampling try per waiting lane
            const unsigned int m = (st >> (8 * phase)) & 0xffu;
            const bool active = m != 0u;
            const int j = active ? __ffs(m) - 1 : 0;
            float* __restrict__ sp = slots + j * 32;
            float edep = 0.0f;
            unsigned int voxel = 0;
            if (active) {
                float px = sp[kWPx * kStride], py = sp[kWPy * kStride], pz = sp[kWPz * kStride];
                float dx = sp[kWDx * kStride], dy = sp[kWDy * kStride], dz = sp[kWDz * kStride];
                float E = sp[kWE * kStride], w = sp[kWW * kStride];
                const unsigned int hlo = __float_as_uint(sp[kWHlo * kStride]);
                unsigned int meta = __float_as_uint(sp[kWMeta * kStride]);
                const unsigned int hhi = P.hbase_hi + (hlo < P.hbase_lo ? 1u : 0u);
                unsigned int blk = meta & kMetaBlkMask;
                const int mat = static_cast<int>((meta >> kMetaMatShift) & 0xffu);
                int newPhase = phase;
                bool scattered = false; // an accepted scatter: cut-off, roulette, exit distance
                const PhiloxBlock rb = philox4x32_10(P.round_key, hlo, hhi, blk++);
                if (phase == kPhInt) {
                    bool compton = (meta & kMetaRetry) != 0u;
                    if (!compton) {
                        ++nInteractions;
                        const TabPos epos = energyPos(E);
                        const float4 a = __ldg(P.tab.att + mat * kDevNE + epos.i);
                        const float4 b = __ldg(P.tab.att + mat * kDevNE + epos.i + 1);
                        const float aPhoto = lerp(a.x, b.x, epos.f);
                        const float aIncoh = lerp(a.y, b.y, epos.f);
                        const float aTot = lerp(a.w, b.w, epos.f);
                        const float r2 = rb.u(0) * aTot;
                        if (r2 < aPhoto) {
                            const float ef = MODE >= 2 ? photoFluorescence(P.tab, mat, E, rb.u(1), rb.u(2)) : 0.0f;
                            if (MODE >= 2 && ef > 0.0f) {
                                // fluorescence photon: isotropic, one extra block for its direction
                                const PhiloxBlock rf = philox4x32_10(P.round_key, hlo, hhi, blk++);
                                isotropic(rf.u(0), rf.u(1), dx, dy, dz);
                                edep = (E - ef) * w;
                                E = ef;
                                scattered = true;
                            } else {
                                edep = E * w;
                                E = 0.0f;
                                newPhase = kPhDead;
                            }
                        } else if (r2 < aPhoto + aIncoh) {
                            compton = true;
                        } else {
                            newPhase = kPhRay;
                        }
                    }
                    if (compton) {
                        float e, cosT;
                        bool ok = comptonTry<MODE>(P.tab, mat, E, rb.u(1), rb.u(2), e, cosT);
                        if (MODE >= 2 && ok) {
                            // impulse approximation: shell + Doppler broadening from one extra block
                            const PhiloxBlock ri = philox4x32_10(P.round_key, hlo, hhi, blk++);
                            ok = dopplerBroaden(P.tab, mat, E, e, cosT, ri.u(0), ri.u(1), e);
                        }
                        if (ok) {
                            deflect(dx, dy, dz, cosT, kTwoPi * rb.u(3));
                            const float E0 = E;
                            E = E0 * e;
                            edep = (E0 - E) * w;
                            scattered = true;
                        } else {
                            meta |= kMetaRetry;
                        }
                    }
                } else {
                    float cosT;
                    if (rayleighTry<MODE>(P.tab, mat, E, rb.u(0), rb.u(1), cosT)) {
                        deflect(dx, dy, dz, cosT, kTwoPi * rb.u(2));
                        scattered = true;
                    }
                }
                if (scattered) {
                    newPhase = kPhStep;
                    if (E < kMinEnergy) {
                        edep += E * w;
                        E = 0.0f;
                        newPhase = kPhDead;
                    } else if (w < kRouletteThreshold) {
                        const PhiloxBlock rr = philox4x32_10(P.round_key, hlo, hhi, blk++);
                        if (rr.u(0) < kRouletteKill)
                            newPhase = kPhDead;
                        else
                            w *= 1.0f / (1.0f - kRouletteKill);
                    }
                    if (newPhase == kPhStep) {
                        sp[kWDx * kStride] = dx;
                        sp[kWDy * kStride] = dy;
                        sp[kWDz * kStride] = dz;
                        sp[kWE * kStride] = E;
                        sp[kWW * kStride] = w;
                        meta &= ~kMetaRetry;
                    }
                }
                if (newPhase != kPhDead)
                    sp[kWMeta * kStride] = __uint_as_float((meta & ~kMetaBlkMask) | blk);
                if (CALIB)
                    edep = 0.0f;
                if (edep > 0.0f)
                    voxelIndex(G, px, py, pz, voxel);
                st ^= (1u << (8 * phase + j)) ^ (1u << (8 * newPhase + j));
            }
            if (!CALIB && phase == kPhInt) {
                const unsigned int mScore = __ballot_sync(kFull, edep > 0.0f);
                if (edep > 0.0f) {
                    ++nDeposits;
                    scoreEnergy(mScore, G.tally, voxel, edep, P.tally_scale_e, P.tally_scale_e2);
                }
            }
        } else {
            // ------------------------------------------------------------ refill
            const unsigned int laneLt = (1u << lane) - 1u;
            if (bufCount == 0) {
                // warp-cooperative source sampling of the next (up to) 32 histories into the warp's buffer
                if (poolNext == poolEnd && !drained) {
                    constexpr unsigned long long kPiece = 256;
                    unsigned long long base = 0;
                    if (lane == 0)
                        base = atomicAdd(P.work_counter, kPiece);
                    base = __shfl_sync(kFull, base, 0);
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
                bufBase = (sblk * P.world + P.rank) * kShardBlock + (poolNext % kShardBlock);
                poolNext += nb;
                const unsigned long long h = bufBase + lane;
                bool hit = false;
                float qpx = 0.f, qpy = 0.f, qpz = 0.f, qdx = 0.f, qdy = 0.f, qdz = 0.f, qE = 0.f, qw = 0.f;
                if (lane < nb && h < P.n_total) {
                    SourceSample q;
                    hit = sampleSource(P, h, q);
                    qpx = q.px, qpy = q.py, qpz = q.pz;
                    qdx = q.dx, qdy = q.dy, qdz = q.dz;
                    qE = q.E;
                    qw = q.w;
                    ++nHistories;
                    emitted += static_cast<unsigned long long>(__float2ll_rn(q.E * q.w * 65536.0f));
                }
                const unsigned int mHit = __ballot_sync(kFull, hit);
                if (hit) {
                    const int k = __popc(mHit & laneLt);
                    sbuf[0 * 32 + k] = qpx;
                    sbuf[1 * 32 + k] = qpy;
                    sbuf[2 * 32 + k] = qpz;
                    sbuf[3 * 32 + k] = qdx;
                    sbuf[4 * 32 + k] = qdy;
                    sbuf[5 * 32 + k] = qdz;
                    sbuf[6 * 32 + k] = qE;
                    sbuf[7 * 32 + k] = qw;
                    sbuf[8 * 32 + k] = __int_as_float(lane);
                }
                bufCount = __popc(mHit);
                __syncwarp();
            }
            // lanes with a dead slot pop from the top of the buffer
            const unsigned int md = st >> (8 * kPhDead);
            const bool wants = md != 0u;
            const unsigned int mWant = __ballot_sync(kFull, wants);
            if (wants) {
                const int r = __popc(mWant & laneLt);
                if (r < bufCount) {
                    const int k = bufCount - 1 - r;
                    const int j = __ffs(md) - 1;
                    float* __restrict__ sp = slots + j * 32;
                    sp[kWPx * kStride] = sbuf[0 * 32 + k];
                    sp[kWPy * kStride] = sbuf[1 * 32 + k];
                    sp[kWPz * kStride] = sbuf[2 * 32 + k];
                    sp[kWDx * kStride] = sbuf[3 * 32 + k];
                    sp[kWDy * kStride] = sbuf[4 * 32 + k];
                    sp[kWDz * kStride] = sbuf[5 * 32 + k];
                    sp[kWE * kStride] = sbuf[6 * 32 + k];
                    sp[kWW * kStride] = sbuf[7 * 32 + k];
                    const unsigned long long h = bufBase + static_cast<unsigned int>(__float_as_int(sbuf[8 * 32 + k]));
                    sp[kWHlo * kStride] = __uint_as_float(static_cast<unsigned int>(h));
                    sp[kWMeta * kStride] = __uint_as_float(2u); // blocks 0-1 belong to the source
                    st ^= (1u << (8 * kPhDead + j)) ^ (1u << j);
                }
            }
            bufCount -= min(__popc(mWant), bufCount);
            __syncwarp();
        }
    }

    // ---------------- statistics
    unsigned long long v[5] = { nSteps, nInteractions, nDeposits, emitted, nHistories };
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        unsigned long long x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            x += __shfl_xor_sync(kFull, x, o);
        if (lane == 0 && x)
            atomicAdd(P.stats + k, x);
    }
}

template <int MODE, bool CALIB, bool SMEM, int M>
cudaError_t launchMux(const RunParams& p, const LaunchConfig& cfg, cudaStream_t stream)
{
    auto kern = transportKernelMux<MODE, CALIB, SMEM, M>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(cfg.smem));
    if (e != cudaSuccess)
        return e;
    kern<<<cfg.blocks, cfg.threads, cfg.smem, stream>>>(p);
    return cudaGetLastError();
}

template <int MODE, bool CALIB, bool SMEM, int M>
int occupancyMux(int threads, size_t smem)
{
    auto kern = transportKernelMux<MODE, CALIB, SMEM, M>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return nb;
}

// applies CALL(MODE, CALIB, SMEM, M) for the run-time (mode, calib, smem, slots).  The slot count is a tuning
// knob of the production variant only (mode 1, scoring); every other variant is built with 4 slots per lane.
#define DXB_MUX_DISPATCH(CALL)                                                  \
    const int md = mode <= 0 ? 0 : (mode == 1 ? 1 : 2);                         \
    const int key = (md << 2) | (calib ? 2 : 0) | (smemTable ? 1 : 0);          \
    switch (key) {                                                              \
    case 8: return CALL(2, false, false, 4);                                    \
    case 9: return CALL(2, false, true, 4);                                     \
    case 10: return CALL(2, true, false, 4);                                    \
    case 11: return CALL(2, true, true, 4);                                     \
    case 0: return CALL(0, false, false, 4);                                    \
    case 1: return CALL(0, false, true, 4);                                     \
    case 2: return CALL(0, true, false, 4);                                     \
    case 3: return CALL(0, true, true, 4);                                      \
    case 4: return CALL(1, false, false, 4);                                    \
    case 6: return CALL(1, true, false, 4);                                     \
    case 7: return CALL(1, true, true, 4);                                      \
    default: break;                                                             \
    }                                                                           \
    switch (slots) {                                                            \
    case 2: return CALL(1, false, true, 2);                                     \
    case 3: return CALL(1, false, true, 3);                                     \
    case 6: return CALL(1, false, true, 6);                                     \
    default: return CALL(1, false, true, 4);                                    \
    }

} // namespace

cudaError_t launchTransportMux(const RunParams& p, int mode, bool calib, const LaunchConfig& cfg, cudaStream_t stream)
{
    const bool smemTable = cfg.table_in_smem;
    const int slots = cfg.slots;
#define DXB_CALL(MO, CA, SM, MM) launchMux<MO, CA, SM, MM>(p, cfg, stream)
    DXB_MUX_DISPATCH(DXB_CALL)
#undef DXB_CALL
}

int transportMuxSlots(int mode, bool calib, bool smemTable, int slots)
{
    // must mirror DXB_MUX_DISPATCH: only the production variant is built for several slot counts
    if (mode == 1 && !calib && smemTable && (slots == 2 || slots == 3 || slots == 6))
        return slots;
    return 4;
}

int transportMuxOccupancy(int mode, bool calib, bool smemTable, int slots, int threads, size_t smem)
{
#define DXB_CALL(MO, CA, SM, MM) occupancyMux<MO, CA, SM, MM>(threads, smem)
    DXB_MUX_DISPATCH(DXB_CALL)
#undef DXB_CALL
}

} // namespace dxb
