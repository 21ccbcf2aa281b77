// transport.cu — the photon-history kernels for sm_100a.
//
// One launch = a batch of histories of one beam.  Every lane owns one photon; when it dies the
// lane takes the next history id from its warp's pool (pools are carved from one global cursor in
// 256-history pieces), so warps stay full until the batch drains and concurrently resident
// histories belong to the same / adjacent exposures — the primary-beam slab of the voxel grid
// then lives in the 126 MB L2.
//
// Replaces (recalled DXMClib): Transport::runWorker -> exposure.sampleParticle
// -> World::transport -> AAVoxelGrid::woodcockTransport -> interactions::interact -> EnergyScore.
// The algorithm and its random-number protocol are restated independently in oracle/oracle.cpp.
//
// No tensor cores: nothing here is a dense contraction.  The hot loop is one dependent 8-byte
// voxel gather per tentative step (HBM/L2 sector bound) + FP32/INT issue (Philox, log, lerp).
#include "device_types.cuh"

#include <cstdio>

namespace dxb {

namespace {

constexpr float kElectronMass = 510.99895f;
constexpr float kHc = 12.398419843f;
constexpr float kMinEnergy = 1.0f;           // keV cut-off
constexpr float kRouletteThreshold = 0.1f;   // weight below which roulette is played
constexpr float kRouletteKill = 0.9f;        // kill probability
constexpr float kTwoPi = 6.283185307179586f;
constexpr float kPiF = 3.14159265358979f;

// ------------------------------------------------------------------ Philox4x32-10
struct Rng {
    unsigned int k0, k1;       // key
    unsigned int c0, c1, c2;   // counter words: history id lo/hi, block index
    unsigned int buf[4];
    int have;

    __device__ __forceinline__ void init(unsigned int seed_lo, unsigned int seed_hi, unsigned long long history)
    {
        k0 = seed_lo;
        k1 = seed_hi;
        c0 = static_cast<unsigned int>(history);
        c1 = static_cast<unsigned int>(history >> 32);
        c2 = 0;
        have = 0;
    }
    __device__ __forceinline__ void refill()
    {
        unsigned int x0 = c0, x1 = c1, x2 = c2, x3 = 0u;
        unsigned int ka = k0, kb = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const unsigned int hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
            const unsigned int hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
            x0 = hi1 ^ x1 ^ ka;
            x1 = lo1;
            x2 = hi0 ^ x3 ^ kb;
            x3 = lo0;
            ka += 0x9E3779B9u;
            kb += 0xBB67AE85u;
        }
        buf[0] = x0;
        buf[1] = x1;
        buf[2] = x2;
        buf[3] = x3;
        ++c2;
        have = 4;
    }
    // k * 2^-24, k in [0, 2^24): exactly representable, in [0,1)
    __device__ __forceinline__ float uniform()
    {
        if (have == 0)
            refill();
        // consume in order buf[0], buf[1], buf[2], buf[3]
        const int idx = 4 - have;
        --have;
        const unsigned int r = idx == 0 ? buf[0] : (idx == 1 ? buf[1] : (idx == 2 ? buf[2] : buf[3]));
        return static_cast<float>(r >> 8) * 5.9604644775390625e-8f;
    }
};

// ------------------------------------------------------------------ table coordinates
struct TabPos {
    int i;
    float f;
};

// grid coordinate of a value on a log-uniform grid with PER nodes per octave starting at 1.0:
// integer part from the float exponent, fraction from log2(mantissa) only.
template <int PER, int N>
__device__ __forceinline__ TabPos tabPos(float v)
{
    TabPos p;
    if (!(v > 1.0f)) {
        p.i = 0;
        p.f = 0.0f;
        return p;
    }
    const int bits = __float_as_int(v);
    const int ex = (bits >> 23) - 127;
    const float m = __int_as_float((bits & 0x007fffff) | 0x3f800000);
    const float t = log2f(m) * static_cast<float>(PER);
    int ti = static_cast<int>(t);
    ti = min(ti, PER - 1);
    int i = ex * PER + ti;
    float f = t - static_cast<float>(ti);
    if (i >= N - 1) {
        i = N - 2;
        f = 1.0f;
    }
    p.i = i;
    p.f = f;
    return p;
}

__device__ __forceinline__ float lerp(float a, float b, float f) { return fmaf(f, b - a, a); }

// ------------------------------------------------------------------ geometry helpers
__device__ __forceinline__ float exitDistance(const GridDev& g, float px, float py, float pz, float dx, float dy, float dz)
{
    const float tx = ((dx > 0.0f ? g.x1 : g.x0) - px) / dx;
    const float ty = ((dy > 0.0f ? g.y1 : g.y0) - py) / dy;
    const float tz = ((dz > 0.0f ? g.z1 : g.z0) - pz) / dz;
    // a zero direction component gives +-inf or NaN; fminf drops NaN, and -inf cannot occur for a
    // point inside the box except exactly on a face
    float t = 3.0e38f;
    if (dx != 0.0f)
        t = fminf(t, tx);
    if (dy != 0.0f)
        t = fminf(t, ty);
    if (dz != 0.0f)
        t = fminf(t, tz);
    return fmaxf(t, 0.0f);
}

__device__ __forceinline__ void deflect(float& dx, float& dy, float& dz, float cosT, float phi)
{
    // dxmc::vectormath::peturb: rotate the direction by polar angle theta and azimuth phi
    const float sinT = sqrtf(fmaxf(0.0f, 1.0f - cosT * cosT));
    float sinP, cosP;
    __sincosf(phi, &sinP, &cosP);
    float nx, ny, nz;
    if (fabsf(dz) < 0.99999f) {
        const float tmp = sqrtf(1.0f - dz * dz);
        const float inv = 1.0f / tmp;
        nx = dx * cosT + sinT * (dx * dz * cosP - dy * sinP) * inv;
        ny = dy * cosT + sinT * (dy * dz * cosP + dx * sinP) * inv;
        nz = dz * cosT - tmp * sinT * cosP;
    } else {
        nx = sinT * cosP;
        ny = sinT * sinP;
        nz = dz > 0.0f ? cosT : -cosT;
    }
    const float n = rsqrtf(nx * nx + ny * ny + nz * nz);
    dx = nx * n;
    dy = ny * n;
    dz = nz * n;
}

// ------------------------------------------------------------------ scoring
// Warp-aggregated fixed-point tally update.  `mask` = the lanes that score in this phase (all of them call
// this function together); lanes of the mask that hit the same voxel are merged before the atomics (exact
// integer sums, so the result does not depend on the merge or on arrival order).
__device__ __forceinline__ void scoreEnergy(unsigned int mask, unsigned long long* __restrict__ tally, unsigned int voxel, float edep,
    float scale_e, float scale_e2)
{
    unsigned long long e = static_cast<unsigned long long>(__float2ll_rn(edep * scale_e));
    unsigned long long e2 = static_cast<unsigned long long>(__float2ll_rn(edep * edep * scale_e2));
    unsigned int n = 1;
    const unsigned int peers = __match_any_sync(mask, voxel);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    if (peers != (1u << lane)) {
        // rare path: several lanes on one voxel; every peer walks the peer set
        unsigned long long se = 0, se2 = 0;
        unsigned int walk = peers;
        while (walk) {
            const int src = __ffs(walk) - 1;
            walk &= walk - 1;
            se += __shfl_sync(peers, e, src);
            se2 += __shfl_sync(peers, e2, src);
        }
        e = se;
        e2 = se2;
        n = __popc(peers);
    }
    if (lane == leader) {
        unsigned long long* t = tally + static_cast<size_t>(voxel) * 4;
        atomicAdd(t + 0, e);
        atomicAdd(t + 1, e2);
        atomicAdd(t + 2, static_cast<unsigned long long>(n));
    }
}

// ------------------------------------------------------------------ interaction samplers
template <int MODE>
__device__ __forceinline__ float comptonScatter(Rng& rng, const TablesDev& tab, int mat, float& E, float& dx, float& dy, float& dz)
{
    // Klein-Nishina by rejection on g(e)/gmax with e = E'/E uniform in [emin, 1];
    // MODE 1 multiplies the acceptance by the incoherent scatter function S(x)/Z.
    const float k = E * (1.0f / kElectronMass);
    const float emin = 1.0f / (1.0f + 2.0f * k);
    const float gmaxInv = emin / (1.0f + emin * emin);
    float e, cosT;
    bool rejected;
    do {
        const float r1 = rng.uniform();
        e = r1 + (1.0f - r1) * emin;
        const float t = fminf((1.0f - e) / (k * e), 2.0f);
        const float sin2 = t * (2.0f - t);
        cosT = 1.0f - t;
        float g = (1.0f / e + e - sin2) * gmaxInv;
        if (MODE >= 1) {
            const float x = E * (1.0f / kHc) * sqrtf(0.5f * t); // momentum transfer [1/A]
            float sfv;
            const float xs = x * kDevXMinInv;
            if (xs <= 1.0f) {
                sfv = __ldg(tab.sf + mat * kDevNX) * xs * xs;
            } else {
                const TabPos p = tabPos<kDevXPerOctave, kDevNX>(xs);
                const float* s = tab.sf + mat * kDevNX + p.i;
                sfv = lerp(__ldg(s), __ldg(s + 1), p.f);
            }
            g *= sfv;
        }
        rejected = rng.uniform() > g;
    } while (rejected);
    const float phi = kTwoPi * rng.uniform();
    deflect(dx, dy, dz, cosT, phi);
    const float E0 = E;
    E = E0 * e;
    return E0 - E;
}

template <int MODE>
__device__ __forceinline__ void rayleighScatter(Rng& rng, const TablesDev& tab, int mat, float E, float& dx, float& dy, float& dz)
{
    float cosT;
    if (MODE == 0) {
        // Thomson: pdf ~ (1 + cos^2) sin(theta), rejection from a box
        bool reject;
        do {
            const float r1 = rng.uniform() * 1.0886621079036347f; // 4 sqrt2 / (3 sqrt3)
            const float theta = kPiF * rng.uniform();
            float s, c;
            __sincosf(theta, &s, &c);
            cosT = c;
            reject = r1 > (2.0f - s * s) * s;
        } while (reject);
    } else {
        // q^2 ~ F(q)^2 on [0, qmax^2] via the tabulated cumulative A(x^2) (piecewise linear in x^2),
        // then accept with (1 + cos^2)/2.
        const float xmax = E * (1.0f / kHc);
        const float xmax2 = xmax * xmax;
        const float* cdf = tab.ffcdf + mat * kDevNX;
        float amax;
        {
            const float xs = xmax * kDevXMinInv;
            if (xs <= 1.0f) {
                amax = __ldg(cdf) * xs * xs; // F^2 flat below the first node: A = F0^2 x^2
            } else {
                const TabPos p = tabPos<kDevXPerOctave, kDevNX>(xs);
                const float xa = exp2f(static_cast<float>(p.i) * (1.0f / kDevXPerOctave)) * (1.0f / kDevXMinInv);
                const float xb = exp2f(static_cast<float>(p.i + 1) * (1.0f / kDevXPerOctave)) * (1.0f / kDevXMinInv);
                const float fa = (xmax2 - xa * xa) / (xb * xb - xa * xa);
                amax = lerp(__ldg(cdf + p.i), __ldg(cdf + p.i + 1), fminf(fmaxf(fa, 0.0f), 1.0f));
            }
        }
        bool reject;
        do {
            const float target = rng.uniform() * amax;
            float x2;
            const float a0 = __ldg(cdf);
            if (target <= a0) {
                const float x0 = 1.0f / kDevXMinInv;
                x2 = target / a0 * x0 * x0;
            } else {
                // binary search: largest i with A[i] <= target
                int lo = 0, hi = kDevNX - 1;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (__ldg(cdf + mid) <= target)
                        lo = mid;
                    else
                        hi = mid;
                }
                const float al = __ldg(cdf + lo), ah = __ldg(cdf + lo + 1);
                const float xa = exp2f(static_cast<float>(lo) * (1.0f / kDevXPerOctave)) * (1.0f / kDevXMinInv);
                const float xb = exp2f(static_cast<float>(lo + 1) * (1.0f / kDevXPerOctave)) * (1.0f / kDevXMinInv);
                const float f = ah > al ? (target - al) / (ah - al) : 0.0f;
                x2 = xa * xa + f * (xb * xb - xa * xa);
            }
            x2 = fminf(x2, xmax2);
            cosT = 1.0f - 2.0f * x2 / xmax2;
            reject = (1.0f + cosT * cosT) * 0.5f < rng.uniform();
        } while (reject);
    }
    const float phi = kTwoPi * rng.uniform();
    deflect(dx, dy, dz, cosT, phi);
}

// ------------------------------------------------------------------ the history kernel
//
// Warp-level phase machine.  Every lane owns at most one photon and is in one of three states:
//   STEP  tentative Woodcock steps (cheap, executed by most lanes every iteration),
//   WAIT  a tentative collision was accepted as real; the (expensive, divergent) interaction sampler has
//         not run yet,
//   DEAD  no photon; waiting for a new history.
// Each iteration the warp picks ONE phase by vote: the interaction sampler only runs once `interact_threshold`
// lanes wait for it (or nobody can step), dead lanes are only refilled once `refill_threshold` of them are dead.
// Source sampling is warp-cooperative: all 32 lanes sample one history each (full SIMD efficiency, independent
// of how many lanes are dead), photons that hit the grid are compacted into a per-warp shared-memory buffer,
// and dead lanes pop from it.  This replaces the one-lane-at-a-time refill/interaction of the first version,
// whose ncu capture showed 5.4 of 32 lanes active per instruction (profiles/r01_transport_v1.md).
struct Photon {
    float px, py, pz;
    float dx, dy, dz;
    float E, w;
    float remaining; // distance to the grid exit along the current direction
};

constexpr int kStStep = 0, kStWait = 1, kStDead = 2;
constexpr int kWarpBufFloats = kSourceBufWords * 32; // px py pz dx dy dz E w remaining histOffset epos.i epos.f muMax
static_assert(kSourceBufWords == 13, "source buffer layout");

template <int MODE, bool CALIB, bool SMEM_TABLE>
__global__ void __launch_bounds__(256, 3) transportKernel(const __grid_constant__ RunParams P)
{
    extern __shared__ float s_dyn[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nWarps = blockDim.x >> 5;
    float* __restrict__ sbuf = s_dyn + warp * kWarpBufFloats; // SoA: word f of entry k at sbuf[f * 32 + k]
    float* __restrict__ s_tot = s_dyn + nWarps * kWarpBufFloats;
    if (SMEM_TABLE) {
        const int n = P.tab.n_mat * kDevNE;
        for (int i = threadIdx.x; i < n; i += blockDim.x)
            s_tot[i] = P.tab.tot[i];
        __syncthreads();
    }
    const float* __restrict__ totTable = SMEM_TABLE ? s_tot : P.tab.tot;
    const GridDev& G = P.grid;
    const unsigned int laneLt = (1u << lane) - 1u;

    // warp-level pool of local history indices (carved from the global cursor in 256-history pieces)
    unsigned long long poolNext = 0, poolEnd = 0;
    bool drained = false;
    int bufCount = 0;                 // entries in the warp's source buffer
    unsigned long long bufBase = 0;   // global history id of buffer offset 0

    Rng rng;
    Photon ph;
    int status = kStDead;
    TabPos epos;
    float muMax = 1.0f, muMaxInv = 1.0f;
    unsigned int voxel = 0;
    int mat = 0;
    epos.i = 0;
    epos.f = 0.0f;
    ph.remaining = 0.0f;
    rng.init(P.seed_lo, P.seed_hi, 0ull);

    unsigned int nSteps = 0, nInter = 0, nDep = 0, nHist = 0;
    unsigned long long emitted = 0;

    for (;;) {
        const unsigned int mDead = __ballot_sync(0xffffffffu, status == kStDead);
        const unsigned int mWait = __ballot_sync(0xffffffffu, status == kStWait);
        const int nDead = __popc(mDead), nWait = __popc(mWait), nStep = 32 - nDead - nWait;
        const bool canRefill = !(drained && bufCount == 0);
        int phase; // 0 step, 1 interact, 2 refill
        if (canRefill && nDead >= P.refill_threshold)
            phase = 2;
        else if (nWait >= P.interact_threshold)
            phase = 1;
        else if (nStep > 0)
            phase = 0;
        else if (nWait > 0)
            phase = 1;
        else if (canRefill)
            phase = 2;
        else
            break;

        if (phase == 2) {
            // ------------------------------------------------------------ refill
            if (bufCount == 0) {
                // warp-cooperative source sampling of the next (up to) 32 histories
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
                const unsigned long long blk = poolNext / kShardBlock;
                bufBase = (blk * P.world + P.rank) * kShardBlock + (poolNext % kShardBlock);
                const unsigned long long h = bufBase + lane;
                bool hit = false;
                Photon q;
                TabPos qpos;
                float qmu = 1.0f;
                qpos.i = 0;
                qpos.f = 0.0f;
                q.px = q.py = q.pz = q.dx = q.dy = q.dz = q.E = q.w = q.remaining = 0.0f;
                if (lane < nb && h < P.n_total) {
                    Rng srng;
                    srng.init(P.seed_lo, P.seed_hi, h);
                    const unsigned long long ei = h / P.ppe;
                    const ExposureDev* ex = P.exposures + ei;
                    const float hx = __ldg(&ex->hx), hy = __ldg(&ex->hy);
                    const float angx = (2.0f * srng.uniform() - 1.0f) * hx;
                    const float angy = (2.0f * srng.uniform() - 1.0f) * hy;
                    const int tube = __ldg(&ex->tube);
                    const SpectrumDev& sp = P.spec[tube];
                    float E;
                    if (sp.n <= 1) {
                        E = sp.e0;
                    } else {
                        const float r0 = srng.uniform();
                        int idx = min(static_cast<int>(r0 * static_cast<float>(sp.n)), sp.n - 1);
                        const float r1 = srng.uniform();
                        if (!(r1 < __ldg(sp.prob + idx)))
                            idx = __ldg(sp.alias + idx);
                        const float r2 = srng.uniform();
                        E = sp.e0 + static_cast<float>(idx) * sp.step;
                        if (idx < sp.n - 1)
                            E += r2 * sp.step;
                    }
                    float w = __ldg(&ex->weight);
                    const BowtieDev& bt = P.bow[tube];
                    if (bt.n > 0) {
                        const float a = fabsf(angx);
                        float bw;
                        if (a <= __ldg(bt.angle)) {
                            bw = __ldg(bt.weight);
                        } else if (a >= __ldg(bt.angle + bt.n - 1)) {
                            bw = __ldg(bt.weight + bt.n - 1);
                        } else {
                            int i = 1;
                            while (__ldg(bt.angle + i) < a)
                                ++i;
                            const float a0 = __ldg(bt.angle + i - 1), a1 = __ldg(bt.angle + i);
                            bw = lerp(__ldg(bt.weight + i - 1), __ldg(bt.weight + i), (a - a0) / (a1 - a0));
                        }
                        w *= bw;
                    }
                    const float sx = __sinf(angx), sy = __sinf(angy);
                    const float sz = sqrtf(fmaxf(0.0f, 1.0f - sx * sx - sy * sy));
                    q.dx = __ldg(&ex->c0[0]) * sx + __ldg(&ex->c1[0]) * sy + __ldg(&ex->dir[0]) * sz;
                    q.dy = __ldg(&ex->c0[1]) * sx + __ldg(&ex->c1[1]) * sy + __ldg(&ex->dir[1]) * sz;
                    q.dz = __ldg(&ex->c0[2]) * sx + __ldg(&ex->c1[2]) * sy + __ldg(&ex->dir[2]) * sz;
                    q.px = __ldg(&ex->pos[0]);
                    q.py = __ldg(&ex->pos[1]);
                    q.pz = __ldg(&ex->pos[2]);
                    q.E = E;
                    q.w = w;
                    ++nHist;
                    emitted += static_cast<unsigned long long>(__float2ll_rn(E * w * 65536.0f));
                    // move to the grid AABB (World::transport)
                    const float ix = 1.0f / q.dx, iy = 1.0f / q.dy, iz = 1.0f / q.dz;
                    float tmin = 0.0f, tmax = 3.0e38f;
                    float t0 = (G.x0 - q.px) * ix, t1 = (G.x1 - q.px) * ix;
                    if (q.dx == 0.0f) {
                        if (q.px < G.x0 || q.px > G.x1)
                            tmax = -1.0f;
                    } else {
                        tmin = fmaxf(tmin, fminf(t0, t1));
                        tmax = fminf(tmax, fmaxf(t0, t1));
                    }
                    t0 = (G.y0 - q.py) * iy;
                    t1 = (G.y1 - q.py) * iy;
                    if (q.dy == 0.0f) {
                        if (q.py < G.y0 || q.py > G.y1)
                            tmax = -1.0f;
                    } else {
                        tmin = fmaxf(tmin, fminf(t0, t1));
                        tmax = fminf(tmax, fmaxf(t0, t1));
                    }
                    t0 = (G.z0 - q.pz) * iz;
                    t1 = (G.z1 - q.pz) * iz;
                    if (q.dz == 0.0f) {
                        if (q.pz < G.z0 || q.pz > G.z1)
                            tmax = -1.0f;
                    } else {
                        tmin = fmaxf(tmin, fminf(t0, t1));
                        tmax = fminf(tmax, fmaxf(t0, t1));
                    }
                    if (tmax > tmin && E >= kMinEnergy) {
                        q.px = fmaf(q.dx, tmin, q.px);
                        q.py = fmaf(q.dy, tmin, q.py);
                        q.pz = fmaf(q.dz, tmin, q.pz);
                        q.remaining = tmax - tmin;
                        qpos = tabPos<kDevEPerOctave, kDevNE>(E);
                        qmu = lerp(__ldg(P.tab.majorant + qpos.i), __ldg(P.tab.majorant + qpos.i + 1), qpos.f);
                        hit = true;
                    }
                }
                poolNext += nb;
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
                    sbuf[8 * 32 + k] = q.remaining;
                    sbuf[9 * 32 + k] = __int_as_float(lane);
                    sbuf[10 * 32 + k] = __int_as_float(qpos.i);
                    sbuf[11 * 32 + k] = qpos.f;
                    sbuf[12 * 32 + k] = qmu;
                }
                bufCount = __popc(mHit);
                __syncwarp();
            }
            // dead lanes pop from the top of the buffer
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
                    ph.remaining = sbuf[8 * 32 + k];
                    const unsigned long long h = bufBase + static_cast<unsigned int>(__float_as_int(sbuf[9 * 32 + k]));
                    epos.i = __float_as_int(sbuf[10 * 32 + k]);
                    epos.f = sbuf[11 * 32 + k];
                    muMax = sbuf[12 * 32 + k];
                    muMaxInv = 1.0f / muMax;
                    // the transport stream of a history starts at Philox block 2 (blocks 0-1 belong to the source)
                    rng.init(P.seed_lo, P.seed_hi, h);
                    rng.c2 = 2u;
                    status = kStStep;
                }
            }
            bufCount -= min(nDead, bufCount);
            __syncwarp();
        } else if (phase == 0) {
            // ------------------------------------------------------------ one tentative Woodcock step
            float kerma = 0.0f;
            if (status == kStStep) {
                const float s = -__logf(1.0f - rng.uniform()) * muMaxInv;
                if (s >= ph.remaining) {
                    status = kStDead; // left the grid
                } else {
                    ++nSteps;
                    ph.px = fmaf(ph.dx, s, ph.px);
                    ph.py = fmaf(ph.dy, s, ph.py);
                    ph.pz = fmaf(ph.dz, s, ph.pz);
                    ph.remaining -= s;
                    int vx = static_cast<int>((ph.px - G.x0) * G.inv_dx);
                    int vy = static_cast<int>((ph.py - G.y0) * G.inv_dy);
                    int vz = static_cast<int>((ph.pz - G.z0) * G.inv_dz);
                    vx = min(max(vx, 0), G.nx - 1);
                    vy = min(max(vy, 0), G.ny - 1);
                    vz = min(max(vz, 0), G.nz - 1);
                    voxel = (static_cast<unsigned int>(vz) * G.ny + vy) * G.nx + vx;
                    const uint2 vox = __ldg(G.voxels + voxel);
                    const float rho = __uint_as_float(vox.x);
                    mat = static_cast<int>(vox.y);
                    const float* tt = totTable + mat * kDevNE + epos.i;
                    const float mu = rho * lerp(tt[0], tt[1], epos.f);
                    const float r = rng.uniform();
                    if (CALIB && mat == P.score_material) {
                        // collision estimator of air kerma: every tentative collision carries 1/mu_max of track length
                        const float* et = P.tab.etr + mat * kDevNE + epos.i;
                        kerma = ph.w * ph.E * lerp(__ldg(et), __ldg(et + 1), epos.f) * muMaxInv;
                    }
                    if (r * muMax < mu)
                        status = kStWait;
                }
            }
            if (CALIB) {
                const unsigned int mScore = __ballot_sync(0xffffffffu, kerma > 0.0f);
                if (kerma > 0.0f)
                    scoreEnergy(mScore, G.tally, voxel, kerma, P.tally_scale_e, P.tally_scale_e2);
            }
        } else {
            // ------------------------------------------------------------ real interactions of the waiting lanes
            float edep = 0.0f;
            if (status == kStWait) {
                ++nInter;
                const float4 a = __ldg(P.tab.att + mat * kDevNE + epos.i);
                const float4 b = __ldg(P.tab.att + mat * kDevNE + epos.i + 1);
                const float aPhoto = lerp(a.x, b.x, epos.f);
                const float aIncoh = lerp(a.y, b.y, epos.f);
                const float aTot = lerp(a.w, b.w, epos.f);
                const float r2 = rng.uniform() * aTot;
                bool alive = true, energyChanged = false, dirChanged = false;
                if (r2 < aPhoto) {
                    edep = ph.E * ph.w;
                    ph.E = 0.0f;
                    alive = false;
                } else if (r2 < aPhoto + aIncoh) {
                    const float de = comptonScatter<MODE>(rng, P.tab, mat, ph.E, ph.dx, ph.dy, ph.dz);
                    edep = de * ph.w;
                    energyChanged = true;
                    dirChanged = true;
                } else {
                    rayleighScatter<MODE>(rng, P.tab, mat, ph.E, ph.dx, ph.dy, ph.dz);
                    dirChanged = true;
                }
                if (alive) {
                    if (ph.E < kMinEnergy) {
                        edep += ph.E * ph.w;
                        ph.E = 0.0f;
                        alive = false;
                    } else if (ph.w < kRouletteThreshold) {
                        if (rng.uniform() < kRouletteKill)
                            alive = false;
                        else
                            ph.w *= 1.0f / (1.0f - kRouletteKill);
                    }
                }
                if (alive) {
                    if (energyChanged) {
                        epos = tabPos<kDevEPerOctave, kDevNE>(ph.E);
                        muMax = lerp(__ldg(P.tab.majorant + epos.i), __ldg(P.tab.majorant + epos.i + 1), epos.f);
                        muMaxInv = 1.0f / muMax;
                    }
                    if (dirChanged)
                        ph.remaining = exitDistance(G, ph.px, ph.py, ph.pz, ph.dx, ph.dy, ph.dz);
                    status = kStStep;
                } else {
                    status = kStDead;
                }
                if (CALIB)
                    edep = 0.0f;
            }
            if (!CALIB) {
                const unsigned int mScore = __ballot_sync(0xffffffffu, edep > 0.0f);
                if (edep > 0.0f) {
                    ++nDep;
                    scoreEnergy(mScore, G.tally, voxel, edep, P.tally_scale_e, P.tally_scale_e2);
                }
            }
        }
    }

    // ---------------- statistics
    unsigned long long v[5] = { nSteps, nInter, nDep, emitted, nHist };
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
    uint2* __restrict__ out, size_t n, unsigned int* __restrict__ maxDensityBits /* [256] */)
{
    __shared__ unsigned int s_max[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        s_max[i] = 0u;
    __syncthreads();
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        float rho = static_cast<float>(density[i]);
        if (!(rho > 0.0f))
            rho = 0.0f;
        const unsigned int m = material[i];
        out[i] = make_uint2(__float_as_uint(rho), m);
        atomicMax(&s_max[m], __float_as_uint(rho)); // non-negative floats order like their bit patterns
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (s_max[i])
            atomicMax(&maxDensityBits[i], s_max[i]);
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
__global__ void energyToDoseKernel(const unsigned long long* __restrict__ tally, const uint2* __restrict__ voxels,
    double* __restrict__ dose, double* __restrict__ variance, unsigned long long* __restrict__ events, size_t n,
    double inv_scale_e, double inv_scale_e2, double factor, double voxel_volume)
{
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const ulonglong4 t = reinterpret_cast<const ulonglong4*>(tally)[i];
        if (t.z == 0)
            continue;
        const double rho = static_cast<double>(__uint_as_float(voxels[i].x));
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

// reference post-processing (R:src/libopendxmc/simulationpipeline.cpp:180-185,206-211,221-229)
__global__ void postprocessKernel(const double* __restrict__ in, const uint2* __restrict__ voxels, double* __restrict__ out,
    size_t n, int maskAir, double scale)
{
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        double v = in[i];
        if (maskAir && voxels[i].y == 0)
            v = 0.0;
        out[i] = v * scale;
    }
}

__global__ void u64ToDoubleKernel(const unsigned long long* __restrict__ in, const uint2* __restrict__ voxels,
    double* __restrict__ out, size_t n, int maskAir)
{
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        double v = static_cast<double>(in[i]);
        if (maskAir && voxels[i].y == 0)
            v = 0.0;
        out[i] = v;
    }
}

// block max reduction for the uGy decision (max(dose) < 1)
__global__ void maxKernel(const double* __restrict__ in, const uint2* __restrict__ voxels, size_t n, int maskAir,
    unsigned long long* __restrict__ outBits)
{
    double m = 0.0;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        double v = in[i];
        if (maskAir && voxels[i].y == 0)
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
    const uint2* __restrict__ voxels, const unsigned char* __restrict__ organ, size_t n, double voxel_volume,
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
        const double m = static_cast<double>(__uint_as_float(voxels[i].x)) * voxel_volume;
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
    const TabPos p = tabPos<kDevEPerOctave, kDevNE>(energy[i]);
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
    const TabPos p = tabPos<kDevEPerOctave, kDevNE>(energy[i]);
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
        } else {
            DXB_DISPATCH(1, false)
        }
    } else {
        if (mode == 0) {
            DXB_DISPATCH(0, true)
        } else {
            DXB_DISPATCH(1, true)
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
        } else {
            if (smemTable) DXB_OCC(1, false, true) else DXB_OCC(1, false, false)
        }
    } else {
        if (mode == 0) {
            if (smemTable) DXB_OCC(0, true, true) else DXB_OCC(0, true, false)
        } else {
            if (smemTable) DXB_OCC(1, true, true) else DXB_OCC(1, true, false)
        }
    }
#undef DXB_OCC
    return e == cudaSuccess ? nb : 0;
}

void launchPackVoxels(const double* density, const unsigned char* material, uint2* out, size_t n, unsigned int* maxBits, cudaStream_t s)
{
    packVoxelsKernel<<<148 * 8, 256, 0, s>>>(density, material, out, n, maxBits);
}
void launchMajorant(const float* tot, const unsigned int* maxBits, int n_mat, float* majorant, cudaStream_t s)
{
    majorantKernel<<<(kDevNE + 127) / 128, 128, 0, s>>>(tot, maxBits, n_mat, majorant);
}
void launchEnergyToDose(const unsigned long long* tally, const uint2* voxels, double* dose, double* variance,
    unsigned long long* events, size_t n, double inv_e, double inv_e2, double factor, double vol, cudaStream_t s)
{
    energyToDoseKernel<<<148 * 8, 256, 0, s>>>(tally, voxels, dose, variance, events, n, inv_e, inv_e2, factor, vol);
}
void launchTallyToEnergy(const unsigned long long* tally, double* e, double* e2, unsigned long long* cnt, size_t n,
    double inv_e, double inv_e2, cudaStream_t s)
{
    tallyToEnergyKernel<<<148 * 8, 256, 0, s>>>(tally, e, e2, cnt, n, inv_e, inv_e2);
}
void launchPeerReduce(unsigned long long* dst, const unsigned long long* const* peers, int n_peers, size_t n_words, cudaStream_t s)
{
    peerReduceKernel<<<148 * 8, 256, 0, s>>>(dst, peers, n_peers, n_words);
}
void launchPostprocess(const double* in, const uint2* voxels, double* out, size_t n, int maskAir, double scale, cudaStream_t s)
{
    postprocessKernel<<<148 * 8, 256, 0, s>>>(in, voxels, out, n, maskAir, scale);
}
void launchU64ToDouble(const unsigned long long* in, const uint2* voxels, double* out, size_t n, int maskAir, cudaStream_t s)
{
    u64ToDoubleKernel<<<148 * 8, 256, 0, s>>>(in, voxels, out, n, maskAir);
}
void launchMax(const double* in, const uint2* voxels, size_t n, int maskAir, unsigned long long* outBits, cudaStream_t s)
{
    maxKernel<<<148 * 4, 256, 0, s>>>(in, voxels, n, maskAir, outBits);
}
void launchOrganDose(const double* dose, const double* variance, const uint2* voxels, const unsigned char* organ, size_t n,
    double vol, double* energy, double* mass, unsigned long long* count, double* varEnergy, cudaStream_t s)
{
    organDoseKernel<<<148 * 4, 256, 0, s>>>(dose, variance, voxels, organ, n, vol, energy, mass, count, varEnergy);
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
    segmentKernel<<<148 * 8, 256, 0, s>>>(hu, n, sep, n_sep, matAtt, waterAttDens, airAttDens, material, density);
}

} // namespace dxb
