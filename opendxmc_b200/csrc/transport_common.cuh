// transport_common.cuh — device helpers shared by the transport kernels (Philox, table coordinates, geometry,
// scoring, interaction samplers).  Included by transport.cu and transport_mux.cu.
#pragma once
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
constexpr float kU24 = 5.9604644775390625e-8f; // 2^-24
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kInvLn2 = 1.4426950408889634f;

// ------------------------------------------------------------------ Philox4x32-10
struct PhiloxBlock {
    unsigned int w[4];
    // k * 2^-24, k in [0, 2^24): exactly representable, in [0,1)
    __device__ __forceinline__ float u(int i) const { return static_cast<float>(w[i] >> 8) * kU24; }
    // the same uniform scaled by 2^24 (an integer-valued float): lets callers fold the 2^-24 into their own factor
    __device__ __forceinline__ float k(int i) const { return static_cast<float>(w[i] >> 8); }
};

// rk = the ten round keys (key + r * Weyl constants), precomputed on the host and read as constant-bank operands
__device__ __forceinline__ PhiloxBlock philox4x32_10(const unsigned int (&rk)[10][2], unsigned int c0, unsigned int c1, unsigned int c2)
{
    unsigned int x0 = c0, x1 = c1, x2 = c2, x3 = 0u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
        x0 = hi1 ^ x1 ^ rk[r][0];
        x1 = lo1;
        x2 = hi0 ^ x3 ^ rk[r][1];
        x3 = lo0;
    }
    PhiloxBlock b;
    b.w[0] = x0;
    b.w[1] = x1;
    b.w[2] = x2;
    b.w[3] = x3;
    return b;
}

// ------------------------------------------------------------------ table coordinates
struct TabPos {
    int i;
    float f;
};

// Grid coordinate of v >= 1 on a semi-log grid with PER = 2^LOG2PER nodes per octave (uniform inside the
// octave, physics.hpp): index = exponent and top mantissa bits, fraction = remaining mantissa bits.  Exact.
template <int LOG2PER, int N>
__device__ __forceinline__ TabPos tabPos(float v)
{
    TabPos p;
    const int bits = __float_as_int(fmaxf(v, 1.0f));
    constexpr int SH = 23 - LOG2PER;
    int i = (bits >> SH) - (127 << LOG2PER);
    float f = static_cast<float>(bits & ((1 << SH) - 1)) * (1.0f / static_cast<float>(1 << SH));
    if (i >= N - 1) {
        i = N - 2;
        f = 1.0f;
    }
    p.i = i;
    p.f = f;
    return p;
}
// value of grid node k (relative to the grid minimum)
template <int LOG2PER>
__device__ __forceinline__ float tabNode(int k)
{
    return __int_as_float(((127 << LOG2PER) + k) << (23 - LOG2PER));
}
constexpr int kLog2EPer = 6, kLog2XPer = 5;
static_assert((1 << kLog2EPer) == kDevEPerOctave && (1 << kLog2XPer) == kDevXPerOctave, "grid geometry");

__device__ __forceinline__ TabPos energyPos(float e) { return tabPos<kLog2EPer, kDevNE>(e); }

__device__ __forceinline__ float lerp(float a, float b, float f) { return fmaf(f, b - a, a); }

// voxel gather: L2 only (ld.global.cg).  The gathers are random over a 315 MB array, L1 cannot hold them; bypassing it
// leaves L1 to the interaction tables (+3 % measured against ld.global.nc, profiles/r01_tuning_log.txt).
__device__ __forceinline__ unsigned int loadVoxel(const unsigned int* p) { return __ldcg(p); }

// ------------------------------------------------------------------ geometry helpers
// Voxel of a point: linear index (x fastest) and whether the point lies inside the grid.  floor() of a negative
// coordinate is a negative integer, i.e. a huge unsigned one, so one unsigned compare per axis tests both faces.
// This IS the Woodcock exit test: a tentative step ends the history when its end point is outside the AABB.
__device__ __forceinline__ bool voxelIndex(const GridDev& G, float x, float y, float z, unsigned int& index)
{
    const unsigned int ix = static_cast<unsigned int>(__float2int_rd(fmaf(x, G.inv_dx, G.offx)));
    const unsigned int iy = static_cast<unsigned int>(__float2int_rd(fmaf(y, G.inv_dy, G.offy)));
    const unsigned int iz = static_cast<unsigned int>(__float2int_rd(fmaf(z, G.inv_dz, G.offz)));
    index = (iz * G.ny + iy) * G.nx + ix;
    return ix < static_cast<unsigned int>(G.nx) && iy < static_cast<unsigned int>(G.ny) && iz < static_cast<unsigned int>(G.nz);
}

// the same, plus the index of the brick (2^brick_shift voxels per edge) that holds the voxel
// dense box (DB builds of the pool kernel): the tentative point's voxel, and whether it lies inside the box (which lies inside the grid)
__device__ __forceinline__ bool voxelIndexBox(const GridDev& G, const RunParams& P, float x, float y, float z, unsigned int& index)
{
    const unsigned int ix = static_cast<unsigned int>(__float2int_rd(fmaf(x, G.inv_dx, G.offx)));
    const unsigned int iy = static_cast<unsigned int>(__float2int_rd(fmaf(y, G.inv_dy, G.offy)));
    const unsigned int iz = static_cast<unsigned int>(__float2int_rd(fmaf(z, G.inv_dz, G.offz)));
    index = (iz * G.ny + iy) * G.nx + ix;
    return ix - static_cast<unsigned int>(P.db_i0[0]) < static_cast<unsigned int>(P.db_n[0])
        && iy - static_cast<unsigned int>(P.db_i0[1]) < static_cast<unsigned int>(P.db_n[1])
        && iz - static_cast<unsigned int>(P.db_i0[2]) < static_cast<unsigned int>(P.db_n[2]);
}
// distance along (dx, dy, dz) from a point inside the axis-aligned box [lo, hi] to its boundary
__device__ __forceinline__ float boxExitDistance(float x, float y, float z, float dx, float dy, float dz, float lox, float loy, float loz,
    float hix, float hiy, float hiz)
{
    float t = 3.0e38f;
    if (dx != 0.0f)
        t = fminf(t, __fdividef((dx > 0.0f ? hix : lox) - x, dx));
    if (dy != 0.0f)
        t = fminf(t, __fdividef((dy > 0.0f ? hiy : loy) - y, dy));
    if (dz != 0.0f)
        t = fminf(t, __fdividef((dz > 0.0f ? hiz : loz) - z, dz));
    return fmaxf(t, 0.0f);
}
// distance at which the ray enters the dense box (0 when it starts inside); false on a miss
__device__ __forceinline__ bool boxEntryDistance(const RunParams& P, float x, float y, float z, float dx, float dy, float dz, float& tin)
{
    float tmin = 0.0f, tmax = 3.0e38f;
    const float p[3] = { x, y, z }, d[3] = { dx, dy, dz };
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (d[a] == 0.0f) {
            if (p[a] < P.db_lo[a] || p[a] > P.db_hi[a])
                tmax = -1.0f;
        } else {
            const float inv = __fdividef(1.0f, d[a]);
            const float t0 = (P.db_lo[a] - p[a]) * inv, t1 = (P.db_hi[a] - p[a]) * inv;
            tmin = fmaxf(tmin, fminf(t0, t1));
            tmax = fminf(tmax, fmaxf(t0, t1));
        }
    }
    tin = tmin;
    return tmax > tmin;
}
__device__ __forceinline__ bool voxelIndexBrick(const GridDev& G, const RunParams& P, float x, float y, float z, unsigned int& index, unsigned int& brick)
{
    const unsigned int ix = static_cast<unsigned int>(__float2int_rd(fmaf(x, G.inv_dx, G.offx)));
    const unsigned int iy = static_cast<unsigned int>(__float2int_rd(fmaf(y, G.inv_dy, G.offy)));
    const unsigned int iz = static_cast<unsigned int>(__float2int_rd(fmaf(z, G.inv_dz, G.offz)));
    index = (iz * G.ny + iy) * G.nx + ix;
    brick = ((iz >> P.brick_shift) * P.brick_ny + (iy >> P.brick_shift)) * P.brick_nx + (ix >> P.brick_shift);
    return ix < static_cast<unsigned int>(G.nx) && iy < static_cast<unsigned int>(G.ny) && iz < static_cast<unsigned int>(G.nz);
}

__device__ __forceinline__ void deflect(float& dx, float& dy, float& dz, float cosT, float phi)
{
    // dxmc::vectormath::peturb: rotate the direction by polar angle theta and azimuth phi.
    // (The library is built with -fmad=false: every fused multiply-add is written out, so that all kernel builds
    // round identically whatever the compiler's contraction heuristics.)
    const float sinT = sqrtf(fmaxf(0.0f, fmaf(-cosT, cosT, 1.0f)));
    float sinP, cosP;
    __sincosf(phi, &sinP, &cosP);
    float nx, ny, nz;
    if (fabsf(dz) < 0.99999f) {
        const float s2 = fmaf(-dz, dz, 1.0f);
        const float inv = rsqrtf(s2);
        const float tmp = s2 * inv;
        const float sti = sinT * inv;
        nx = fmaf(sti, fmaf(dx * dz, cosP, -(dy * sinP)), dx * cosT);
        ny = fmaf(sti, fmaf(dy * dz, cosP, dx * sinP), dy * cosT);
        nz = fmaf(-(tmp * sinT), cosP, dz * cosT);
    } else {
        nx = sinT * cosP;
        ny = sinT * sinP;
        nz = dz > 0.0f ? cosT : -cosT;
    }
    const float n = rsqrtf(fmaf(nx, nx, fmaf(ny, ny, nz * nz)));
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

// ------------------------------------------------------------------ interaction samplers (one try each)
// Klein-Nishina candidate e = E'/E uniform in [emin, 1], accepted with g(e)/gmax; MODE >= 1 multiplies the
// acceptance by the incoherent scatter function S(x)/Z (Livermore).  Returns true if accepted.
template <int MODE>
__device__ __forceinline__ bool comptonTry(const TablesDev& tab, int mat, float E, float r1, float ra, float& e, float& cosT)
{
    const float k = E * (1.0f / kElectronMass);
    const float emin = __fdividef(1.0f, fmaf(2.0f, k, 1.0f));
    const float gmaxInv = __fdividef(emin, fmaf(emin, emin, 1.0f));
    e = fmaf(1.0f - r1, emin, r1);
    const float einv = __fdividef(1.0f, e);
    const float t = fminf((1.0f - e) * einv * __fdividef(1.0f, k), 2.0f);
    const float sin2 = t * (2.0f - t);
    cosT = 1.0f - t;
    float g = (einv + e - sin2) * gmaxInv;
    if (MODE >= 1) {
        const float xs = E * (kDevXMinInv / kHc) * sqrtf(0.5f * t); // momentum transfer in units of the grid minimum
        float sfv;
        if (xs <= 1.0f) {
            sfv = __ldg(tab.sf + mat * kDevNX) * xs * xs;
        } else {
            const TabPos p = tabPos<kLog2XPer, kDevNX>(xs);
            const float* s = tab.sf + mat * kDevNX + p.i;
            sfv = lerp(__ldg(s), __ldg(s + 1), p.f);
        }
        g *= sfv;
    }
    return !(ra > g);
}

// ---- MODE 2 (impulse approximation), one extra Philox block per accepted Klein-Nishina x S(x) candidate ----
// The struck electron is drawn from the material's shell table by electron share (electrons not covered by the
// table form one unbound group with a common profile).  A bound electron carries a momentum component pz along the scattering vector, sampled from
// the one-parameter profile J(pz) = J0 sech^2(2 J0 pz) (normalised, J(0) = J0, inverse CDF
// pz = ln(u / (1 - u)) / (4 J0), atomic units) [D: profile shape defined by this project, DESIGN.md §6b]; the
// scattered energy follows from energy-momentum conservation for that pz (Doppler broadening, e.g. PENELOPE
// eq. 2.39).  The try is rejected if the shell cannot be ionised (binding >= E) or the energy transfer is below
// the binding energy.  Returns true and the ratio E'/E if accepted.
constexpr float kFineStructure = 7.2973525693e-3f;
__device__ __forceinline__ bool dopplerBroaden(const TablesDev& tab, int mat, float E, float e0, float cosT, float rShell, float rPz, float& eOut)
{
    const ShellDev* __restrict__ sh = tab.shells + mat * kMaxShells;
    const int n = __ldg(tab.n_shells + mat);
    float r = rShell, U = 0.0f, j0 = 0.0f;
    bool bound = false;
    for (int i = 0; i < n; ++i) {
        const float f = __ldg(&sh[i].nel_fraction);
        if (r < f) {
            U = __ldg(&sh[i].binding);
            j0 = __ldg(&sh[i].j0);
            bound = true;
            break;
        }
        r -= f;
    }
    eOut = e0;
    if (!bound) {
        // the electrons outside the table: unbound, one common profile (0: at rest)
        j0 = __ldg(tab.rest_j0 + mat);
        if (!(j0 > 0.0f))
            return true;
    }
    if (!(U < E))
        return false;
    const float u = fminf(fmaxf(rPz, kU24), 1.0f - kU24);
    float pz = __logf(__fdividef(u, 1.0f - u)) * __fdividef(0.25f, j0) * kFineStructure; // in units of m_e c
    pz = fminf(fmaxf(pz, -0.5f), 0.5f);
    // E'/E = e0 / b * (a + sign(pz) sqrt(a^2 - b (1 - t))), t = pz^2, a = 1 - t e0 cos, b = 1 - t e0^2.  Expanded, the
    // discriminant is t (q - t e0^2 sin^2) with q = (1 - e0)^2 + 2 e0 (1 - cos): no cancellation of O(1) terms in f32,
    // and sign(pz) sqrt(t) = pz.
    const float t = pz * pz;
    const float te0 = t * e0;
    const float a = fmaf(-te0, cosT, 1.0f);
    const float b = fmaf(-te0, e0, 1.0f);
    const float ome = 1.0f - e0;
    const float q = fmaf(2.0f * e0, 1.0f - cosT, ome * ome);
    const float disc = fmaxf(fmaf(-te0 * e0, fmaf(-cosT, cosT, 1.0f), q), 0.0f);
    const float e = __fdividef(e0, b) * fmaf(pz, sqrtf(disc), a);
    if (!(e > 0.0f) || !(fmaf(-E, e, E) > U))
        return false;
    eOut = fminf(e, 1.0f);
    return true;
}

// MODE 2 photoelectric absorption: picks the ionised shell by its share of the photoelectric cross section
// (shells with binding < E only) and decides whether a fluorescence photon is emitted.  Returns the energy of
// that photon (0: everything is absorbed locally).
__device__ __forceinline__ float photoFluorescence(const TablesDev& tab, int mat, float E, float rShell, float rYield)
{
    const ShellDev* __restrict__ sh = tab.shells + mat * kMaxShells;
    const int n = __ldg(tab.n_shells + mat);
    float r = rShell;
    for (int i = 0; i < n; ++i) {
        if (!(__ldg(&sh[i].binding) < E))
            continue;
        const float f = __ldg(&sh[i].photo_fraction);
        if (r < f) {
            const float ef = __ldg(&sh[i].fluor_energy);
            if (rYield < __ldg(&sh[i].fluor_yield) && ef >= kMinEnergy && ef < E)
                return ef;
            return 0.0f;
        }
        r -= f;
    }
    return 0.0f;
}

// isotropic direction from two uniforms
__device__ __forceinline__ void isotropic(float r0, float r1, float& dx, float& dy, float& dz)
{
    const float c = fmaf(2.0f, r0, -1.0f);
    const float s = sqrtf(fmaxf(0.0f, fmaf(-c, c, 1.0f)));
    float sp, cp;
    __sincosf(kTwoPi * r1, &sp, &cp);
    dx = s * cp;
    dy = s * sp;
    dz = c;
}

// Rayleigh: MODE 0 Thomson (pdf ~ (1 + cos^2) sin(theta), rejection from a box); MODE >= 1 samples
// q^2 ~ F(q)^2 on [0, qmax^2] from the tabulated cumulative A(x^2) (piecewise linear in x^2) and accepts with
// (1 + cos^2)/2.  Returns true if accepted.
template <int MODE>
__device__ __forceinline__ bool rayleighTry(const TablesDev& tab, int mat, float E, float r0, float r1, float& cosT)
{
    if (MODE == 0) {
        const float rr = r0 * 1.0886621079036347f; // 4 sqrt2 / (3 sqrt3)
        float s, c;
        __sincosf(kPiF * r1, &s, &c);
        cosT = c;
        return !(rr > fmaf(-s, s, 2.0f) * s);
    }
    const float xmaxs = E * (kDevXMinInv / kHc); // in units of the grid minimum
    const float xmax2 = xmaxs * xmaxs;
    const float* cdf = tab.ffcdf + mat * kDevNX;
    const float a0 = __ldg(cdf);
    float amax;
    int top; // the target lies below node `top`
    if (xmaxs <= 1.0f) {
        amax = a0 * xmax2; // F^2 flat below the first node: A = F0^2 x^2
        top = 0;
    } else {
        const TabPos p = tabPos<kLog2XPer, kDevNX>(xmaxs);
        const float xa = tabNode<kLog2XPer>(p.i), xb = tabNode<kLog2XPer>(p.i + 1);
        const float xa2 = xa * xa;
        const float fa = __fdividef(xmax2 - xa2, fmaf(xb, xb, -xa2));
        amax = lerp(__ldg(cdf + p.i), __ldg(cdf + p.i + 1), fminf(fmaxf(fa, 0.0f), 1.0f));
        top = p.i + 1;
    }
    const float target = r0 * amax;
    float x2;
    if (target <= a0) {
        x2 = __fdividef(target, a0);
    } else {
        // binary search: largest i with A[i] <= target (target <= amax <= A[top])
        int lo = 0, hi = top;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(cdf + mid) <= target)
                lo = mid;
            else
                hi = mid;
        }
        const float al = __ldg(cdf + lo), ah = __ldg(cdf + lo + 1);
        const float xa = tabNode<kLog2XPer>(lo), xb = tabNode<kLog2XPer>(lo + 1);
        const float f = ah > al ? __fdividef(target - al, ah - al) : 0.0f;
        const float xa2 = xa * xa;
        x2 = fmaf(f, fmaf(xb, xb, -xa2), xa2);
    }
    x2 = fminf(x2, xmax2);
    cosT = fmaf(-2.0f, __fdividef(x2, xmax2), 1.0f);
    return !(fmaf(cosT, cosT, 1.0f) * 0.5f < r1);
}

// ------------------------------------------------------------------ source sampling
// One history of the beam: exposure = h / ppe, fan / cone angles uniform in the collimation, energy from the tube's
// alias table (+ uniform inside the bin), weight = exposure weight x bowtie(|fan angle|), then the ray is moved to the
// grid's bounding box (World::transport).  Philox blocks 0 and 1 of the history.  Returns false if the ray misses the
// grid; E and w are valid either way (the caller counts the emitted energy E * w).
struct SourceSample {
    float px, py, pz, dx, dy, dz, E, w;
    float chord;  // length of the ray's path through the grid
    float spareK; // SPARE: word 1 of Philox block 1 as a 24-bit integer (the dense-box build draws its first flight from it)
};
template <bool SPARE = false>
__device__ __forceinline__ bool sampleSource(const RunParams& P, unsigned long long h, SourceSample& q)
{
    const GridDev& G = P.grid;
    const unsigned int qlo = static_cast<unsigned int>(h), qhi = static_cast<unsigned int>(h >> 32);
    const PhiloxBlock s0 = philox4x32_10(P.round_key, qlo, qhi, 0u);
    const unsigned long long ei = h / P.ppe;
    const ExposureDev* ex = P.exposures + ei;
    const float hx = __ldg(&ex->hx), hy = __ldg(&ex->hy);
    const float angx = fmaf(2.0f, s0.u(0), -1.0f) * hx;
    const float angy = fmaf(2.0f, s0.u(1), -1.0f) * hy;
    const int tube = __ldg(&ex->tube);
    const SpectrumDev& spc = P.spec[tube];
    float E;
    PhiloxBlock s1;
    if (SPARE) {
        s1 = philox4x32_10(P.round_key, qlo, qhi, 1u);
        q.spareK = s1.k(1);
    }
    if (spc.n <= 1) {
        E = spc.e0;
    } else {
        int idx = min(static_cast<int>(s0.u(2) * static_cast<float>(spc.n)), spc.n - 1);
        if (!(s0.u(3) < __ldg(spc.prob + idx)))
            idx = __ldg(spc.alias + idx);
        E = fmaf(static_cast<float>(idx), spc.step, spc.e0);
        if (idx < spc.n - 1) {
            if (!SPARE)
                s1 = philox4x32_10(P.round_key, qlo, qhi, 1u);
            E = fmaf(s1.u(0), spc.step, E);
        }
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
    const float sz = sqrtf(fmaxf(0.0f, fmaf(-sy, sy, fmaf(-sx, sx, 1.0f))));
    q.dx = fmaf(__ldg(&ex->dir[0]), sz, fmaf(__ldg(&ex->c1[0]), sy, __ldg(&ex->c0[0]) * sx));
    q.dy = fmaf(__ldg(&ex->dir[1]), sz, fmaf(__ldg(&ex->c1[1]), sy, __ldg(&ex->c0[1]) * sx));
    q.dz = fmaf(__ldg(&ex->dir[2]), sz, fmaf(__ldg(&ex->c1[2]), sy, __ldg(&ex->c0[2]) * sx));
    q.px = __ldg(&ex->pos[0]);
    q.py = __ldg(&ex->pos[1]);
    q.pz = __ldg(&ex->pos[2]);
    q.E = E;
    q.w = w;
    // slab test against the grid's bounding box
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
    if (!(tmax > tmin && E >= kMinEnergy))
        return false;
    q.px = fmaf(q.dx, tmin, q.px);
    q.py = fmaf(q.dy, tmin, q.py);
    q.pz = fmaf(q.dz, tmin, q.pz);
    q.chord = tmax - tmin;
    return true;
}

} // namespace
} // namespace dxb
