// oracle.cpp — CPU restatement of the dxmc::Transport photon-history path.  TEST INFRASTRUCTURE.
//
// Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may use it.
// Double precision, std::thread workers claiming exposure indices from one atomic counter — the
// threading scheme of DXMClib's Transport (recalled, SURVEY.md §2 "CPU parallelism").
//
// PARITY UNPINNED (see oracle.h).  Sources this file follows:
//  [V] verified from OpenDXMC:
//      driver order / post-processing      R:src/libopendxmc/simulationpipeline.cpp:124-235
//      array layout, x fastest             R:src/libopendxmc/otherphantomimportpipeline.cpp:44
//      grid centred on the origin          R:src/libopendxmc/datacontainer.cpp:174-178
//      photon direction from half-angles   R:src/libopendxmc/beamactorcontainer.cpp:44-69
//      dual-source interleave              R:src/libopendxmc/beamactorcontainer.cpp:134-146
//      per-organ dose                      R:src/libopendxmc/dosetablepipeline.cpp:60-84
//      CT segmentation                     R:src/libopendxmc/ctsegmentationpipeline.cpp:131-156
//  [R] recalled DXMClib design intent, structured like the library:
//      Particle, RandomState, World::transport, AAVoxelGrid::woodcockTransport,
//      interactions::{photoelectricEffect, comptonScatter, rayleightScatter, interact},
//      EnergyScore, DoseScore, Transport::runWorker, CT*Beam::calibrationFactor.
//  [D] this project's own definitions where DXMClib's are unknown: Philox streams keyed by history
//      id (north-star requirement), table interpolation rule, CTDI phantom voxelisation and the
//      collision kerma estimator, bowtie/AEC normalisation.  DESIGN.md lists them.
//
// Random-number protocol (must match opendxmc_b200/csrc/transport.cu draw for draw):
//   stream(history) = Philox4x32-10, key = seed, counter = (history lo, history hi, block, 0),
//   uniform u = (word >> 8) * 2^-24 in [0,1).  The source sampler draws sequentially from blocks 0-1; transport
//   starts at block 2 and consumes one whole block per event:
//     step pair        w0,w1 = length / acceptance of tentative step A, w2,w3 = the same for step B (B only if A was virtual)
//     interaction try  w0 = channel choice (first try only), w1,w2 = Compton candidate / rejection test, w3 = azimuth
//     Rayleigh try     w0 = CDF target (Thomson: rejection variable), w1 = rejection test (Thomson: polar angle), w2 = azimuth
//     roulette         w0
#include "oracle.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

namespace {

constexpr double PI = 3.14159265358979323846;
constexpr double ELECTRON_REST_MASS = 510.99895; // keV
constexpr double HC = 12.398419843;              // keV Angstrom
constexpr double MIN_ENERGY = 1.0;               // keV
constexpr double RUSSIAN_ROULETTE_THRESHOLD = 0.1;
constexpr double RUSSIAN_ROULETTE_PROBABILITY = 0.9;
constexpr double KEV_PER_GRAM_TO_MGY = 1.602176634e-10;
constexpr uint64_t SHARD_BLOCK = 65536;

// ------------------------------------------------------------------ device mirroring
// By default the oracle reproduces the places where the device stores a QUANTISED copy of its input (24-bit voxel
// densities, f32 majorant / bowtie knots / alias acceptance values / shell constants / exposure geometry), so that both
// sides consume the same numbers and differ by arithmetic rounding only.  orc_set_device_mirroring(0) switches all of
// that off: the oracle then runs on the caller's f64 data as handed over.  tests/test_oracle_known_answers.py and the
// GPU suite use the un-mirrored mode to show that those quantisations are harmless, not merely shared.
bool g_mirror = true;
inline double mf(double v) { return g_mirror ? static_cast<double>(static_cast<float>(v)) : v; }

// ------------------------------------------------------------------ RandomState [D]
void philox(const uint32_t key[2], const uint32_t ctr[4], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int round = 0; round < 10; ++round) {
        const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c0;
        const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c2;
        const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = static_cast<uint32_t>(p1);
        const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = static_cast<uint32_t>(p0);
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

struct RandomState {
    uint32_t key[2];
    uint32_t ctr[4];
    uint32_t buf[4];
    int used = 4;
    RandomState(uint64_t seed, uint64_t history)
    {
        key[0] = static_cast<uint32_t>(seed);
        key[1] = static_cast<uint32_t>(seed >> 32);
        ctr[0] = static_cast<uint32_t>(history);
        ctr[1] = static_cast<uint32_t>(history >> 32);
        ctr[2] = 0;
        ctr[3] = 0;
    }
    // [D] the transport stream of a history starts at Philox block 2: blocks 0-1 belong to the source sampler
    // (the device samples sources warp-cooperatively with a separate generator object, transport.cu)
    void startTransportStream()
    {
        ctr[2] = 2;
        used = 4;
    }
    // four uniforms of the next Philox block (one block per event, see the protocol above)
    std::array<double, 4> block()
    {
        philox(key, ctr, buf);
        ++ctr[2];
        used = 4;
        return { static_cast<double>(buf[0] >> 8) * (1.0 / 16777216.0), static_cast<double>(buf[1] >> 8) * (1.0 / 16777216.0),
            static_cast<double>(buf[2] >> 8) * (1.0 / 16777216.0), static_cast<double>(buf[3] >> 8) * (1.0 / 16777216.0) };
    }
    // word `w` of block `b` of this history, as a uniform number (does not advance the stream)
    double wordOfBlock(uint32_t b, int w) const
    {
        uint32_t c[4] = { ctr[0], ctr[1], b, ctr[3] }, out[4];
        philox(key, c, out);
        return static_cast<double>(out[w] >> 8) * (1.0 / 16777216.0);
    }
    double randomUniform()
    {
        if (used == 4) {
            philox(key, ctr, buf);
            ++ctr[2];
            used = 0;
        }
        return static_cast<double>(buf[used++] >> 8) * (1.0 / 16777216.0);
    }
};

// ------------------------------------------------------------------ vectormath
using Vec = std::array<double, 3>;
Vec cross(const Vec& a, const Vec& b) { return { a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0] }; }
double dot(const Vec& a, const Vec& b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
double length(const Vec& a) { return std::sqrt(dot(a, a)); }
Vec normalized(Vec a)
{
    const double l = length(a);
    if (l > 0)
        for (auto& v : a)
            v /= l;
    return a;
}
Vec rotate(const Vec& v, const Vec& axisIn, double angle)
{
    const Vec k = normalized(axisIn);
    const double c = std::cos(angle), s = std::sin(angle);
    const Vec kv = cross(k, v);
    const double kd = dot(k, v) * (1 - c);
    return { v[0] * c + kv[0] * s + k[0] * kd, v[1] * c + kv[1] * s + k[1] * kd, v[2] * c + kv[2] * s + k[2] * kd };
}
int argmin3abs(const Vec& a)
{
    int m = 0;
    for (int i = 1; i < 3; ++i)
        if (std::fabs(a[i]) < std::fabs(a[m]))
            m = i;
    return m;
}
// dxmc::vectormath::peturb [R]
Vec peturb(const Vec& d, double cosTheta, double phi)
{
    const double sinTheta = std::sqrt(std::max(0.0, 1.0 - cosTheta * cosTheta));
    const double sinPhi = std::sin(phi), cosPhi = std::cos(phi);
    Vec n;
    if (std::fabs(d[2]) < 0.99999) {
        const double tmp = std::sqrt(1.0 - d[2] * d[2]);
        n[0] = d[0] * cosTheta + sinTheta * (d[0] * d[2] * cosPhi - d[1] * sinPhi) / tmp;
        n[1] = d[1] * cosTheta + sinTheta * (d[1] * d[2] * cosPhi + d[0] * sinPhi) / tmp;
        n[2] = d[2] * cosTheta - tmp * sinTheta * cosPhi;
    } else {
        n[0] = sinTheta * cosPhi;
        n[1] = sinTheta * sinPhi;
        n[2] = d[2] > 0 ? cosTheta : -cosTheta;
    }
    return normalized(n);
}

// ------------------------------------------------------------------ Material [D: table rule]
struct OMaterial {
    uint32_t nE = 0, nX = 0;
    double eMin = 1, xMin = 1;
    uint32_t ePerOctave = 64, xPerOctave = 32; // semi-log grids: node(i) = vmin 2^(i/P) (1 + (i%P)/P)
    std::vector<double> photo, incoh, coh, etr, ffCdf, sf;
    uint32_t nShells = 0;
    dxb_shell shells[DXB_MAX_SHELLS];
    double restFraction = 1, restJ0 = 0;

    void load(const dxb_material_tables& t)
    {
        nE = t.n_energy;
        nX = t.n_x;
        eMin = t.e_min_kev;
        xMin = t.x_min;
        ePerOctave = t.nodes_per_octave_e;
        xPerOctave = t.nodes_per_octave_x;
        photo.assign(t.photo, t.photo + nE);
        incoh.assign(t.incoh, t.incoh + nE);
        coh.assign(t.coh, t.coh + nE);
        etr.assign(t.etr, t.etr + nE);
        ffCdf.assign(t.ff_cdf, t.ff_cdf + nX);
        sf.assign(t.sf, t.sf + nX);
        nShells = t.n_shells;
        for (uint32_t i = 0; i < nShells; ++i)
            shells[i] = t.shells[i];
        restFraction = t.rest_electrons_fraction;
        restJ0 = t.rest_compton_j0;
    }
    static double lerpAt(const std::vector<double>& tab, double u)
    {
        const size_t n = tab.size();
        if (!(u > 0))
            return tab[0];
        size_t i = static_cast<size_t>(u);
        if (i >= n - 1)
            return tab[n - 1];
        const double f = u - static_cast<double>(i);
        return tab[i] + f * (tab[i + 1] - tab[i]);
    }
    static double semiLogCoord(double v, double vmin, uint32_t per)
    {
        const double r = v / vmin;
        if (!(r > 1.0))
            return 0.0;
        int ex = 0;
        const double m = 2.0 * std::frexp(r, &ex); // r = m 2^(ex-1), m in [1,2)
        return static_cast<double>(ex - 1) * per + (m - 1.0) * per;
    }
    double eCoord(double e) const { return semiLogCoord(e, eMin, ePerOctave); }
    double xCoord(double x) const { return semiLogCoord(x, xMin, xPerOctave); }
    double xNode(size_t k) const
    {
        return xMin * std::ldexp(1.0 + static_cast<double>(k % xPerOctave) / xPerOctave, static_cast<int>(k / xPerOctave));
    }

    struct AttenuationValues {
        double photoelectric, incoherent, coherent;
        double sum() const { return photoelectric + incoherent + coherent; }
    };
    AttenuationValues attenuationValues(double e) const
    {
        const double u = eCoord(e);
        return { lerpAt(photo, u), lerpAt(incoh, u), lerpAt(coh, u) };
    }
    double massEnergyTransfer(double e) const { return lerpAt(etr, eCoord(e)); }
    // S(x)/Z
    double scatterFactor(double x) const
    {
        if (x <= xMin) {
            const double r = x / xMin;
            return sf[0] * r * r;
        }
        return lerpAt(sf, xCoord(x));
    }
    // A(x^2) = int_0^{x^2} F^2 dt, piecewise linear in x^2 between nodes, F^2 flat below the first node
    double formFactorCumulative(double x) const
    {
        if (x <= xMin) {
            const double r = x / xMin;
            return ffCdf[0] * r * r;
        }
        const double u = xCoord(x);
        size_t i = static_cast<size_t>(u);
        if (i >= nX - 1)
            return ffCdf[nX - 1];
        const double xa = xNode(i), xb = xNode(i + 1);
        const double f = std::clamp((x * x - xa * xa) / (xb * xb - xa * xa), 0.0, 1.0);
        return ffCdf[i] + f * (ffCdf[i + 1] - ffCdf[i]);
    }
    // inverse of the above: x^2 for a given cumulative value
    double sampleSquaredMomentumTransfer(double target) const
    {
        if (target <= ffCdf[0])
            return target / ffCdf[0] * xMin * xMin;
        size_t lo = 0, hi = nX - 1;
        while (hi - lo > 1) {
            const size_t mid = (lo + hi) >> 1;
            if (ffCdf[mid] <= target)
                lo = mid;
            else
                hi = mid;
        }
        const double al = ffCdf[lo], ah = ffCdf[lo + 1];
        const double xa = xNode(lo), xb = xNode(lo + 1);
        const double f = ah > al ? (target - al) / (ah - al) : 0.0;
        return xa * xa + f * (xb * xb - xa * xa);
    }
};

// ------------------------------------------------------------------ Particle
struct Particle {
    Vec pos, dir;
    double energy, weight;
    void translate(double d)
    {
        for (int i = 0; i < 3; ++i)
            pos[i] += dir[i] * d;
    }
};

// ------------------------------------------------------------------ interactions [R]
struct InteractionResult {
    double energyImparted = 0;
    bool particleAlive = true;
    bool particleEnergyChanged = false;
    bool particleDirectionChanged = false;
};

// One Klein-Nishina candidate (r1) and rejection test (ra); true if accepted.
bool comptonTry(double energy, const OMaterial& material, int correction, double r1, double ra, double& e, double& cosTheta)
{
    const double k = energy / ELECTRON_REST_MASS;
    const double emin = 1.0 / (1.0 + 2.0 * k);
    const double gmaxInv = emin / (1.0 + emin * emin);
    e = r1 + (1.0 - r1) * emin;
    const double t = std::min((1.0 - e) / (k * e), 2.0);
    const double sinthetasqr = t * (2.0 - t);
    cosTheta = 1.0 - t;
    double g = (1.0 / e + e - sinthetasqr) * gmaxInv;
    if (correction >= 1) {
        // momentumTransferCosAngle(E, cos) = E/hc * sqrt((1-cos)/2)
        const double q = energy / HC * std::sqrt(0.5 * t);
        g *= material.scatterFactor(q);
    }
    return !(ra > g);
}

// One Rayleigh try; true if accepted.
bool rayleighTry(double energy, const OMaterial& material, int correction, double r0, double r1, double& cosAngle)
{
    if (correction == 0) {
        constexpr double extreme = 1.0886621079036347; // 4 sqrt2 / (3 sqrt3)
        const double rr = r0 * extreme;
        const double theta = PI * r1;
        const double sinang = std::sin(theta);
        cosAngle = std::cos(theta);
        return !(rr > (2.0 - sinang * sinang) * sinang);
    }
    const double qmax = energy / HC;
    const double qmax_squared = qmax * qmax;
    const double amax = material.formFactorCumulative(qmax);
    const double target = r0 * amax;
    const double q_squared = std::min(material.sampleSquaredMomentumTransfer(target), qmax_squared);
    cosAngle = 1.0 - 2.0 * q_squared / qmax_squared;
    return !((1.0 + cosAngle * cosAngle) * 0.5 < r1);
}

// ---- correction 2 (impulse approximation) [D: profile shape; R: intent] ----
// Shell by electron share (uncovered electrons: one unbound group with a common profile); pz from J(pz) = J0 sech^2(2 J0 pz), i.e.
// pz = ln(u/(1-u)) / (4 J0) atomic units; Doppler-broadened energy from energy-momentum conservation.  False: try
// rejected (shell cannot be ionised, or energy transfer below the binding energy).  The device reads the shell
// table as f32, so the same rounded values are used here.
constexpr double FINE_STRUCTURE = 7.2973525693e-3;
constexpr double U24 = 5.9604644775390625e-8;
bool dopplerBroaden(const OMaterial& material, double energy, double e0, double cosTheta, double rShell, double rPz, double& eOut)
{
    double r = rShell, U = 0, j0 = 0;
    bool bound = false;
    for (uint32_t i = 0; i < material.nShells; ++i) {
        const double f = mf(material.shells[i].n_electrons_fraction);
        if (r < f) {
            U = mf(material.shells[i].binding_energy_kev);
            j0 = mf(material.shells[i].compton_j0);
            bound = true;
            break;
        }
        r -= f;
    }
    eOut = e0;
    if (!bound) {
        j0 = mf(material.restJ0); // the electrons outside the table: unbound, one common profile
        if (!(j0 > 0))
            return true;
    }
    if (!(U < energy))
        return false;
    const double u = std::min(std::max(rPz, U24), 1.0 - U24);
    double pz = std::log(u / (1.0 - u)) * (0.25 / j0) * FINE_STRUCTURE;
    pz = std::min(std::max(pz, -0.5), 0.5);
    // E'/E = e0 / b (a + sign(pz) sqrt(a^2 - b (1 - t))), written with the expanded discriminant t (q - t e0^2 sin^2),
    // q = (1 - e0)^2 + 2 e0 (1 - cos), like the device (no cancellation of O(1) terms in f32)
    const double t = pz * pz;
    const double a = 1.0 - t * e0 * cosTheta;
    const double b = 1.0 - t * e0 * e0;
    const double q = (1.0 - e0) * (1.0 - e0) + 2.0 * e0 * (1.0 - cosTheta);
    const double disc = std::max(q - t * e0 * e0 * (1.0 - cosTheta * cosTheta), 0.0);
    const double e = e0 / b * (a + pz * std::sqrt(disc));
    if (!(e > 0) || !(energy - energy * e > U))
        return false;
    eOut = std::min(e, 1.0);
    return true;
}

// correction 2 photoelectric effect: energy of the emitted fluorescence photon (0: all absorbed locally)
double photoFluorescence(const OMaterial& material, double energy, double rShell, double rYield)
{
    double r = rShell;
    for (uint32_t i = 0; i < material.nShells; ++i) {
        const dxb_shell& s = material.shells[i];
        if (!(mf(s.binding_energy_kev) < energy))
            continue;
        const double f = mf(s.photo_fraction_above);
        if (r < f) {
            const double ef = mf(s.fluor_energy_kev);
            if (rYield < mf(s.fluor_yield) && ef >= MIN_ENERGY && ef < energy)
                return ef;
            return 0;
        }
        r -= f;
    }
    return 0;
}

double comptonScatter(Particle& p, const OMaterial& material, int correction, RandomState& state, const std::array<double, 4>* first,
    double* eRatio = nullptr, double* cosOut = nullptr)
{
    // the first try may share its block with the channel choice (words 1..3 of *first)
    std::array<double, 4> u = first ? *first : state.block();
    double e, cosTheta;
    for (;;) {
        bool ok = comptonTry(p.energy, material, correction, u[1], u[2], e, cosTheta);
        if (ok && correction >= 2) {
            const std::array<double, 4> ia = state.block(); // one extra block per accepted candidate
            ok = dopplerBroaden(material, p.energy, e, cosTheta, ia[0], ia[1], e);
        }
        if (ok)
            break;
        u = state.block();
    }
    p.dir = peturb(p.dir, cosTheta, 2.0 * PI * u[3]);
    const double E = p.energy;
    p.energy *= e;
    if (eRatio)
        *eRatio = e;
    if (cosOut)
        *cosOut = cosTheta;
    return (E - p.energy) * p.weight;
}

void rayleightScatter(Particle& p, const OMaterial& material, int correction, RandomState& state, double* cosOut = nullptr)
{
    std::array<double, 4> u = state.block();
    double cosAngle;
    while (!rayleighTry(p.energy, material, correction, u[0], u[1], cosAngle))
        u = state.block();
    p.dir = peturb(p.dir, cosAngle, 2.0 * PI * u[2]);
    if (cosOut)
        *cosOut = cosAngle;
}

InteractionResult interact(const OMaterial::AttenuationValues& att, Particle& p, const OMaterial& material, int correction,
    RandomState& state)
{
    InteractionResult res;
    const std::array<double, 4> u = state.block();
    const double r2 = u[0] * att.sum();
    if (r2 < att.photoelectric) {
        const double ef = correction >= 2 ? photoFluorescence(material, p.energy, u[1], u[2]) : 0.0;
        if (ef > 0) {
            // isotropic fluorescence photon, one extra block for its direction
            const std::array<double, 4> f = state.block();
            const double c = 2.0 * f[0] - 1.0, sn = std::sqrt(std::max(0.0, 1.0 - c * c)), phi = 2.0 * PI * f[1];
            p.dir = { sn * std::cos(phi), sn * std::sin(phi), c };
            res.energyImparted = (p.energy - ef) * p.weight;
            p.energy = ef;
            res.particleDirectionChanged = true;
        } else {
            res.energyImparted = p.energy * p.weight;
            p.energy = 0;
            res.particleAlive = false;
        }
        res.particleEnergyChanged = true;
    } else if (r2 < att.photoelectric + att.incoherent) {
        res.energyImparted = comptonScatter(p, material, correction, state, &u);
        res.particleEnergyChanged = true;
        res.particleDirectionChanged = true;
    } else {
        rayleightScatter(p, material, correction, state);
        res.particleDirectionChanged = true;
    }
    if (res.particleAlive) {
        if (p.energy < MIN_ENERGY) {
            res.energyImparted += p.energy * p.weight;
            p.energy = 0;
            res.particleAlive = false;
        } else if (p.weight < RUSSIAN_ROULETTE_THRESHOLD) {
            if (state.block()[0] < RUSSIAN_ROULETTE_PROBABILITY)
                res.particleAlive = false;
            else
                p.weight *= 1.0 / (1.0 - RUSSIAN_ROULETTE_PROBABILITY);
        }
    }
    return res;
}

// ------------------------------------------------------------------ EnergyScore [R]
inline void atomicAdd(double& target, double v)
{
    auto* a = reinterpret_cast<std::atomic<double>*>(&target);
    double old = a->load(std::memory_order_relaxed);
    while (!a->compare_exchange_weak(old, old + v, std::memory_order_relaxed)) { }
}
inline void atomicInc(uint64_t& target, uint64_t v = 1)
{
    reinterpret_cast<std::atomic<uint64_t>*>(&target)->fetch_add(v, std::memory_order_relaxed);
}

struct WorkerStats {
    uint64_t histories = 0, steps = 0, interactions = 0, deposits = 0, hops = 0;
    double emitted = 0;
};

// ------------------------------------------------------------------ AAVoxelGrid [R]
struct AAVoxelGrid {
    uint64_t dim[3];
    double spacing[3];
    double aabb[6]; // min xyz, max xyz
    std::vector<double> density;
    std::vector<uint8_t> materialIndex;
    std::vector<OMaterial> materials;
    std::vector<double> woodcockStepTable; // majorant on the energy grid
    // EnergyScore arrays
    std::vector<double> energyImparted, energyImpartedSquared;
    std::vector<uint64_t> nEvents;
    int scoreMaterial = -1; // calibration: collision kerma estimator in this material
    // [D] slab-local majorants (transport_pool.cu, LM builds): slabs of 2^lmShift voxel layers along z; inside slab s and
    // energy band b (= energy node index >> 5) the tracking majorant is majorant(E) * lmRatio[s * 16 + b]
    int lmShift = 0, lmSlabs = 0;
    std::vector<double> lmRatio;
    // [D] dense-box tracking (transport_pool.cu, DB builds): an axis-aligned box holds every voxel that is not "thin" (air);
    // inside it the tracking majorant is majorant(E), in the rest of the grid majorant(E) * boxRatio[band]
    bool boxOn = false;
    double boxLo[3] = { 0, 0, 0 }, boxHi[3] = { 0, 0, 0 }; // faces [cm] (the device's f32 values)
    double boxRatio[16] = { 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1 };

    size_t size() const { return density.size(); }
    double voxelVolume() const { return spacing[0] * spacing[1] * spacing[2]; }

    void build()
    {
        for (int i = 0; i < 3; ++i) {
            aabb[i] = -0.5 * dim[i] * spacing[i];
            aabb[i + 3] = 0.5 * dim[i] * spacing[i];
        }
        // per-material maximum density (as float, like the packed device voxels)
        std::vector<double> maxDens(materials.size(), 0.0);
        for (size_t i = 0; i < density.size(); ++i) {
            // the device grid stores the top 24 bits of the f32 density (round to nearest; DESIGN.md: voxel layout)
            double rho = density[i] > 0 ? density[i] : 0.0;
            if (g_mirror) {
                const float rf = static_cast<float>(rho);
                uint32_t bits;
                std::memcpy(&bits, &rf, 4);
                bits = (bits + 0x80u) & 0xFFFFFF00u;
                float rq;
                std::memcpy(&rq, &bits, 4);
                rho = static_cast<double>(rq);
            }
            density[i] = rho;
            maxDens[materialIndex[i]] = std::max(maxDens[materialIndex[i]], rho);
        }
        const uint32_t nE = materials[0].nE;
        woodcockStepTable.assign(nE, 0.0);
        for (uint32_t e = 0; e < nE; ++e) {
            double m = 0;
            for (size_t k = 0; k < materials.size(); ++k) {
                // the device majorant is built from the f32 total table
                const double tot = mf(materials[k].photo[e] + materials[k].incoh[e] + materials[k].coh[e]);
                m = std::max(m, g_mirror ? static_cast<double>(static_cast<float>(maxDens[k]) * static_cast<float>(tot)) : maxDens[k] * tot);
            }
            woodcockStepTable[e] = std::max(m, 1e-12);
        }
        clearEnergyScored();
    }
    void clearEnergyScored()
    {
        energyImparted.assign(size(), 0.0);
        energyImpartedSquared.assign(size(), 0.0);
        nEvents.assign(size(), 0);
    }
    double majorant(double e) const { return OMaterial::lerpAt(woodcockStepTable, materials[0].eCoord(e)); }

    void scoreEnergy(size_t idx, double e)
    {
        if (e <= 0)
            return;
        atomicAdd(energyImparted[idx], e);
        atomicAdd(energyImpartedSquared[idx], e * e);
        atomicInc(nEvents[idx]);
    }

    // distance to the exit of the AABB along dir from a point inside
    double exitDistance(const Particle& p) const
    {
        double t = 3.0e38;
        for (int i = 0; i < 3; ++i) {
            if (p.dir[i] != 0) {
                const double plane = p.dir[i] > 0 ? aabb[i + 3] : aabb[i];
                t = std::min(t, (plane - p.pos[i]) / p.dir[i]);
            }
        }
        return std::max(t, 0.0);
    }
    // entry/exit of a ray that may start outside; returns false on a miss
    bool intersect(const Particle& p, double& tmin, double& tmax) const
    {
        tmin = 0;
        tmax = 3.0e38;
        for (int i = 0; i < 3; ++i) {
            if (p.dir[i] == 0) {
                if (p.pos[i] < aabb[i] || p.pos[i] > aabb[i + 3])
                    return false;
            } else {
                const double t0 = (aabb[i] - p.pos[i]) / p.dir[i], t1 = (aabb[i + 3] - p.pos[i]) / p.dir[i];
                tmin = std::max(tmin, std::min(t0, t1));
                tmax = std::min(tmax, std::max(t0, t1));
            }
        }
        return tmax > tmin;
    }
    size_t flatIndex(const Vec& pos) const
    {
        size_t idx[3];
        for (int i = 0; i < 3; ++i) {
            long v = static_cast<long>((pos[i] - aabb[i]) / spacing[i]);
            v = std::min<long>(std::max<long>(v, 0), static_cast<long>(dim[i]) - 1);
            idx[i] = static_cast<size_t>(v);
        }
        return (idx[2] * dim[1] + idx[1]) * dim[0] + idx[0];
    }

    void woodcockTransport(Particle& p, int correction, RandomState& state, WorkerStats& st)
    {
        bool still_inside = true;
        bool updateAtt = true;
        double attMax = 1, attMaxInv = 1;
        while (still_inside) {
            if (updateAtt) {
                attMax = majorant(p.energy);
                attMaxInv = 1.0 / attMax;
                updateAtt = false;
            }
            // one Philox block serves a pair of tentative steps; the second half is discarded if the first step
            // was a real interaction or left the grid [D]
            const std::array<double, 4> u = state.block();
            for (int half = 0; half < 2 && still_inside; ++half) {
                const double steplen = -std::log(1.0 - u[2 * half]) * attMaxInv;
                const double toExit = exitDistance(p);
                if (!(steplen < toExit)) {
                    still_inside = false;
                    break;
                }
                ++st.steps;
                p.translate(steplen);
                const size_t flat = flatIndex(p.pos);
                const uint8_t matInd = materialIndex[flat];
                const double dens = density[flat];
                const OMaterial& mat = materials[matInd];
                const auto att = mat.attenuationValues(p.energy);
                const double attSum = att.sum() * dens;
                if (scoreMaterial >= 0 && matInd == scoreMaterial) {
                    // [D] collision estimator of air kerma: each tentative collision stands for 1/mu_max of track
                    scoreEnergy(flat, p.weight * p.energy * mat.massEnergyTransfer(p.energy) * attMaxInv);
                }
                if (u[2 * half + 1] * attMax < attSum) {
                    ++st.interactions;
                    const auto res = interact(att, p, mat, correction, state);
                    if (scoreMaterial < 0 && res.energyImparted > 0) {
                        ++st.deposits;
                        scoreEnergy(flat, res.energyImparted);
                    }
                    still_inside = res.particleAlive;
                    updateAtt = res.particleEnergyChanged;
                    break;
                }
            }
        }
    }

    long layerOf(double z) const { return static_cast<long>(std::floor((z - aabb[2]) / spacing[2])); }

    // [D] Woodcock tracking with slab-local majorants.  Same random-number protocol as woodcockTransport (one Philox block
    // per pair of tentative steps).  The optical depth tau = -ln(1 - u) drawn for a tentative step is marched through the
    // slabs: while it exceeds what the rest of the current slab absorbs at the slab's majorant, the ray moves to the slab
    // face and tau shrinks accordingly; the tentative collision lies where tau is used up and is accepted with
    // mu(x) / (slab majorant).  Unbiased for any table that bounds mu inside each slab; crossing a face consumes no
    // random number.
    void woodcockTransportSlabs(Particle& p, int correction, RandomState& state, WorkerStats& st)
    {
        const double thickness = static_cast<double>(1 << lmShift) * spacing[2];
        bool still_inside = true;
        while (still_inside) {
            const double attMax = majorant(p.energy);
            const size_t node = std::min<size_t>(static_cast<size_t>(std::max(0.0, materials[0].eCoord(p.energy))), materials[0].nE - 2);
            const int band = static_cast<int>(node >> 5);
            long slab = std::min<long>(std::max<long>(layerOf(p.pos[2]), 0), static_cast<long>(dim[2]) - 1) >> lmShift;
            const std::array<double, 4> u = state.block();
            for (int half = 0; half < 2 && still_inside; ++half) {
                double tau = -std::log(1.0 - u[2 * half]);
                double ratio = 1.0;
                const bool up = p.dir[2] > 0;
                for (;;) {
                    ratio = lmRatio[static_cast<size_t>(slab) * 16 + band];
                    const double mus = attMax * ratio;
                    const double zf = aabb[2] + static_cast<double>(slab + (up ? 1 : 0)) * thickness;
                    const double tb = p.dir[2] != 0 ? (zf - p.pos[2]) / p.dir[2] : 3.0e38;
                    const double need = tau / mus;
                    if (need < tb) {
                        if (!(need < exitDistance(p)))
                            still_inside = false;
                        else
                            p.translate(need);
                        break;
                    }
                    tau = std::max(tau - tb * mus, 0.0);
                    if (!(tb < exitDistance(p))) {
                        still_inside = false; // leaves through the side (or the top / bottom) before the face
                        break;
                    }
                    p.translate(tb);
                    p.pos[2] = zf;
                    slab += up ? 1 : -1;
                    ++st.hops;
                    if (slab < 0 || slab >= lmSlabs) {
                        still_inside = false;
                        break;
                    }
                }
                if (!still_inside)
                    break;
                ++st.steps;
                const size_t flat = flatIndex(p.pos);
                const uint8_t matInd = materialIndex[flat];
                const OMaterial& mat = materials[matInd];
                const auto att = mat.attenuationValues(p.energy);
                const double attSum = att.sum() * density[flat];
                if (u[2 * half + 1] * attMax * ratio < attSum) {
                    ++st.interactions;
                    const auto res = interact(att, p, mat, correction, state);
                    if (res.energyImparted > 0) {
                        ++st.deposits;
                        scoreEnergy(flat, res.energyImparted);
                    }
                    still_inside = res.particleAlive;
                    break; // the energy (band, majorant) or the direction may have changed: new block
                }
            }
        }
    }

    // distance to the exit of the dense box along dir from a point inside it
    double boxExitDistance(const Particle& p) const
    {
        double t = 3.0e38;
        for (int i = 0; i < 3; ++i) {
            if (p.dir[i] != 0) {
                const double plane = p.dir[i] > 0 ? boxHi[i] : boxLo[i];
                t = std::min(t, (plane - p.pos[i]) / p.dir[i]);
            }
        }
        return std::max(t, 0.0);
    }
    // distance at which the ray enters the dense box (0 if it starts inside); false on a miss
    bool boxIntersect(const Particle& p, double& tmin) const
    {
        tmin = 0;
        double tmax = 3.0e38;
        for (int i = 0; i < 3; ++i) {
            if (p.dir[i] == 0) {
                if (p.pos[i] < boxLo[i] || p.pos[i] > boxHi[i])
                    return false;
            } else {
                const double t0 = (boxLo[i] - p.pos[i]) / p.dir[i], t1 = (boxHi[i] - p.pos[i]) / p.dir[i];
                tmin = std::max(tmin, std::min(t0, t1));
                tmax = std::min(tmax, std::max(t0, t1));
            }
        }
        return tmax > tmin;
    }

    // [D] Woodcock tracking with a dense box.  Outside the box (air around the patient) the majorant is majorant(E) *
    // boxRatio[band], inside it majorant(E).  A region boundary is crossed without a collision, so by the memoryless
    // property the flight simply restarts there with a fresh optical depth:
    //   * the first flight, from the grid's face to the box, draws its optical depth from the spare word 1 of the source's
    //     second Philox block (block 1) and consumes no block of the transport stream;
    //   * outside ("flight", one Philox block per flight): tau = -ln(1 - u0); the ray either has its tentative collision
    //     in the outside region (accepted with mu / (majorant * ratio) by u1 of the NEXT block, after which u2 of that block
    //     starts the next flight), or reaches the box and continues inside, or leaves the grid;
    //   * inside: pairs of tentative steps at the global majorant exactly as woodcockTransport; a tentative step that
    //     ends beyond the box face leaves the box, and the acceptance number of that step (unused otherwise) is the
    //     optical depth of the flight from the face to the grid's boundary.
    void woodcockTransportBox(Particle& p, int correction, RandomState& state, WorkerStats& st)
    {
        bool out = true;  // the history starts on the grid's face: fly to the box first
        bool air = false; // a tentative collision in the outside region waits for its acceptance number
        bool alive = true;
        bool first = true;
        while (alive) {
            const double attMax = majorant(p.energy);
            const size_t node = std::min<size_t>(static_cast<size_t>(std::max(0.0, materials[0].eCoord(p.energy))), materials[0].nE - 2);
            const double attOut = g_mirror ? static_cast<double>(static_cast<float>(attMax) * static_cast<float>(boxRatio[node >> 5]))
                                           : attMax * boxRatio[node >> 5];
            std::array<double, 4> u { state.wordOfBlock(1, 1), 0, 0, 0 };
            if (!first)
                u = state.block();
            first = false;
            if (out) {
                double tau;
                if (air) {
                    ++st.steps;
                    const size_t flat = flatIndex(p.pos);
                    const uint8_t matInd = materialIndex[flat];
                    const OMaterial& mat = materials[matInd];
                    const auto att = mat.attenuationValues(p.energy);
                    const double attSum = att.sum() * density[flat];
                    air = false;
                    if (u[1] * attOut < attSum) {
                        ++st.interactions;
                        const auto res = interact(att, p, mat, correction, state);
                        if (res.energyImparted > 0) {
                            ++st.deposits;
                            scoreEnergy(flat, res.energyImparted);
                        }
                        alive = res.particleAlive;
                        continue; // still outside the box: the next block starts a flight
                    }
                    tau = -std::log(1.0 - u[2]);
                } else {
                    tau = -std::log(1.0 - u[0]);
                }
                double tIn = 0;
                const bool hit = boxIntersect(p, tIn);
                const double need = tau / attOut;
                if (need < (hit ? tIn : exitDistance(p))) {
                    p.translate(need);
                    air = true;
                } else if (hit) {
                    p.translate(tIn);
                    for (int i = 0; i < 3; ++i)
                        p.pos[i] = std::min(std::max(p.pos[i], boxLo[i]), boxHi[i]);
                    out = false;
                    ++st.hops;
                } else {
                    alive = false;
                }
                continue;
            }
            for (int half = 0; half < 2; ++half) {
                const double steplen = -std::log(1.0 - u[2 * half]) / attMax;
                const double tB = boxExitDistance(p);
                if (!(steplen < tB)) {
                    const double need = -std::log(1.0 - u[2 * half + 1]) / attOut;
                    if (need < exitDistance(p) - tB) {
                        p.translate(tB + need);
                        out = true;
                        air = true;
                    } else {
                        alive = false;
                    }
                    break;
                }
                ++st.steps;
                p.translate(steplen);
                const size_t flat = flatIndex(p.pos);
                const uint8_t matInd = materialIndex[flat];
                const OMaterial& mat = materials[matInd];
                const auto att = mat.attenuationValues(p.energy);
                const double attSum = att.sum() * density[flat];
                if (u[2 * half + 1] * attMax < attSum) {
                    ++st.interactions;
                    const auto res = interact(att, p, mat, correction, state);
                    if (res.energyImparted > 0) {
                        ++st.deposits;
                        scoreEnergy(flat, res.energyImparted);
                    }
                    alive = res.particleAlive;
                    break;
                }
            }
        }
    }

    // World::transport: move to the AABB, then track
    void transport(Particle& p, int correction, RandomState& state, WorkerStats& st)
    {
        double tmin, tmax;
        if (p.energy < MIN_ENERGY || !intersect(p, tmin, tmax))
            return;
        p.translate(tmin);
        if (lmSlabs >= 2 && scoreMaterial < 0)
            woodcockTransportSlabs(p, correction, state, st);
        else if (boxOn && scoreMaterial < 0)
            woodcockTransportBox(p, correction, state, st);
        else
            woodcockTransport(p, correction, state, st);
    }
};

// ------------------------------------------------------------------ filters [D normalisation]
struct Bowtie {
    std::vector<double> angle, weight;
    explicit Bowtie(const dxb_bowtie& b)
    {
        if (b.n < 2 || !b.angle_rad || !b.weight)
            return;
        std::vector<std::pair<double, double>> d;
        for (uint32_t i = 0; i < b.n; ++i)
            d.emplace_back(std::fabs(b.angle_rad[i]), b.weight[i]);
        std::sort(d.begin(), d.end());
        for (const auto& p : d) {
            if (!angle.empty() && p.first - angle.back() < 1e-12)
                weight.back() = 0.5 * (weight.back() + p.second);
            else {
                angle.push_back(p.first);
                weight.push_back(std::max(0.0, p.second));
            }
        }
        if (angle.size() < 2) {
            angle.clear();
            weight.clear();
            return;
        }
        double area = weight.front() * angle.front();
        for (size_t i = 1; i < angle.size(); ++i)
            area += 0.5 * (weight[i] + weight[i - 1]) * (angle[i] - angle[i - 1]);
        const double mean = area / angle.back();
        if (mean > 0)
            for (auto& w : weight)
                w /= mean;
        // the device evaluates the profile from f32 knots
        for (size_t i = 0; i < angle.size(); ++i) {
            angle[i] = mf(angle[i]);
            weight[i] = mf(weight[i]);
        }
    }
    double operator()(double a) const
    {
        if (angle.size() < 2)
            return 1.0;
        a = std::fabs(a);
        if (a <= angle.front())
            return weight.front();
        if (a >= angle.back())
            return weight.back();
        size_t i = 1;
        while (angle[i] < a)
            ++i;
        return weight[i - 1] + (a - angle[i - 1]) / (angle[i] - angle[i - 1]) * (weight[i] - weight[i - 1]);
    }
};

struct AEC {
    std::vector<double> w;
    Vec start { 0, 0, 0 }, dir { 0, 0, 1 };
    double len = 0;
    explicit AEC(const dxb_aec& a)
    {
        if (a.n < 2 || !a.weights)
            return;
        Vec d { a.stop[0] - a.start[0], a.stop[1] - a.start[1], a.stop[2] - a.start[2] };
        len = length(d);
        if (!(len > 0))
            return;
        start = { a.start[0], a.start[1], a.start[2] };
        dir = normalized(d);
        w.assign(a.weights, a.weights + a.n);
        double mean = 0;
        for (double v : w)
            mean += v;
        mean /= w.size();
        if (mean > 0)
            for (auto& v : w)
                v /= mean;
    }
    double operator()(const Vec& pos) const
    {
        if (w.size() < 2 || !(len > 0))
            return 1.0;
        const Vec rel { pos[0] - start[0], pos[1] - start[1], pos[2] - start[2] };
        double u = std::clamp(dot(rel, dir) / len, 0.0, 1.0) * (w.size() - 1);
        const size_t i = static_cast<size_t>(u);
        if (i >= w.size() - 1)
            return w.back();
        return w[i] + (u - i) * (w[i + 1] - w[i]);
    }
};

double wrap2pi(double a)
{
    a = std::fmod(a, 2 * PI);
    return a < 0 ? a + 2 * PI : a;
}
double organAecHigh(const dxb_organ_aec& o)
{
    if (!o.use_filter || !o.compensate_outside)
        return 1.0;
    const double low = std::clamp(o.low_weight, 0.0, 1.0);
    const double L = wrap2pi(o.stop_angle - o.start_angle);
    const double r = std::min(std::max(0.0, o.ramp_angle), 0.5 * (2 * PI - L));
    const double den = 2 * PI - L - r;
    return den > 1e-9 ? (2 * PI - low * (L + r)) / den : 1.0;
}
double organAec(const dxb_organ_aec& o, double angle)
{
    if (!o.use_filter)
        return 1.0;
    const double low = std::clamp(o.low_weight, 0.0, 1.0), high = organAecHigh(o);
    const double L = wrap2pi(o.stop_angle - o.start_angle);
    const double r = std::min(std::max(0.0, o.ramp_angle), 0.5 * (2 * PI - L));
    const double a = wrap2pi(angle - o.start_angle);
    if (a <= L)
        return low;
    if (r > 0 && a < L + r)
        return low + (high - low) * (a - L) / r;
    if (r > 0 && a > 2 * PI - r)
        return low + (high - low) * (2 * PI - a) / r;
    return high;
}

// ------------------------------------------------------------------ beams [V geometry formulas / R gantry]
uint64_t ceilCount(double x)
{
    if (!(x > 0) || !std::isfinite(x))
        return 1;
    return std::max<uint64_t>(1, static_cast<uint64_t>(std::ceil(x - 1e-9)));
}
double stepOf(const dxb_beam_desc& b) { return std::fabs(b.step_angle) > 0 ? std::fabs(b.step_angle) : PI / 180.0; }
bool dualTube(const dxb_beam_desc& b)
{
    return b.type == DXB_BEAM_CT_SPIRAL_DUAL || (b.type == DXB_BEAM_CTDI && b.spectrum[1].n > 0);
}

uint64_t numberOfExposures(const dxb_beam_desc& b)
{
    switch (b.type) {
    case DXB_BEAM_DX:
    case DXB_BEAM_PENCIL:
        return std::max<uint64_t>(1, b.n_exposures);
    case DXB_BEAM_CT_SPIRAL:
    case DXB_BEAM_CT_SPIRAL_DUAL: {
        const Vec d { b.stop[0] - b.start[0], b.stop[1] - b.start[1], b.stop[2] - b.start[2] };
        const double feed = std::fabs(b.pitch * b.collimation);
        const double total = feed > 0 ? length(d) / feed * 2 * PI : 2 * PI;
        const uint64_t n = ceilCount(total / stepOf(b));
        return b.type == DXB_BEAM_CT_SPIRAL_DUAL ? 2 * n : n;
    }
    case DXB_BEAM_CBCT:
        return ceilCount(std::fabs(b.stop_angle - b.start_angle) / stepOf(b));
    case DXB_BEAM_CT_SEQUENTIAL:
        return std::max<uint64_t>(1, b.n_slices) * ceilCount(2 * PI / stepOf(b));
    case DXB_BEAM_CTDI:
        return ceilCount(2 * PI / stepOf(b)) * (dualTube(b) ? 2 : 1);
    }
    return 0;
}

struct Exposure {
    Vec pos, dirCosX, dirCosY, dir;
    double halfAngles[2] = { 0, 0 };
    double weight = 1;
    int tube = 0;
};

void gantry(const Vec& center, Vec axis, double angle, double sdd, Exposure& e)
{
    axis = normalized(axis);
    Vec unit { 0, 0, 0 };
    unit[argmin3abs(axis)] = 1;
    const Vec n0 = normalized(cross(unit, axis));
    e.dirCosX = rotate(n0, axis, angle);
    e.dirCosY = axis;
    e.dir = cross(e.dirCosX, e.dirCosY);
    for (int i = 0; i < 3; ++i)
        e.pos[i] = center[i] - e.dir[i] * sdd * 0.5;
}

void tubeRelativeWeights(const dxb_beam_desc& b, double& wa, double& wb)
{
    auto total = [](const dxb_spectrum& s) {
        double t = 0;
        for (uint32_t i = 0; i < s.n; ++i)
            t += s.weight ? s.weight[i] : 0.0;
        return t;
    };
    const double ma = b.relative_mas_a > 0 ? b.relative_mas_a : 1.0, mb = b.relative_mas_b > 0 ? b.relative_mas_b : 1.0;
    double a = ma * total(b.spectrum[0]), c = mb * total(b.spectrum[1]);
    if (!(a > 0) || !(c > 0)) {
        a = ma;
        c = mb;
    }
    wa = 2 * a / (a + c);
    wb = 2 * c / (a + c);
}

bool exposureOf(const dxb_beam_desc& b, uint64_t index, const AEC& aec, Exposure& e)
{
    if (index >= numberOfExposures(b))
        return false;
    const double step = stepOf(b);
    e = Exposure {};
    switch (b.type) {
    case DXB_BEAM_DX: {
        e.pos = { b.position[0], b.position[1], b.position[2] };
        e.dirCosX = normalized({ b.cosines[0][0], b.cosines[0][1], b.cosines[0][2] });
        e.dirCosY = normalized({ b.cosines[1][0], b.cosines[1][1], b.cosines[1][2] });
        e.dir = cross(e.dirCosX, e.dirCosY);
        e.halfAngles[0] = std::fabs(b.half_angles[0]);
        e.halfAngles[1] = std::fabs(b.half_angles[1]);
        return true;
    }
    case DXB_BEAM_PENCIL: {
        Vec d { b.direction[0], b.direction[1], b.direction[2] };
        if (!(length(d) > 0))
            d = { 0, 0, 1 };
        d = normalized(d);
        Vec unit { 0, 0, 0 };
        unit[argmin3abs(d)] = 1;
        e.dirCosX = normalized(cross(unit, d));
        e.dirCosY = cross(d, e.dirCosX);
        e.dir = cross(e.dirCosX, e.dirCosY);
        e.pos = { b.position[0], b.position[1], b.position[2] };
        return true;
    }
    case DXB_BEAM_CT_SPIRAL:
    case DXB_BEAM_CT_SPIRAL_DUAL: {
        const bool dual = b.type == DXB_BEAM_CT_SPIRAL_DUAL;
        const uint64_t k = dual ? index / 2 : index;
        const int tube = dual ? static_cast<int>(index & 1) : 0;
        Vec axis { b.stop[0] - b.start[0], b.stop[1] - b.start[1], b.stop[2] - b.start[2] };
        if (!(length(axis) > 0))
            axis = { 0, 0, 1 };
        axis = normalized(axis);
        const double rot = static_cast<double>(k) * step;
        const double feed = b.pitch * b.collimation * rot / (2 * PI);
        const Vec center { b.start[0] + axis[0] * feed, b.start[1] + axis[1] * feed, b.start[2] + axis[2] * feed };
        const double angle = b.start_angle + rot + (tube ? b.tube_b_offset_angle : 0.0);
        gantry(center, axis, angle, b.sdd, e);
        e.halfAngles[0] = std::atan((tube ? b.fov_b : b.fov) / b.sdd);
        e.halfAngles[1] = std::atan(b.collimation / b.sdd);
        e.tube = tube;
        e.weight = aec(center) * organAec(b.organ_aec, angle);
        if (dual) {
            double wa, wb;
            tubeRelativeWeights(b, wa, wb);
            e.weight *= tube ? wb : wa;
        }
        return true;
    }
    case DXB_BEAM_CT_SEQUENTIAL:
    case DXB_BEAM_CTDI: {
        const bool dual = dualTube(b);
        const uint64_t perRot = ceilCount(2 * PI / step);
        const uint64_t k = dual ? index / 2 : index;
        const int tube = dual ? static_cast<int>(index & 1) : 0;
        const uint64_t slice = k / perRot, a = k % perRot;
        Vec axis { b.direction[0], b.direction[1], b.direction[2] };
        if (!(length(axis) > 0))
            axis = { 0, 0, 1 };
        axis = normalized(axis);
        const double off = b.slice_spacing * static_cast<double>(slice);
        const Vec center { b.position[0] + axis[0] * off, b.position[1] + axis[1] * off, b.position[2] + axis[2] * off };
        const double angle = b.start_angle + static_cast<double>(a) * step + (tube ? b.tube_b_offset_angle : 0.0);
        gantry(center, axis, angle, b.sdd, e);
        e.halfAngles[0] = std::atan((tube ? b.fov_b : b.fov) / b.sdd);
        e.halfAngles[1] = std::atan(b.collimation / b.sdd);
        e.tube = tube;
        e.weight = b.type == DXB_BEAM_CTDI ? 1.0 : organAec(b.organ_aec, angle);
        if (dual) {
            double wa, wb;
            tubeRelativeWeights(b, wa, wb);
            e.weight *= tube ? wb : wa;
        }
        return true;
    }
    case DXB_BEAM_CBCT: {
        const double sign = b.stop_angle >= b.start_angle ? 1.0 : -1.0;
        const double angle = b.start_angle + sign * static_cast<double>(index) * step;
        Vec axis { b.direction[0], b.direction[1], b.direction[2] };
        if (!(length(axis) > 0))
            axis = { 0, 0, 1 };
        gantry({ b.isocenter[0], b.isocenter[1], b.isocenter[2] }, axis, angle, b.sdd, e);
        e.halfAngles[0] = std::fabs(b.half_angles[0]);
        e.halfAngles[1] = std::fabs(b.half_angles[1]);
        return true;
    }
    }
    return false;
}

// SpecterDistribution: Walker alias table (Vose's construction) [R]
struct SpecterDistribution {
    std::vector<double> energies, prob;
    std::vector<size_t> alias;
    double step = 0;
    void setup(const dxb_spectrum& s)
    {
        const size_t n = s.n;
        energies.assign(s.energy_kev, s.energy_kev + n);
        step = n > 1 ? energies[1] - energies[0] : 0.0;
        prob.assign(n, 1.0);
        alias.resize(n);
        double sum = 0;
        for (size_t i = 0; i < n; ++i)
            sum += std::max(0.0, s.weight[i]);
        std::vector<double> q(n);
        std::vector<size_t> small, large;
        for (size_t i = 0; i < n; ++i) {
            q[i] = sum > 0 ? std::max(0.0, s.weight[i]) / sum * static_cast<double>(n) : 1.0;
            alias[i] = i;
        }
        for (size_t i = 0; i < n; ++i)
            (q[i] < 1.0 ? small : large).push_back(i);
        while (!small.empty() && !large.empty()) {
            const size_t sm = small.back();
            small.pop_back();
            const size_t lg = large.back();
            large.pop_back();
            prob[sm] = mf(q[sm]); // device keeps f32 acceptance values
            alias[sm] = lg;
            q[lg] = (q[lg] + q[sm]) - 1.0;
            (q[lg] < 1.0 ? small : large).push_back(lg);
        }
    }
    double sampleValue(RandomState& state) const
    {
        const size_t n = energies.size();
        if (n <= 1)
            return energies[0];
        const double r0 = state.randomUniform();
        size_t idx = std::min(static_cast<size_t>(r0 * static_cast<double>(n)), n - 1);
        const double r1 = state.randomUniform();
        if (!(r1 < prob[idx]))
            idx = alias[idx];
        const double r2 = state.randomUniform();
        double e = energies[0] + static_cast<double>(idx) * step;
        if (idx < n - 1)
            e += r2 * step;
        return e;
    }
};

struct BeamModel {
    const dxb_beam_desc& b;
    AEC aec;
    Bowtie bowtie[2];
    SpecterDistribution specter[2];
    bool pencil;
    bool useBowtie;
    explicit BeamModel(const dxb_beam_desc& beam)
        : b(beam)
        , aec(beam.aec)
        , bowtie { Bowtie(beam.bowtie[0]), Bowtie(dualTube(beam) ? beam.bowtie[1] : beam.bowtie[0]) }
    {
        pencil = b.type == DXB_BEAM_PENCIL;
        useBowtie = !(b.type == DXB_BEAM_PENCIL || b.type == DXB_BEAM_DX || b.type == DXB_BEAM_CBCT);
        if (!pencil) {
            specter[0].setup(b.spectrum[0]);
            specter[1].setup(dualTube(b) ? b.spectrum[1] : b.spectrum[0]);
        }
    }
    // Exposure::sampleParticle [R], direction formula [V] R:src/libopendxmc/beamactorcontainer.cpp:44-69
    Particle sampleParticle(const Exposure& e, RandomState& state) const
    {
        // exposure geometry is held in f32 on the device; mirror that rounding so that both sides start
        // from the same rays
        auto f = [](double v) { return mf(v); };
        const double angx = (2.0 * state.randomUniform() - 1.0) * f(e.halfAngles[0]);
        const double angy = (2.0 * state.randomUniform() - 1.0) * f(e.halfAngles[1]);
        Particle p;
        p.energy = pencil ? f(b.energy) : specter[e.tube].sampleValue(state);
        p.weight = f(e.weight) * (useBowtie ? bowtie[e.tube](angx) : 1.0);
        const double sinx = std::sin(angx), siny = std::sin(angy);
        const double sinz = std::sqrt(std::max(0.0, 1.0 - sinx * sinx - siny * siny));
        for (int i = 0; i < 3; ++i) {
            p.dir[i] = f(e.dirCosX[i]) * sinx + f(e.dirCosY[i]) * siny + f(e.dir[i]) * sinz;
            p.pos[i] = f(e.pos[i]);
        }
        return p;
    }
};

} // namespace

// =========================================================================== World
struct orc_world {
    AAVoxelGrid grid;
    OMaterial air, pmma;
    double airDensity = 1.20479e-3, pmmaDensity = 1.19;
    bool haveReference = false;
};

namespace {

// Transport::runWorker [R]
void runWorker(AAVoxelGrid& grid, const BeamModel& beam, int correction, uint64_t seed, uint64_t nExposures, uint64_t ppe,
    uint64_t rank, uint64_t world, std::atomic<uint64_t>& idx, WorkerStats& st)
{
    uint64_t n;
    while ((n = idx.fetch_add(1)) < nExposures) {
        Exposure exposure;
        exposureOf(beam.b, n, beam.aec, exposure);
        for (uint64_t i = 0; i < ppe; ++i) {
            const uint64_t history = n * ppe + i;
            if (world > 1 && (history / SHARD_BLOCK) % world != rank)
                continue;
            RandomState state(seed, history);
            Particle p = beam.sampleParticle(exposure, state);
            state.startTransportStream();
            ++st.histories;
            st.emitted += p.energy * p.weight;
            grid.transport(p, correction, state, st);
        }
    }
}

void runBeam(AAVoxelGrid& grid, const dxb_beam_desc& b, int correction, uint64_t seed, int nThreads, uint64_t rank, uint64_t world,
    orc_stats* stats)
{
    const auto t0 = std::chrono::steady_clock::now();
    BeamModel beam(b);
    const uint64_t nExp = numberOfExposures(b);
    if (nThreads <= 0)
        nThreads = std::max(1u, std::thread::hardware_concurrency());
    std::atomic<uint64_t> idx { 0 };
    std::vector<WorkerStats> ws(nThreads);
    std::vector<std::thread> threads;
    for (int t = 1; t < nThreads; ++t)
        threads.emplace_back(runWorker, std::ref(grid), std::cref(beam), correction, seed, nExp, b.particles_per_exposure, rank, world,
            std::ref(idx), std::ref(ws[t]));
    runWorker(grid, beam, correction, seed, nExp, b.particles_per_exposure, rank, world, idx, ws[0]);
    for (auto& t : threads)
        t.join();
    if (stats) {
        *stats = orc_stats {};
        for (const auto& w : ws) {
            stats->histories += w.histories;
            stats->steps += w.steps;
            stats->hops += w.hops;
            stats->interactions += w.interactions;
            stats->deposits += w.deposits;
            stats->energy_emitted_kev += w.emitted;
        }
        for (double e : grid.energyImparted)
            stats->energy_deposited_kev += e;
        stats->threads = nThreads;
        stats->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
}

bool isCt(int type) { return type == DXB_BEAM_CT_SPIRAL || type == DXB_BEAM_CT_SPIRAL_DUAL || type == DXB_BEAM_CT_SEQUENTIAL; }

// CT*Beam::calibrationFactor [R intent, D phantom]: one axial rotation on a PMMA CTDI cylinder.
double ctCalibration(const orc_world& w, const dxb_beam_desc& b, int correction, uint64_t seed, uint64_t calibHistories, int nThreads)
{
    const double diameter = b.ctdi_diameter > 0 ? b.ctdi_diameter : 32.0;
    AAVoxelGrid ph;
    const double sxy = 0.2, sz = 0.5;
    const int nxy = static_cast<int>(std::ceil((diameter + 4.0) / sxy / 2.0)) * 2, nz = 30;
    ph.dim[0] = ph.dim[1] = nxy;
    ph.dim[2] = nz;
    ph.spacing[0] = ph.spacing[1] = sxy;
    ph.spacing[2] = sz;
    const size_t n = static_cast<size_t>(nxy) * nxy * nz;
    ph.density.assign(n, w.airDensity);
    ph.materialIndex.assign(n, 0);
    std::vector<int> hole(n, -1);
    const double R = 0.5 * diameter, rh = 0.655, off = R - 1.0;
    const double hx[5] = { 0, off, -off, 0, 0 }, hy[5] = { 0, 0, 0, off, -off };
    for (int k = 0; k < nz; ++k) {
        const double z = (k + 0.5) * sz - 0.5 * nz * sz;
        for (int j = 0; j < nxy; ++j) {
            const double y = (j + 0.5) * sxy - 0.5 * nxy * sxy;
            for (int i = 0; i < nxy; ++i) {
                const double x = (i + 0.5) * sxy - 0.5 * nxy * sxy;
                const size_t idx = (static_cast<size_t>(k) * nxy + j) * nxy + i;
                if (x * x + y * y <= R * R) {
                    ph.density[idx] = w.pmmaDensity;
                    ph.materialIndex[idx] = 1;
                    for (int h = 0; h < 5; ++h) {
                        const double ddx = x - hx[h], ddy = y - hy[h];
                        if (ddx * ddx + ddy * ddy <= rh * rh) {
                            ph.density[idx] = w.airDensity;
                            ph.materialIndex[idx] = 0;
                            if (std::fabs(z) <= 5.0) {
                                ph.materialIndex[idx] = 2;
                                hole[idx] = h;
                            }
                        }
                    }
                }
            }
        }
    }
    ph.materials = { w.air, w.pmma, w.air };
    ph.scoreMaterial = 2;
    ph.build();

    dxb_beam_desc cb = b;
    cb.type = DXB_BEAM_CTDI;
    cb.position[0] = cb.position[1] = cb.position[2] = 0;
    cb.direction[0] = cb.direction[1] = 0;
    cb.direction[2] = 1;
    cb.step_angle = PI / 180.0;
    cb.start_angle = 0;
    cb.n_slices = 1;
    cb.aec.n = 0;
    cb.organ_aec.use_filter = 0;
    if (b.type != DXB_BEAM_CT_SPIRAL_DUAL)
        cb.spectrum[1].n = 0;
    const uint64_t nExp = numberOfExposures(cb);
    cb.particles_per_exposure = std::max<uint64_t>(1, calibHistories / nExp);
    // the nested run draws from its own stream: beam key ^ DXB_CALIBRATION_KEY_XOR (include/dxb.h)
    runBeam(ph, cb, correction, seed ^ DXB_CALIBRATION_KEY_XOR, nThreads, 0, 1, nullptr);

    double sum[5] = { 0, 0, 0, 0, 0 };
    size_t cnt[5] = { 0, 0, 0, 0, 0 };
    for (size_t i = 0; i < n; ++i)
        if (hole[i] >= 0) {
            sum[hole[i]] += ph.energyImparted[i];
            ++cnt[hole[i]];
        }
    double kerma[5];
    for (int h = 0; h < 5; ++h)
        kerma[h] = cnt[h] ? sum[h] / (cnt[h] * ph.voxelVolume()) : 0.0;
    const double c = kerma[0] * 10.0 / b.collimation;
    const double p = 0.25 * (kerma[1] + kerma[2] + kerma[3] + kerma[4]) * 10.0 / b.collimation;
    const double ctdiw = c / 3.0 + 2.0 * p / 3.0;
    const double perHistory = ctdiw / static_cast<double>(nExp * cb.particles_per_exposure);
    double perRot = static_cast<double>(b.particles_per_exposure) * (2 * PI / stepOf(b));
    if (b.type == DXB_BEAM_CT_SPIRAL_DUAL)
        perRot *= 2;
    const double target = b.type == DXB_BEAM_CT_SEQUENTIAL ? b.ctdi : b.ctdi * b.pitch;
    return target / (perHistory * perRot);
}

double analyticCalibration(const orc_world& w, const dxb_beam_desc& b)
{
    const double nTotal = static_cast<double>(numberOfExposures(b)) * static_cast<double>(b.particles_per_exposure);
    if (b.type == DXB_BEAM_PENCIL) {
        const double e = std::clamp(b.energy, 1.0, 150.0);
        const double per = e * w.air.massEnergyTransfer(e);
        return b.air_kerma > 0 && per > 0 ? b.air_kerma / (per * nTotal) : KEV_PER_GRAM_TO_MGY;
    }
    if (b.type == DXB_BEAM_DX || b.type == DXB_BEAM_CBCT) {
        const dxb_spectrum& s = b.spectrum[0];
        double sw = 0, k = 0;
        for (uint32_t i = 0; i < s.n; ++i) {
            sw += s.weight[i];
            k += s.weight[i] * s.energy_kev[i] * w.air.massEnergyTransfer(std::clamp(s.energy_kev[i], 1.0, 150.0));
        }
        if (!(sw > 0) || !(k > 0) || !(b.dap > 0))
            return KEV_PER_GRAM_TO_MGY;
        return b.dap / (k / sw * nTotal);
    }
    return KEV_PER_GRAM_TO_MGY;
}

} // namespace

// =========================================================================== C interface
extern "C" {

// slab-local majorants: the table the device built (dxb_get_local_majorant), or none (n_slabs < 2)
void orc_world_set_local_majorant(orc_world* w, int shift, int n_slabs, const float* ratio)
{
    AAVoxelGrid& g = w->grid;
    g.lmShift = shift;
    g.lmSlabs = (n_slabs >= 2 && ratio) ? n_slabs : 0;
    g.lmRatio.clear();
    if (g.lmSlabs)
        g.lmRatio.assign(ratio, ratio + static_cast<size_t>(n_slabs) * 16);
}
// ... or the oracle's own table in f64 (no GPU needed): slabs of 2^shift voxel layers; returns the number of slabs
int orc_world_build_local_majorant(orc_world* w, int shift)
{
    AAVoxelGrid& g = w->grid;
    const long nz = static_cast<long>(g.dim[2]);
    const int slabs = static_cast<int>((nz + (1 << shift) - 1) >> shift);
    g.lmShift = shift;
    g.lmSlabs = slabs >= 2 ? slabs : 0;
    g.lmRatio.assign(static_cast<size_t>(slabs) * 16, 1.0);
    if (!g.lmSlabs)
        return 0;
    const size_t layer = g.dim[0] * g.dim[1];
    const uint32_t nE = g.materials[0].nE;
    for (int sl = 0; sl < slabs; ++sl) {
        std::vector<double> maxDens(g.materials.size(), 0.0);
        const size_t b = static_cast<size_t>(sl << shift) * layer, e = std::min(g.density.size(), static_cast<size_t>((sl + 1) << shift) * layer);
        for (size_t i = b; i < e; ++i)
            maxDens[g.materialIndex[i]] = std::max(maxDens[g.materialIndex[i]], g.density[i]);
        for (uint32_t band = 0; band * 32 < nE; ++band) {
            double r = 0;
            for (uint32_t node = band * 32; node <= std::min(band * 32 + 32, nE - 1); ++node) {
                double mu = 0;
                for (size_t k = 0; k < g.materials.size(); ++k)
                    mu = std::max(mu, maxDens[k] * (g.materials[k].photo[node] + g.materials[k].incoh[node] + g.materials[k].coh[node]));
                r = std::max(r, mu / g.woodcockStepTable[node]);
            }
            r = std::min(1.0, std::max(r, 1e-6) * (1.0 + 1e-6));
            g.lmRatio[static_cast<size_t>(sl) * 16 + band] = r;
        }
    }
    return g.lmSlabs;
}

// dense-box tracking: the box and outside ratios the device built (dxb_get_dense_box), or none (faces == NULL)
void orc_world_set_dense_box(orc_world* w, const float* faces, const float* ratio)
{
    AAVoxelGrid& g = w->grid;
    g.boxOn = faces && ratio;
    if (!g.boxOn)
        return;
    for (int i = 0; i < 3; ++i) {
        g.boxLo[i] = faces[i];
        g.boxHi[i] = faces[i + 3];
    }
    for (int b = 0; b < 16; ++b)
        g.boxRatio[b] = ratio[b];
}
// ... or the oracle's own box in f64 (no GPU needed): a voxel is thin when its attenuation stays below theta x majorant at
// every energy; the box is the bounding box of all other voxels.  Returns 1 if the box is a proper part of the grid.
int orc_world_build_dense_box(orc_world* w, double theta)
{
    AAVoxelGrid& g = w->grid;
    g.boxOn = false;
    const uint32_t nE = g.materials[0].nE;
    std::vector<double> gm(g.materials.size(), 0.0);
    for (size_t k = 0; k < g.materials.size(); ++k)
        for (uint32_t node = 0; node < nE; ++node)
            gm[k] = std::max(gm[k], (g.materials[k].photo[node] + g.materials[k].incoh[node] + g.materials[k].coh[node]) / g.woodcockStepTable[node]);
    long lo[3] = { static_cast<long>(g.dim[0]), static_cast<long>(g.dim[1]), static_cast<long>(g.dim[2]) }, hi[3] = { -1, -1, -1 };
    size_t i = 0;
    for (long z = 0; z < static_cast<long>(g.dim[2]); ++z)
        for (long y = 0; y < static_cast<long>(g.dim[1]); ++y)
            for (long x = 0; x < static_cast<long>(g.dim[0]); ++x, ++i) {
                if (g.density[i] * gm[g.materialIndex[i]] <= theta)
                    continue;
                const long c[3] = { x, y, z };
                for (int a = 0; a < 3; ++a) {
                    lo[a] = std::min(lo[a], c[a]);
                    hi[a] = std::max(hi[a], c[a]);
                }
            }
    if (hi[0] < 0)
        return 0;
    double vol = 1;
    for (int a = 0; a < 3; ++a) {
        g.boxLo[a] = mf(g.aabb[a] + static_cast<double>(lo[a]) * g.spacing[a]);
        g.boxHi[a] = mf(g.aabb[a] + static_cast<double>(hi[a] + 1) * g.spacing[a]);
        vol *= static_cast<double>(hi[a] + 1 - lo[a]) / static_cast<double>(g.dim[a]);
    }
    std::vector<double> maxDens(g.materials.size(), 0.0);
    i = 0;
    for (long z = 0; z < static_cast<long>(g.dim[2]); ++z)
        for (long y = 0; y < static_cast<long>(g.dim[1]); ++y)
            for (long x = 0; x < static_cast<long>(g.dim[0]); ++x, ++i)
                if (x < lo[0] || x > hi[0] || y < lo[1] || y > hi[1] || z < lo[2] || z > hi[2])
                    maxDens[g.materialIndex[i]] = std::max(maxDens[g.materialIndex[i]], g.density[i]);
    for (uint32_t band = 0; band < 16; ++band) {
        double r = 0;
        for (uint32_t node = band * 32; node <= std::min(band * 32 + 32, nE - 1) && band * 32 < nE; ++node) {
            double mu = 0;
            for (size_t k = 0; k < g.materials.size(); ++k)
                mu = std::max(mu, maxDens[k] * (g.materials[k].photo[node] + g.materials[k].incoh[node] + g.materials[k].coh[node]));
            r = std::max(r, mu / g.woodcockStepTable[node]);
        }
        g.boxRatio[band] = mf(std::min(1.0, std::max(r, 1e-6) * (1.0 + 1e-6)));
    }
    g.boxOn = vol < 1.0;
    return g.boxOn ? 1 : 0;
}

void orc_set_device_mirroring(int on) { g_mirror = on != 0; }
int orc_get_device_mirroring(void) { return g_mirror ? 1 : 0; }

orc_world* orc_world_create(const uint64_t dim[3], const double spacing[3], const double* density, const uint8_t* material,
    uint32_t n_materials, const dxb_material_tables* tables)
{
    auto* w = new orc_world();
    AAVoxelGrid& g = w->grid;
    const size_t n = static_cast<size_t>(dim[0]) * dim[1] * dim[2];
    for (int i = 0; i < 3; ++i) {
        g.dim[i] = dim[i];
        g.spacing[i] = spacing[i];
    }
    g.density.assign(density, density + n);
    g.materialIndex.assign(material, material + n);
    g.materials.resize(n_materials);
    for (uint32_t m = 0; m < n_materials; ++m)
        g.materials[m].load(tables[m]);
    g.build();
    return w;
}
void orc_world_destroy(orc_world* w) { delete w; }

void orc_world_set_reference_materials(orc_world* w, const dxb_material_tables* air, const dxb_material_tables* pmma,
    double air_density, double pmma_density)
{
    w->air.load(*air);
    w->pmma.load(*pmma);
    w->airDensity = air_density;
    w->pmmaDensity = pmma_density;
    w->haveReference = true;
}

int orc_run(orc_world* w, const dxb_beam_desc* b, int physics_mode, uint64_t seed, int n_threads, uint64_t rank, uint64_t world,
    double* energy, double* energy_sq, uint64_t* n_events, orc_stats* stats)
{
    if (!w || !b)
        return DXB_EINVAL;
    w->grid.clearEnergyScored();
    w->grid.scoreMaterial = -1;
    runBeam(w->grid, *b, physics_mode, seed, n_threads, rank, world == 0 ? 1 : world, stats);
    const size_t n = w->grid.size();
    if (energy)
        std::copy(w->grid.energyImparted.begin(), w->grid.energyImparted.end(), energy);
    if (energy_sq)
        std::copy(w->grid.energyImpartedSquared.begin(), w->grid.energyImpartedSquared.end(), energy_sq);
    if (n_events)
        std::copy(w->grid.nEvents.begin(), w->grid.nEvents.end(), n_events);
    (void)n;
    return DXB_OK;
}

double orc_ct_calibration(orc_world* w, const dxb_beam_desc* b, int physics_mode, uint64_t seed, uint64_t calibration_histories,
    int n_threads)
{
    if (!w || !b || !w->haveReference)
        return 0.0;
    return ctCalibration(*w, *b, physics_mode, seed, calibration_histories, n_threads);
}

int orc_transport(orc_world* w, const dxb_beam_desc* b, int physics_mode, int use_beam_calibration, uint64_t seed,
    uint64_t calibration_histories, int n_threads, double* dose, double* variance, uint64_t* n_events, orc_stats* stats)
{
    if (!w || !b)
        return DXB_EINVAL;
    // Transport::operator() [R]: clear energy, run, calibration factor, addEnergyScoredToDoseScore
    w->grid.clearEnergyScored();
    w->grid.scoreMaterial = -1;
    runBeam(w->grid, *b, physics_mode, seed, n_threads, 0, 1, stats);
    double factor = KEV_PER_GRAM_TO_MGY;
    if (use_beam_calibration) {
        if (!w->haveReference)
            return DXB_ESTATE;
        factor = isCt(b->type) ? ctCalibration(*w, *b, physics_mode, seed, calibration_histories, n_threads) : analyticCalibration(*w, *b);
    }
    if (stats)
        stats->calibration_factor = factor;
    const AAVoxelGrid& g = w->grid;
    const double vol = g.voxelVolume();
    for (size_t i = 0; i < g.size(); ++i) {
        const uint64_t nn = g.nEvents[i];
        if (nn == 0 || !(g.density[i] > 0))
            continue;
        // DoseScore::addScoredEnergy [R]: dose += E k/(rho V); variance of the sum of nn events
        const double e = g.energyImparted[i], e2 = g.energyImpartedSquared[i];
        const double varE = std::max(0.0, e2 - e * e / static_cast<double>(nn));
        const double f = factor / (g.density[i] * vol);
        if (dose)
            dose[i] += e * f;
        if (variance)
            variance[i] += varE * f * f;
        if (n_events)
            n_events[i] += nn;
    }
    return DXB_OK;
}

void orc_philox4x32_10(const uint32_t key[2], const uint32_t ctr[4], uint32_t out[4]) { philox(key, ctr, out); }
uint64_t orc_beam_number_of_exposures(const dxb_beam_desc* b) { return b ? numberOfExposures(*b) : 0; }
int orc_beam_exposure(const dxb_beam_desc* b, uint64_t index, dxb_exposure* out)
{
    if (!b || !out)
        return DXB_EINVAL;
    AEC aec(b->aec);
    Exposure e;
    if (!exposureOf(*b, index, aec, e))
        return DXB_EINVAL;
    std::memset(out, 0, sizeof(*out));
    for (int i = 0; i < 3; ++i) {
        out->position[i] = e.pos[i];
        out->cosines[0][i] = e.dirCosX[i];
        out->cosines[1][i] = e.dirCosY[i];
        out->direction[i] = e.dir[i];
    }
    out->half_angles[0] = e.halfAngles[0];
    out->half_angles[1] = e.halfAngles[1];
    out->weight = e.weight;
    out->n_particles = b->particles_per_exposure;
    out->tube = e.tube;
    return DXB_OK;
}
double orc_bowtie_weight(const dxb_bowtie* b, double angle) { return b ? Bowtie(*b)(angle) : 1.0; }
double orc_aec_weight(const dxb_aec* a, const double pos[3]) { return a ? AEC(*a)(Vec { pos[0], pos[1], pos[2] }) : 1.0; }
double orc_organ_aec_weight(const dxb_organ_aec* o, double angle) { return o ? organAec(*o, angle) : 1.0; }

void orc_attenuation(const dxb_material_tables* t, double e, double out[4])
{
    OMaterial m;
    m.load(*t);
    const auto a = m.attenuationValues(e);
    out[0] = a.photoelectric;
    out[1] = a.incoherent;
    out[2] = a.coherent;
    out[3] = a.sum();
}
double orc_majorant(const orc_world* w, double e) { return w->grid.majorant(e); }

void orc_sample_compton(const dxb_material_tables* t, int mode, double energy, uint64_t seed, uint64_t n, double* cosT, double* ratio)
{
    OMaterial m;
    m.load(*t);
    for (uint64_t i = 0; i < n; ++i) {
        RandomState st(seed, i);
        Particle p { { 0, 0, 0 }, { 0, 0, 1 }, energy, 1.0 };
        comptonScatter(p, m, mode, st, nullptr, ratio ? ratio + i : nullptr, cosT ? cosT + i : nullptr);
    }
}
void orc_sample_rayleigh(const dxb_material_tables* t, int mode, double energy, uint64_t seed, uint64_t n, double* cosT)
{
    OMaterial m;
    m.load(*t);
    for (uint64_t i = 0; i < n; ++i) {
        RandomState st(seed, i);
        Particle p { { 0, 0, 0 }, { 0, 0, 1 }, energy, 1.0 };
        rayleightScatter(p, m, mode, st, cosT + i);
    }
}
void orc_sample_source(const dxb_beam_desc* b, uint64_t seed, uint64_t first, uint64_t n, double* pos3, double* dir3, double* energy,
    double* weight)
{
    BeamModel beam(*b);
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t h = first + i;
        Exposure e;
        exposureOf(*b, h / b->particles_per_exposure, beam.aec, e);
        RandomState st(seed, h);
        const Particle p = beam.sampleParticle(e, st);
        for (int k = 0; k < 3; ++k) {
            if (pos3)
                pos3[3 * i + k] = p.pos[k];
            if (dir3)
                dir3[3 * i + k] = p.dir[k];
        }
        if (energy)
            energy[i] = p.energy;
        if (weight)
            weight[i] = p.weight;
    }
}

void orc_organ_dose(const double* dose, const double* density, const uint8_t* organ, uint64_t n, double voxel_volume,
    uint32_t n_organs, double* dose_out, double* mass_out, uint64_t* count_out)
{
    // R:src/libopendxmc/dosetablepipeline.cpp:60-84: energy_imparted = dose*mass; per organ energy/mass
    std::vector<double> energy(n_organs, 0.0), mass(n_organs, 0.0);
    std::vector<uint64_t> cnt(n_organs, 0);
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t o = organ[i];
        if (o >= n_organs)
            continue;
        const double m = voxel_volume * density[i];
        energy[o] += dose[i] * m;
        mass[o] += m;
        ++cnt[o];
    }
    for (uint32_t o = 0; o < n_organs; ++o) {
        if (dose_out)
            dose_out[o] = mass[o] > 0 ? energy[o] / mass[o] : 0.0;
        if (mass_out)
            mass_out[o] = mass[o];
        if (count_out)
            count_out[o] = cnt[o];
    }
}

int orc_postprocess(double* dose, double* variance, double* events, const uint8_t* material, uint64_t n, int delete_air)
{
    // R:src/libopendxmc/simulationpipeline.cpp:180-195, 206-211, 221-229
    if (delete_air)
        for (uint64_t i = 0; i < n; ++i)
            if (material[i] == 0) {
                dose[i] = 0;
                if (events)
                    events[i] = 0;
                if (variance)
                    variance[i] = 0;
            }
    double mx = 0;
    for (uint64_t i = 0; i < n; ++i)
        mx = std::max(mx, dose[i]);
    const bool micro = mx < 1.0;
    if (micro)
        for (uint64_t i = 0; i < n; ++i) {
            dose[i] *= 1e3;
            if (variance)
                variance[i] *= 1e6;
        }
    return micro ? 1 : 0;
}

void orc_segment(const double* hu, uint64_t n, const double* sep, int n_sep, const double* mat_att, double water_att_dens,
    double air_att_dens, uint8_t* material, double* density)
{
    // R:src/libopendxmc/ctsegmentationpipeline.cpp:136-156
    for (uint64_t i = 0; i < n; ++i) {
        uint8_t m = static_cast<uint8_t>(n_sep);
        for (int t = 0; t < n_sep; ++t)
            if (hu[i] < sep[t]) {
                m = static_cast<uint8_t>(t);
                break;
            }
        material[i] = m;
        const double dens = ((water_att_dens - air_att_dens) * hu[i] / 1000.0 + water_att_dens) / mat_att[m];
        density[i] = std::max(dens, 0.0);
    }
}

} // extern "C"
