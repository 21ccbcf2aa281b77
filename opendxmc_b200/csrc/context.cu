// context.cu — dxb_ctx: World<AAVoxelGrid<5,C,255>> + Transport on one or more B200s.
//
// Mirrors the driver at R:src/libopendxmc/simulationpipeline.cpp:124-235:
//   setData/setSpacing/build  -> dxb_set_materials + dxb_set_grid (pack voxels, per-energy majorant)
//   transport(world, beam, progress, true) -> dxb_run (tallies -> reduce -> calibration -> dose)
//   doseScored(i).dose()/variance()/numberOfEvents() -> dxb_get_dose
#include "context_types.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

using namespace dxb;

static_assert(kDevNE == static_cast<int>(kNEnergy), "energy grid mismatch");
static_assert(kDevNX == static_cast<int>(kNX), "x grid mismatch");
static_assert(kDevEPerOctave == static_cast<int>(kENodesPerOctave), "energy grid mismatch");
static_assert(kDevXPerOctave == static_cast<int>(kXNodesPerOctave), "x grid mismatch");
static_assert(sizeof(ExposureDev) == 64, "ExposureDev must be 64 bytes");
static_assert(kShardBlock == DXB_SHARD_BLOCK, "shard block");

namespace dxb {

// Vose's alias method (DXMClib RandomDistribution; restated independently in oracle/oracle.cpp)
void buildAlias(const std::vector<double>& w, std::vector<float>& prob, std::vector<unsigned short>& alias)
{
    const size_t n = w.size();
    prob.assign(n, 1.0f);
    alias.resize(n);
    double sum = 0;
    for (double v : w)
        sum += std::max(0.0, v);
    std::vector<double> q(n);
    std::vector<size_t> small, large;
    for (size_t i = 0; i < n; ++i) {
        q[i] = sum > 0 ? std::max(0.0, w[i]) / sum * static_cast<double>(n) : 1.0;
        alias[i] = static_cast<unsigned short>(i);
    }
    for (size_t i = 0; i < n; ++i)
        (q[i] < 1.0 ? small : large).push_back(i);
    while (!small.empty() && !large.empty()) {
        const size_t s = small.back();
        small.pop_back();
        const size_t l = large.back();
        large.pop_back();
        prob[s] = static_cast<float>(q[s]);
        alias[s] = static_cast<unsigned short>(l);
        q[l] = (q[l] + q[s]) - 1.0;
        (q[l] < 1.0 ? small : large).push_back(l);
    }
    for (size_t i : small)
        prob[i] = 1.0f;
    for (size_t i : large)
        prob[i] = 1.0f;
}

int uploadTables(dxb_ctx* c, World& w, const std::vector<std::shared_ptr<Material>>& mats, cudaStream_t s)
{
    const int n = static_cast<int>(mats.size());
    std::vector<float4> att(static_cast<size_t>(n) * kDevNE);
    std::vector<float> tot(att.size()), etr(att.size());
    std::vector<float> ff(static_cast<size_t>(n) * kDevNX), sf(ff.size());
    std::vector<ShellDev> shells(static_cast<size_t>(n) * kMaxShells);
    std::vector<int> nsh(n);
    std::vector<float> restJ(n);
    for (int m = 0; m < n; ++m) {
        const Material& M = *mats[m];
        for (int i = 0; i < kDevNE; ++i) {
            const double p = M.photo[i], in = M.incoh[i], co = M.coh[i];
            // the total is rounded from the double sum (what the oracle uses), not re-summed in float
            att[static_cast<size_t>(m) * kDevNE + i] = make_float4(static_cast<float>(p), static_cast<float>(in),
                static_cast<float>(co), static_cast<float>(p + in + co));
            tot[static_cast<size_t>(m) * kDevNE + i] = static_cast<float>(p + in + co);
            etr[static_cast<size_t>(m) * kDevNE + i] = static_cast<float>(M.etr[i]);
        }
        for (int i = 0; i < kDevNX; ++i) {
            ff[static_cast<size_t>(m) * kDevNX + i] = static_cast<float>(M.ffCdf[i]);
            sf[static_cast<size_t>(m) * kDevNX + i] = static_cast<float>(M.sf[i]);
        }
        nsh[m] = static_cast<int>(M.nShells);
        restJ[m] = static_cast<float>(M.restComptonJ0);
        for (uint32_t k = 0; k < M.nShells; ++k) {
            ShellDev& d = shells[static_cast<size_t>(m) * kMaxShells + k];
            const dxb_shell& h = M.shells[k];
            d.binding = static_cast<float>(h.binding_energy_kev);
            d.nel_fraction = static_cast<float>(h.n_electrons_fraction);
            d.photo_fraction = static_cast<float>(h.photo_fraction_above);
            d.fluor_yield = static_cast<float>(h.fluor_yield);
            d.fluor_energy = static_cast<float>(h.fluor_energy_kev);
            d.j0 = static_cast<float>(h.compton_j0);
            d.pad0 = d.pad1 = 0.0f;
        }
    }
    w.n_mat = n;
    w.hostTot = tot;
    CUDA_TRY(c, w.att.upload(att, w.device, s));
    CUDA_TRY(c, w.tot.upload(tot, w.device, s));
    CUDA_TRY(c, w.etr.upload(etr, w.device, s));
    CUDA_TRY(c, w.ffcdf.upload(ff, w.device, s));
    CUDA_TRY(c, w.sf.upload(sf, w.device, s));
    CUDA_TRY(c, w.shells.upload(shells, w.device, s));
    CUDA_TRY(c, w.nshells.upload(nsh, w.device, s));
    CUDA_TRY(c, w.restJ0.upload(restJ, w.device, s));
    CUDA_TRY(c, w.majorant.alloc(kDevNE, w.device));
    CUDA_TRY(c, cudaStreamSynchronize(s)); // host vectors go out of scope
    w.hasTables = true;
    return DXB_OK;
}

constexpr int kRefNode = 378; // energy node of 60 keV: the auto rules of the slab table and of the dense box look at this energy

// Slab-local majorants (transport_pool.cu, LM builds).  Slabs of 2^shift voxel layers along z, about opt.slabCm thick.  The
// device finds, per slab and material, the largest density (exact integer maxima of the 24-bit densities); the small table
//   ratio(slab, band) = max over the band's energy nodes of [ max_m rho_max(slab, m) * tot_m(node) / majorant(node) ]
// is then computed HERE with single IEEE f32 operations on the same f32 tables the device reads, so that the oracle, which
// receives the table through dxb_get_local_majorant, tracks with exactly the numbers the kernel uses.
int buildLocalMajorant(dxb_ctx* c, World& w, cudaStream_t s)
{
    w.lmSlabs = 0;
    w.lmUseful = false;
    const int nz = static_cast<int>(w.dim[2]);
    int shift = 0;
    while ((2 << shift) * w.spacing[2] <= c->opt.slabCm * 1.5 && (nz >> (shift + 1)) >= 1 && shift < 12)
        ++shift;
    int slabs = (nz + (1 << shift) - 1) >> shift;
    while (slabs > kLmMaxSlabs) {
        ++shift;
        slabs = (nz + (1 << shift) - 1) >> shift;
    }
    if (slabs < 2 || w.hostTot.size() != static_cast<size_t>(w.n_mat) * kDevNE)
        return DXB_OK; // a single slab is the global majorant
    CUDA_TRY(c, w.slabMax.alloc(static_cast<size_t>(slabs) * 256, w.device));
    CUDA_TRY(c, cudaMemsetAsync(w.slabMax.p, 0, static_cast<size_t>(slabs) * 256 * sizeof(unsigned int), s));
    launchSlabMax(w.voxels.p, static_cast<size_t>(w.dim[0]) * w.dim[1], nz, shift, slabs, w.slabMax.p, s);
    CUDA_TRY(c, cudaGetLastError());
    std::vector<unsigned int> bits(static_cast<size_t>(slabs) * 256);
    std::vector<float> maj(kDevNE);
    CUDA_TRY(c, cudaMemcpyAsync(bits.data(), w.slabMax.p, bits.size() * sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaMemcpyAsync(maj.data(), w.majorant.p, kDevNE * sizeof(float), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaStreamSynchronize(s));
    w.lmHost.assign(static_cast<size_t>(slabs) * kLmBands, 1.0f);
    double meanRatio = 0;
    const int refBand = kRefNode >> 5; // the band of 60 keV: where a diagnostic spectrum has most of its photons
    for (int sl = 0; sl < slabs; ++sl) {
        for (int b = 0; b < kLmBands; ++b) {
            const int n0 = b * 32, n1 = std::min(b * 32 + 32, kDevNE - 1);
            if (n0 >= kDevNE)
                break;
            float r = 0.0f;
            for (int node = n0; node <= n1; ++node) {
                float mu = 0.0f;
                for (int m = 0; m < w.n_mat; ++m) {
                    float rho;
                    const unsigned int q = bits[static_cast<size_t>(sl) * 256 + m];
                    std::memcpy(&rho, &q, sizeof(float));
                    const float v = rho * w.hostTot[static_cast<size_t>(m) * kDevNE + node];
                    mu = v > mu ? v : mu;
                }
                const float ratio = mu / maj[node];
                r = ratio > r ? ratio : r;
            }
            // a little head-room (2^-20) keeps the local attenuation below majorant * ratio under f32 rounding
            r = std::min(1.0f, std::max(r, 1.0e-6f) * (1.0f + 9.5367431640625e-7f));
            w.lmHost[static_cast<size_t>(sl) * kLmBands + b] = r;
            if (b == refBand)
                meanRatio += r;
        }
    }
    meanRatio /= slabs;
    CUDA_TRY(c, w.lmRatio.upload(w.lmHost, w.device, s));
    CUDA_TRY(c, cudaStreamSynchronize(s));
    w.lmShift = shift;
    w.lmSlabs = slabs;
    w.lmUseful = meanRatio < 0.8; // hops and the longer step code must be paid for
    return DXB_OK;
}

// Dense box (DB builds of the pool kernel): most of a CT volume is air around the patient, and with one majorant a photon pays
// a tentative step every ~2 cm of it.  A voxel is THIN when its attenuation stays below theta x majorant at every energy node;
// the dense box is the bounding box of all other voxels.  Inside it the kernel tracks with the global majorant, in the rest
// of the grid with majorant x ratio[band], ratio = largest attenuation occurring outside the box / majorant (maximised over
// the band's nodes, as for the slab table).  Box and ratios are what the oracle receives through dxb_get_dense_box.
int buildDenseBox(dxb_ctx* c, World& w, cudaStream_t s)
{
    w.dbBuilt = false;
    w.dbUseful = false;
    if (w.hostTot.size() != static_cast<size_t>(w.n_mat) * kDevNE)
        return DXB_OK;
    std::vector<float> maj(kDevNE);
    CUDA_TRY(c, cudaMemcpyAsync(maj.data(), w.majorant.p, kDevNE * sizeof(float), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaStreamSynchronize(s));
    // thin limit per material: rho * max_E tot_m(E) / majorant(E) <= theta
    std::vector<unsigned int> scratch(256 + 256 + 6, 0u);
    for (int m = 0; m < w.n_mat; ++m) {
        double g = 0;
        for (int node = 0; node < kDevNE; ++node)
            if (maj[node] > 0)
                g = std::max(g, static_cast<double>(w.hostTot[static_cast<size_t>(m) * kDevNE + node]) / maj[node]);
        const float limit = g > 0 ? static_cast<float>(std::min(c->opt.denseTheta / g, 3.0e38)) : 3.0e38f;
        unsigned int bits;
        std::memcpy(&bits, &limit, sizeof(bits));
        scratch[m] = bits & 0xFFFFFF00u; // (rounded down: a voxel exactly at the limit counts as dense)
    }
    const int nx = static_cast<int>(w.dim[0]), ny = static_cast<int>(w.dim[1]), nz = static_cast<int>(w.dim[2]);
    int* boxInit = reinterpret_cast<int*>(scratch.data() + 512);
    boxInit[0] = nx, boxInit[1] = ny, boxInit[2] = nz, boxInit[3] = boxInit[4] = boxInit[5] = -1;
    CUDA_TRY(c, w.dbScratch.upload(scratch, w.device, s));
    launchDenseBox(w.voxels.p, nx, ny, nz, w.dbScratch.p, reinterpret_cast<int*>(w.dbScratch.p + 512), s);
    CUDA_TRY(c, cudaGetLastError());
    launchOutsideMax(w.voxels.p, nx, ny, nz, reinterpret_cast<const int*>(w.dbScratch.p + 512), w.dbScratch.p + 256, s);
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaMemcpyAsync(scratch.data(), w.dbScratch.p, scratch.size() * sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaStreamSynchronize(s));
    const int* box = reinterpret_cast<const int*>(scratch.data() + 512);
    if (box[3] < 0)
        return DXB_OK; // nothing but thin voxels: no box
    double part = 1.0;
    for (int a = 0; a < 3; ++a) {
        w.dbBox[a] = box[a];
        w.dbBox[a + 3] = box[a + 3] + 1;
        const float d = static_cast<float>(w.spacing[a]);
        const float lo = static_cast<float>(-0.5 * static_cast<double>(w.dim[a]) * w.spacing[a]);
        w.dbFaces[a] = std::fmaf(static_cast<float>(w.dbBox[a]), d, lo);
        w.dbFaces[a + 3] = std::fmaf(static_cast<float>(w.dbBox[a + 3]), d, lo);
        part *= static_cast<double>(w.dbBox[a + 3] - w.dbBox[a]) / static_cast<double>(w.dim[a]);
    }
    const int refBand = kRefNode >> 5;
    for (int b = 0; b < kLmBands; ++b) {
        const int n0 = b * 32, n1 = std::min(b * 32 + 32, kDevNE - 1);
        float r = 0.0f;
        for (int node = n0; node <= n1 && n0 < kDevNE; ++node) {
            float mu = 0.0f;
            for (int m = 0; m < w.n_mat; ++m) {
                float rho;
                const unsigned int q = scratch[256 + m];
                std::memcpy(&rho, &q, sizeof(float));
                const float v = rho * w.hostTot[static_cast<size_t>(m) * kDevNE + node];
                mu = v > mu ? v : mu;
            }
            const float ratio = mu / maj[node];
            r = ratio > r ? ratio : r;
        }
        w.dbRatio[b] = std::min(1.0f, std::max(r, 1.0e-6f) * (1.0f + 9.5367431640625e-7f));
    }
    w.dbBuilt = true;
    // Worth it when the flights replace enough tentative steps to pay for themselves (a flight costs about three steps,
    // profiles/r02_sweep_densebox.txt): the steps saved per history are about mu_max x the path outside the box, estimated
    // from the linear sizes of grid and box (C2 / C4: 4 steps -> +13 / +15 %; the CTDI phantom: 0.6 steps -> off, it lost 6 %).
    double vol = 1.0;
    for (int a = 0; a < 3; ++a)
        vol *= static_cast<double>(w.dim[a]) * w.spacing[a];
    const double outsidePath = std::cbrt(vol) - std::cbrt(vol * part);
    const double savedSteps = static_cast<double>(maj[kRefNode]) * outsidePath;
    w.dbUseful = part < 0.85 && w.dbRatio[refBand] < 0.25f && savedSteps >= 2.5;
    return DXB_OK;
}

// majorant from the per-material density maxima + the reference's material-index check; the grid becomes usable
int finishGrid(dxb_ctx* c, World& w, cudaStream_t s)
{
    launchMajorant(w.tot.p, w.maxDensityBits.p, w.n_mat, w.majorant.p, s);
    CUDA_TRY(c, cudaGetLastError());
    // the reference checks max(material) < n_materials before running (R:src/libopendxmc/simulationpipeline.cpp:54-57);
    // here the pack kernel finds the largest index while it reads the array anyway
    unsigned int mmax = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&mmax, w.maxDensityBits.p + 256, sizeof(mmax), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(c, cudaStreamSynchronize(s));
    if (mmax >= static_cast<unsigned int>(w.n_mat))
        return fail(c, DXB_EINVAL, "set_grid: material index out of range");
    w.brickN[0] = 0;
    if (c->opt.brickFilter) {
        int shift = 0;
        while ((2 << shift) <= c->opt.brickVoxels && shift < 8)
            ++shift;
        const int nb[3] = { static_cast<int>((w.dim[0] + (1u << shift) - 1) >> shift), static_cast<int>((w.dim[1] + (1u << shift) - 1) >> shift),
            static_cast<int>((w.dim[2] + (1u << shift) - 1) >> shift) };
        const size_t bricks = static_cast<size_t>(nb[0]) * nb[1] * nb[2];
        if (bricks < (1u << 28)) {
            CUDA_TRY(c, w.brickBound.alloc(bricks * 8, w.device));
            launchBrickBound(w.voxels.p, static_cast<int>(w.dim[0]), static_cast<int>(w.dim[1]), static_cast<int>(w.dim[2]), shift, nb[0], nb[1], nb[2],
                w.tot.p, w.majorant.p, w.n_mat, w.brickBound.p, s);
            CUDA_TRY(c, cudaGetLastError());
            w.brickShift = shift;
            for (int i = 0; i < 3; ++i)
                w.brickN[i] = nb[i];
        }
    }
    if (c->opt.localMajorant != 0) {
        const int rc = buildLocalMajorant(c, w, s);
        if (rc != DXB_OK)
            return rc;
    } else {
        w.lmSlabs = 0;
        w.lmUseful = false;
    }
    if (c->opt.denseBox != 0) {
        const int rc = buildDenseBox(c, w, s);
        if (rc != DXB_OK)
            return rc;
    } else {
        w.dbBuilt = false;
        w.dbUseful = false;
    }
    w.hasGrid = true;
    return DXB_OK;
}

// Uploads and packs voxels [begin, end) of the caller's arrays (the whole grid for dxb_set_grid; one slab per rank for
// dxb_set_grid_sharded, where the ranks then exchange their packed slabs and density maxima over NVLink).
int uploadGrid(dxb_ctx* c, World& w, const uint64_t dim[3], const double spacing[3], const double* density,
    const uint8_t* material, cudaStream_t s, size_t begin, size_t end, bool finish)
{
    const size_t n = static_cast<size_t>(dim[0]) * dim[1] * dim[2];
    for (int i = 0; i < 3; ++i) {
        w.dim[i] = dim[i];
        w.spacing[i] = spacing[i];
    }
    w.nvox = n;
    w.hasGrid = false;
    if (c->ipc && w.tally.p && w.tally.n != n * 4)
        return fail(c, DXB_ESTATE, "set_grid: the grid size changed while peers map the tally buffers (dxb_exchange_export / _import again)");
    CUDA_TRY(c, w.voxels.alloc(n, w.device));
    CUDA_TRY(c, w.tally.alloc(n * 4, w.device));
    CUDA_TRY(c, w.maxDensityBits.alloc(257, w.device));
    const size_t m = end - begin;
    // staging for the slab of the caller's f64 density / u8 material this device packs (kept between set_grid calls)
    CUDA_TRY(c, w.stageDensity.alloc(m, w.device));
    CUDA_TRY(c, w.stageMaterial.alloc(m, w.device));
    CUDA_TRY(c, cudaMemsetAsync(w.maxDensityBits.p, 0, 257 * sizeof(unsigned int), s));
    if (m > 0) {
        CUDA_TRY(c, cudaMemcpyAsync(w.stageDensity.p, density + begin, m * sizeof(double), cudaMemcpyHostToDevice, s));
        CUDA_TRY(c, cudaMemcpyAsync(w.stageMaterial.p, material + begin, m, cudaMemcpyHostToDevice, s));
        launchPackVoxels(w.stageDensity.p, w.stageMaterial.p, w.voxels.p + begin, m, w.maxDensityBits.p, s);
        CUDA_TRY(c, cudaGetLastError());
    }
    w.cur = 0;
    CUDA_TRY(c, cudaMemsetAsync(w.tally.p, 0, n * 4 * sizeof(unsigned long long), s));
    return finish ? finishGrid(c, w, s) : DXB_OK;
}

// FNV-1a over everything that defines a beam: the descriptor (pointers blanked) and the arrays it points to
uint64_t beamHash(const dxb_beam_desc& b)
{
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t n) {
        const auto* c = static_cast<const unsigned char*>(p);
        for (size_t i = 0; i < n; ++i)
            h = (h ^ c[i]) * 1099511628211ull;
    };
    dxb_beam_desc d = b;
    for (int t = 0; t < 2; ++t) {
        d.spectrum[t].energy_kev = d.spectrum[t].weight = nullptr;
        d.bowtie[t].angle_rad = d.bowtie[t].weight = nullptr;
    }
    d.aec.weights = nullptr;
    mix(&d, sizeof(d));
    for (int t = 0; t < 2; ++t) {
        if (b.spectrum[t].n && b.spectrum[t].energy_kev && b.spectrum[t].weight) {
            mix(b.spectrum[t].energy_kev, b.spectrum[t].n * sizeof(double));
            mix(b.spectrum[t].weight, b.spectrum[t].n * sizeof(double));
        }
        if (b.bowtie[t].n && b.bowtie[t].angle_rad && b.bowtie[t].weight) {
            mix(b.bowtie[t].angle_rad, b.bowtie[t].n * sizeof(double));
            mix(b.bowtie[t].weight, b.bowtie[t].n * sizeof(double));
        }
    }
    if (b.aec.n && b.aec.weights)
        mix(b.aec.weights, b.aec.n * sizeof(double));
    return h ? h : 1;
}

int prepareBeam(dxb_ctx* c, const dxb_beam_desc& b, PreparedBeam& out)
{
    const uint64_t nExp = beamNumberOfExposures(b);
    if (nExp == 0 || b.particles_per_exposure == 0)
        return fail(c, DXB_EINVAL, "beam has no exposures or no particles");
    if (nExp > (1ull << 26))
        return fail(c, DXB_EINVAL, "too many exposures");
    const AecTable aec = makeAec(b.aec);
    out.exposures.resize(nExp);
    double wmax = 0;
    for (uint64_t i = 0; i < nExp; ++i) {
        dxb_exposure e;
        if (beamExposure(b, i, aec, e) != DXB_OK)
            return fail(c, DXB_EINVAL, "bad exposure");
        ExposureDev& d = out.exposures[i];
        for (int k = 0; k < 3; ++k) {
            d.pos[k] = static_cast<float>(e.position[k]);
            d.c0[k] = static_cast<float>(e.cosines[0][k]);
            d.c1[k] = static_cast<float>(e.cosines[1][k]);
            d.dir[k] = static_cast<float>(e.direction[k]);
        }
        d.hx = static_cast<float>(e.half_angles[0]);
        d.hy = static_cast<float>(e.half_angles[1]);
        d.weight = static_cast<float>(e.weight);
        d.tube = e.tube;
        wmax = std::max(wmax, e.weight);
    }
    out.ppe = b.particles_per_exposure;
    out.nTotal = nExp * b.particles_per_exposure;
    const int nTubes = (b.type == DXB_BEAM_CT_SPIRAL_DUAL || (b.type == DXB_BEAM_CTDI && b.spectrum[1].n > 0)) ? 2 : 1;
    double bowMax = 1.0;
    for (int t = 0; t < 2; ++t) {
        const dxb_spectrum& s = b.spectrum[t < nTubes ? t : 0];
        if (b.type == DXB_BEAM_PENCIL) {
            out.specN[t] = 1;
            out.specE0[t] = static_cast<float>(b.energy);
            out.specStep[t] = 0;
            out.prob[t].assign(1, 1.0f);
            out.alias[t].assign(1, 0);
        } else {
            if (s.n == 0 || !s.energy_kev || !s.weight)
                return fail(c, DXB_EINVAL, "beam spectrum missing");
            if (s.n > 60000)
                return fail(c, DXB_EINVAL, "spectrum too long");
            // the device samples E = e0 + (bin + u) * step: the grid must be uniform and ascending (what dxmc::Tube::getEnergy()
            // produces); an arbitrary energy array would silently be sampled at the wrong energies
            if (s.n > 1) {
                const double step = s.energy_kev[1] - s.energy_kev[0];
                if (!(step > 0))
                    return fail(c, DXB_EINVAL, "beam spectrum: energies must ascend");
                for (uint32_t i = 2; i < s.n; ++i)
                    if (std::fabs((s.energy_kev[i] - s.energy_kev[i - 1]) - step) > 1e-6 * step)
                        return fail(c, DXB_EINVAL, "beam spectrum: the energy grid must be uniform");
            }
            out.specN[t] = static_cast<int>(s.n);
            out.specE0[t] = static_cast<float>(s.energy_kev[0]);
            out.specStep[t] = s.n > 1 ? static_cast<float>(s.energy_kev[1] - s.energy_kev[0]) : 0.0f;
            std::vector<double> w(s.weight, s.weight + s.n);
            buildAlias(w, out.prob[t], out.alias[t]);
        }
        const BowtieTable bt = makeBowtie(b.bowtie[t < nTubes ? t : 0]);
        out.bowA[t].clear();
        out.bowW[t].clear();
        if (!bt.empty() && b.type != DXB_BEAM_PENCIL && b.type != DXB_BEAM_DX && b.type != DXB_BEAM_CBCT) {
            for (size_t i = 0; i < bt.angle.size(); ++i) {
                out.bowA[t].push_back(static_cast<float>(bt.angle[i]));
                out.bowW[t].push_back(static_cast<float>(bt.weight[i]));
                bowMax = std::max(bowMax, bt.weight[i]);
            }
        }
    }
    out.maxWeight = std::max(1.0, wmax) * bowMax / (1.0 - 0.9);
    return DXB_OK;
}

int uploadBeam(dxb_ctx* c, DeviceState& d, const PreparedBeam& pb)
{
    if (pb.hash && d.uploadedBeam == pb.hash)
        return DXB_OK; // the exposures, alias tables and bowtie knots of this very beam are already on the device
    d.uploadedBeam = 0;
    CUDA_TRY(c, d.exposures.upload(pb.exposures, d.device, d.stream));
    for (int t = 0; t < 2; ++t) {
        CUDA_TRY(c, d.specProb[t].upload(pb.prob[t], d.device, d.stream));
        CUDA_TRY(c, d.specAlias[t].upload(pb.alias[t], d.device, d.stream));
        CUDA_TRY(c, d.bowAngle[t].upload(pb.bowA[t], d.device, d.stream));
        CUDA_TRY(c, d.bowWeight[t].upload(pb.bowW[t], d.device, d.stream));
    }
    // (pageable source vectors: the copies above have been staged by the time the calls return)
    d.uploadedBeam = pb.hash;
    return DXB_OK;
}

// number of local linear indices owned by shard (rank, world) for nTotal global histories
uint64_t localCount(uint64_t nTotal, uint64_t rank, uint64_t world)
{
    const uint64_t nBlocks = (nTotal + kShardBlock - 1) / kShardBlock;
    // blocks rank, rank+world, ...
    if (rank >= nBlocks)
        return 0;
    const uint64_t mine = (nBlocks - rank + world - 1) / world;
    return mine * kShardBlock; // padded; the kernel skips ids >= nTotal
}

struct TransportResult {
    double ms = 0;
    uint64_t launches = 0;
    uint64_t stats[7] = { 0, 0, 0, 0, 0, 0, 0 }; // steps, interactions, deposits, emitted, histories, hops, voxel fetches
    bool cancelled = false;
};

// Runs the histories of one prepared beam through `world` on device state d for shard (rank, world).
int runOnDevice(dxb_ctx* c, DeviceState& d, World& w, const PreparedBeam& pb, int mode, bool calib, int scoreMaterial,
    uint64_t rank, uint64_t world, dxb_progress* progress, bool asyncOnly, TransportResult* res)
{
    CUDA_TRY(c, cudaSetDevice(d.device));
    RunParams P;
    std::memset(&P, 0, sizeof(P));
    P.grid = w.gridDev();
    P.tab = w.tablesDev();
    for (int t = 0; t < 2; ++t) {
        P.spec[t].n = pb.specN[t];
        P.spec[t].e0 = pb.specE0[t];
        P.spec[t].step = pb.specStep[t];
        P.spec[t].prob = d.specProb[t].p;
        P.spec[t].alias = d.specAlias[t].p;
        P.bow[t].n = static_cast<int>(pb.bowA[t].size());
        P.bow[t].angle = d.bowAngle[t].p;
        P.bow[t].weight = d.bowWeight[t].p;
    }
    P.exposures = d.exposures.p;
    P.n_exposures = pb.exposures.size();
    P.ppe = pb.ppe;
    P.n_total = pb.nTotal;
    P.world = static_cast<unsigned int>(world);
    P.rank = static_cast<unsigned int>(rank);
    P.seed_lo = static_cast<unsigned int>(c->runKey);
    P.seed_hi = static_cast<unsigned int>(c->runKey >> 32);
    for (unsigned int r = 0; r < 10; ++r) {
        P.round_key[r][0] = P.seed_lo + r * 0x9E3779B9u;
        P.round_key[r][1] = P.seed_hi + r * 0xBB67AE85u;
    }
    P.tally_scale_e = c->scaleE;
    P.tally_scale_e2 = c->scaleE2;
    P.score_material = calib ? scoreMaterial : -1;
    const bool pool = c->opt.poolSlots > 0;
    const bool mux = pool || c->opt.slots >= 2; // both keep photons in shared memory and share the policy knobs
    // defaults (dead / waiting / Rayleigh lanes that trigger a phase): register kernel 4 / 12 / 4; mux kernel 8 / 33 (bias rule
    // only) / 8; pool kernel 28 / 28 / 20 (profiles/r01_pool_policy_sweep.txt)
    P.refill_threshold = std::clamp(c->opt.refillThreshold > 0 ? c->opt.refillThreshold : (pool ? 28 : mux ? 8 : 4), 1, 32);
    P.interact_threshold = std::clamp(c->opt.interactThreshold > 0 ? c->opt.interactThreshold : (pool ? 28 : mux ? 33 : 12), 1, 33);
    P.rayleigh_threshold = std::clamp(c->opt.rayleighThreshold > 0 ? c->opt.rayleighThreshold : (pool ? 20 : mux ? 8 : 4), 1, 32);
    P.step_pairs = std::clamp(c->opt.stepPairs > 0 ? c->opt.stepPairs : (pool ? 2 : 1), 1, 8);
    P.interact_bias = std::clamp(c->opt.interactBias == -999 ? (pool ? 24 : 16) : c->opt.interactBias, -32, 32);
    P.service_warps = std::clamp(c->opt.serviceWarps, 0, 32);
    P.diag = c->opt.diag;
    P.step_quad = (pool && c->opt.stepQuad && P.step_pairs == 2) ? c->opt.stepQuad : 0;
    P.work_counter = d.counters.p;
    P.stats = d.counters.p + 8;

    LaunchConfig cfg;
    cfg.threads = c->opt.threads;
    const size_t tableBytes = static_cast<size_t>(w.n_mat) * kDevNE * sizeof(float);
    cfg.table_in_smem = c->opt.tableInSmem && tableBytes <= 200 * 1024;
    cfg.slots = 0;
    cfg.pool = pool;
    cfg.min_blocks = c->opt.poolMinBlocks;
    cfg.local_majorant = false;
    cfg.brick_filter = false;
    if (pool) {
        // slab-local majorants: the pool kernel's scoring builds; auto = when the table built with the grid predicts a gain
        const bool lm = !calib && w.lmSlabs >= 2 && (c->opt.localMajorant == 1 || (c->opt.localMajorant < 0 && w.lmUseful));
        cfg.local_majorant = lm;
        const int lmSlabs = lm ? w.lmSlabs : 0;
        cfg.threads = std::clamp(c->opt.poolThreads, 64, 512) / 32 * 32;
        const bool bfWanted = !lm && !calib && c->opt.brickFilter && w.brickN[0] > 0 && w.brickBound.p && c->opt.stepQuad && P.step_pairs == 2;
        cfg.slots = transportPoolSlots(mode, calib, cfg.table_in_smem, c->opt.poolSlots, lm, bfWanted);
        cfg.smem = poolSmemBytes(cfg.slots, cfg.table_in_smem ? w.n_mat * kDevNE : 0, lmSlabs);
        if (cfg.table_in_smem && cfg.smem > 56 * 1024) {
            cfg.table_in_smem = false;
            cfg.slots = transportPoolSlots(mode, calib, false, c->opt.poolSlots, lm, bfWanted);
            cfg.smem = poolSmemBytes(cfg.slots, 0, lmSlabs);
        }
        const bool bf = !lm && !calib && c->opt.brickFilter && w.brickN[0] > 0 && w.brickBound.p && c->opt.stepQuad && P.step_pairs == 2;
        cfg.brick_filter = bf;
        // dense box: the production variant of the quad step (16 slots, 64 registers); auto = when the box built with the grid
        // leaves enough of the grid outside
        const bool db = !lm && !bf && !calib && w.dbBuilt && P.step_quad == 1 && c->opt.poolSlots == 16 && c->opt.poolMinBlocks == 0
            && (c->opt.denseBox == 1 || (c->opt.denseBox < 0 && w.dbUseful));
        cfg.dense_box = db;
        if (db) {
            for (int a = 0; a < 3; ++a) {
                P.db_lo[a] = w.dbFaces[a];
                P.db_hi[a] = w.dbFaces[a + 3];
                P.db_i0[a] = w.dbBox[a];
                P.db_n[a] = w.dbBox[a + 3] - w.dbBox[a];
            }
            for (int b = 0; b < 16; ++b)
                P.db_ratio[b] = w.dbRatio[b];
            double diag2 = 0;
            for (int a = 0; a < 3; ++a)
                diag2 += static_cast<double>(w.dim[a]) * w.spacing[a] * static_cast<double>(w.dim[a]) * w.spacing[a];
            P.db_diag = static_cast<float>(std::sqrt(diag2) * 1.001);
        }
        if (bf) {
            P.brick = w.brickBound.p;
            P.brick_shift = w.brickShift;
            P.brick_nx = w.brickN[0];
            P.brick_ny = w.brickN[1];
        }
        if (lm) {
            P.lm_ratio = w.lmRatio.p;
            P.lm_slabs = w.lmSlabs;
            P.lm_shift = w.lmShift;
            P.lm_thickness = static_cast<float>(static_cast<double>(1 << w.lmShift) * w.spacing[2]);
        }
    } else if (mux) {
        // the table shares the SM's shared memory with the photon slots: keep it only while two blocks still fit
        cfg.slots = transportMuxSlots(mode, calib, cfg.table_in_smem, c->opt.slots);
        cfg.smem = muxSmemBytes(cfg.threads, cfg.slots, cfg.table_in_smem ? w.n_mat * kDevNE : 0);
        if (cfg.table_in_smem && cfg.smem > 110 * 1024) {
            cfg.table_in_smem = false;
            cfg.slots = transportMuxSlots(mode, calib, false, c->opt.slots);
            cfg.smem = muxSmemBytes(cfg.threads, cfg.slots, 0);
        }
    } else {
        cfg.smem = transportSmemBytes(cfg.threads, cfg.table_in_smem ? w.n_mat * kDevNE : 0);
    }
    cfg.smem += static_cast<size_t>(std::max(0, c->opt.smemPadKb)) * 1024;
    int perSm = c->opt.blocksPerSm;
    if (perSm <= 0) {
        perSm = pool ? transportPoolOccupancy(mode, calib, cfg.table_in_smem, cfg.slots, cfg.threads, cfg.smem, cfg.min_blocks, cfg.local_majorant, cfg.brick_filter, cfg.dense_box)
            : mux  ? transportMuxOccupancy(mode, calib, cfg.table_in_smem, cfg.slots, cfg.threads, cfg.smem)
                   : transportOccupancy(mode, calib, cfg.table_in_smem, cfg.threads, cfg.smem);
        if (perSm <= 0)
            return fail(c, DXB_ECUDA, "transport kernel cannot be resident (occupancy 0)");
    }
    cfg.blocks = c->smCount * perSm;
    if (!calib && d.part == 0) // (several devices launch from their own host threads: one of them reports)
    {
        c->stats.local_majorant = cfg.local_majorant ? 1 : 0;
        c->stats.dense_box = cfg.dense_box ? 1 : 0;
    }

    const uint64_t nLocal = localCount(pb.nTotal, rank, world);
    CUDA_TRY(c, cudaMemsetAsync(d.counters.p + 8, 0, 24 * sizeof(unsigned long long), d.stream));
    CUDA_TRY(c, cudaEventRecord(d.evStart, d.stream));
    uint64_t launches = 0;
    bool cancelled = false;
    // the mux kernel keeps 32 bits of the history id per photon: the ids of one launch must span < 2^32
    uint64_t batch = c->opt.batch;
    if (mux)
        batch = std::max<uint64_t>(kShardBlock, std::min<uint64_t>(batch, ((1ull << 31) / world) / kShardBlock * kShardBlock));
    for (uint64_t begin = 0; begin < nLocal; begin += batch) {
        if (progress && progress->stop.load(std::memory_order_relaxed)) {
            cancelled = true;
            break;
        }
        const uint64_t end = std::min(nLocal, begin + batch);
        P.local_begin = begin;
        P.local_end = end;
        const uint64_t firstId = ((begin / kShardBlock) * world + rank) * kShardBlock + begin % kShardBlock;
        P.hbase_lo = static_cast<unsigned int>(firstId);
        P.hbase_hi = static_cast<unsigned int>(firstId >> 32);
        CUDA_TRY(c, cudaMemsetAsync(d.counters.p, 0, sizeof(unsigned long long), d.stream));
        CUDA_TRY(c, pool ? launchTransportPool(P, mode, calib, cfg, d.stream)
                : mux ? launchTransportMux(P, mode, calib, cfg, d.stream)
                      : launchTransport(P, mode, calib, cfg, d.stream));
        ++launches;
        if (progress && !asyncOnly && end < nLocal) {
            // stop must be observed within one batch; progress is published per batch
            CUDA_TRY(c, cudaStreamSynchronize(d.stream));
            progress->done.fetch_add(end - begin, std::memory_order_relaxed);
        }
    }
    CUDA_TRY(c, cudaEventRecord(d.evTransport, d.stream));
    if (res) {
        res->launches = launches;
        res->cancelled = cancelled;
    }
    return DXB_OK;
}

int collectStats(dxb_ctx* c, DeviceState& d, TransportResult& res)
{
    CUDA_TRY(c, cudaSetDevice(d.device));
    CUDA_TRY(c, cudaStreamSynchronize(d.stream));
    unsigned long long h[7];
    CUDA_TRY(c, cudaMemcpy(h, d.counters.p + 8, sizeof(h), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 7; ++i)
        res.stats[i] = h[i];
    if (res.stats[6] == 0)
        res.stats[6] = res.stats[0]; // kernels without the pre-filter fetch the voxel of every tentative step
    if (c->opt.diag) {
        // pool kernel diagnostics: executions and claimed lanes per phase (step, interaction, Rayleigh, refill)
        unsigned long long g[8];
        CUDA_TRY(c, cudaMemcpy(g, d.counters.p + 16, sizeof(g), cudaMemcpyDeviceToHost));
        static const char* names[4] = { "step", "interact", "rayleigh", "refill" };
        for (int i = 0; i < 4; ++i)
            std::fprintf(stderr, "dxb diag: phase %-8s executions %12llu  lanes/execution %5.2f\n", names[i], g[i],
                g[i] ? static_cast<double>(g[4 + i]) / static_cast<double>(g[i]) : 0.0);
    }
    float ms = 0;
    CUDA_TRY(c, cudaEventElapsedTime(&ms, d.evStart, d.evTransport));
    res.ms = ms;
    return DXB_OK;
}

void chooseScales(dxb_ctx* c, const PreparedBeam& pb, bool kerma)
{
    // largest power-of-two quanta such that nTotal * Emax * wmax cannot overflow 2^63
    const double worstE = static_cast<double>(pb.nTotal) * 150.0 * pb.maxWeight * (kerma ? 50.0 : 1.0);
    int se = 24;
    while (se > 0 && worstE * std::ldexp(1.0, se) > 4.0e18)
        --se;
    const double worstE2 = static_cast<double>(pb.nTotal) * 150.0 * 150.0 * pb.maxWeight * pb.maxWeight * (kerma ? 2500.0 : 1.0);
    int s2 = 16;
    while (s2 > -20 && worstE2 * std::ldexp(1.0, s2) > 4.0e18)
        --s2;
    c->scaleE = static_cast<float>(std::ldexp(1.0, se));
    c->scaleE2 = static_cast<float>(std::ldexp(1.0, s2));
}

// ---- nested CTDI calibration (CTSpiralBeam/CTSequentialBeam::calibrationFactor, recalled) ----
// A PMMA cylinder (diameter = beam.ctdi_diameter, length 15 cm) with five 1.31 cm air holes, voxelised at
// 0.2 x 0.2 x 0.5 cm; one axial rotation of the same tube/bowtie/collimation; air kerma in the central 10 cm
// of the holes from a collision estimator; CTDIw = 1/3 centre + 2/3 periphery.
void buildCtdiPhantom(double diameter, CtdiPhantom& ph)
{
    const double sxy = 0.2, sz = 0.5;
    const int nxy = static_cast<int>(std::ceil((diameter + 4.0) / sxy / 2.0)) * 2;
    const int nz = 30; // 15 cm
    ph.diameter = diameter;
    ph.dim[0] = ph.dim[1] = nxy;
    ph.dim[2] = nz;
    ph.spacing[0] = ph.spacing[1] = sxy;
    ph.spacing[2] = sz;
    const size_t n = static_cast<size_t>(nxy) * nxy * nz;
    ph.density.assign(n, 0.0);
    ph.material.assign(n, 0);
    ph.hole.assign(n, -1);
    for (size_t& v : ph.holeCount)
        v = 0;
    const double airRho = nistFind("Air, Dry (near sea level)")->density;
    const double pmmaRho = nistFind("Polymethyl Methacralate (Lucite, Perspex)")->density;
    const double R = 0.5 * diameter, rh = 0.655, off = R - 1.0;
    const double hx[5] = { 0, off, -off, 0, 0 }, hy[5] = { 0, 0, 0, off, -off };
    for (int k = 0; k < nz; ++k) {
        const double z = (k + 0.5) * sz - 0.5 * nz * sz;
        for (int j = 0; j < nxy; ++j) {
            const double y = (j + 0.5) * sxy - 0.5 * nxy * sxy;
            for (int i = 0; i < nxy; ++i) {
                const double x = (i + 0.5) * sxy - 0.5 * nxy * sxy;
                const size_t idx = (static_cast<size_t>(k) * nxy + j) * nxy + i;
                ph.density[idx] = airRho;
                if (x * x + y * y <= R * R) {
                    ph.density[idx] = pmmaRho;
                    ph.material[idx] = 1;
                    for (int h = 0; h < 5; ++h) {
                        const double ddx = x - hx[h], ddy = y - hy[h];
                        if (ddx * ddx + ddy * ddy <= rh * rh) {
                            ph.density[idx] = airRho;
                            ph.material[idx] = 0;
                            if (std::fabs(z) <= 5.0) {
                                ph.material[idx] = 2;
                                ph.hole[idx] = static_cast<signed char>(h);
                            }
                        }
                    }
                }
            }
        }
    }
    for (size_t i = 0; i < n; ++i)
        if (ph.hole[i] >= 0)
            ++ph.holeCount[ph.hole[i]];
    ph.mats = { Material::byNistName("Air, Dry (near sea level)"), Material::byNistName("Polymethyl Methacralate (Lucite, Perspex)"),
        Material::byNistName("Air, Dry (near sea level)") };
}

bool isCtBeam(int type)
{
    return type == DXB_BEAM_CT_SPIRAL || type == DXB_BEAM_CT_SPIRAL_DUAL || type == DXB_BEAM_CT_SEQUENTIAL;
}

// The nested run.  The phantom is built once per diameter and stays resident on every device of the context; the
// histories of the calibration beam are sharded over the devices like those of any beam, each device sums the kerma
// tallies of the five holes itself (integers), and the host adds five numbers per device.  The run uses its own
// Philox stream (beam key ^ DXB_CALIBRATION_KEY_XOR).
int ctCalibration(dxb_ctx* c, const dxb_beam_desc& b, int mode, double* factorOut, double* msOut)
{
    const double diameter = b.ctdi_diameter > 0 ? b.ctdi_diameter : 32.0;
    if (!c->ctdi || c->ctdi->diameter != diameter) {
        c->ctdi = std::make_unique<CtdiPhantom>();
        buildCtdiPhantom(diameter, *c->ctdi);
    }
    const CtdiPhantom& ph = *c->ctdi;
    const size_t nPh = static_cast<size_t>(ph.dim[0]) * ph.dim[1] * ph.dim[2];
    // shards of the nested beam: the devices of this context, or - one process per GPU with the library-managed exchange -
    // the ranks of the job, whose five sums travel through IPC mailboxes (exchange.cu).  A context that is sharded by the
    // caller alone (dxb_set_history_range without dxb_exchange_import) runs the whole nested beam itself: it is
    // deterministic, so every rank derives the same factor without a broadcast.
    const uint64_t nDev = c->devs.size();
    const bool overRanks = c->ipc && c->exchanging && c->world > 1 && c->devs[0]->mailbox.p;
    const uint64_t shardBase = overRanks ? c->rank : 0, shardWorld = overRanks ? c->world : nDev;
    for (uint64_t i = 0; i < nDev; ++i) {
        DeviceState& d = *c->devs[i];
        CUDA_TRY(c, cudaSetDevice(d.device));
        if (!d.ctdi || d.ctdiDiameter != diameter) {
            d.ctdi = std::make_unique<World>();
            d.ctdi->device = d.device;
            int rc = uploadTables(c, *d.ctdi, ph.mats, d.stream);
            if (rc != DXB_OK)
                return rc;
            const bool ipc = c->ipc;
            c->ipc = false; // the phantom's tally buffer is never shared
            rc = uploadGrid(c, *d.ctdi, ph.dim, ph.spacing, ph.density.data(), ph.material.data(), d.stream, 0, nPh, true);
            c->ipc = ipc;
            if (rc != DXB_OK)
                return rc;
            CUDA_TRY(c, d.ctdiHole.upload(ph.hole, d.device, d.stream));
            CUDA_TRY(c, d.holeSums.alloc(8, d.device));
            CUDA_TRY(c, cudaStreamSynchronize(d.stream));
            d.ctdiDiameter = diameter;
        }
    }

    // the internal axial beam: same tube(s), bowtie(s), collimation, SDD and FOV; 1 degree steps, no AEC
    dxb_beam_desc cb = b;
    cb.type = DXB_BEAM_CTDI;
    cb.position[0] = cb.position[1] = cb.position[2] = 0;
    cb.direction[0] = cb.direction[1] = 0;
    cb.direction[2] = 1;
    cb.step_angle = kPi / 180.0;
    cb.start_angle = 0;
    cb.n_slices = 1;
    cb.aec.n = 0;
    cb.organ_aec.use_filter = 0;
    if (b.type != DXB_BEAM_CT_SPIRAL_DUAL)
        cb.spectrum[1].n = 0;
    const uint64_t nExp = beamNumberOfExposures(cb);
    cb.particles_per_exposure = std::max<uint64_t>(1, c->calibHistories / nExp);

    PreparedBeam pb;
    int rc = prepareBeam(c, cb, pb);
    if (rc != DXB_OK)
        return rc;
    const float saveE = c->scaleE, saveE2 = c->scaleE2;
    const uint64_t saveKey = c->runKey;
    chooseScales(c, pb, true);
    c->runKey = c->beamKey ^ DXB_CALIBRATION_KEY_XOR;
    std::vector<TransportResult> tr(nDev);
    for (uint64_t i = 0; i < nDev && rc == DXB_OK; ++i) {
        DeviceState& d = *c->devs[i];
        World& w = *d.ctdi;
        if (cudaSetDevice(d.device) != cudaSuccess || cudaMemsetAsync(w.tally.p, 0, w.nvox * 4 * sizeof(unsigned long long), d.stream) != cudaSuccess
            || cudaMemsetAsync(d.holeSums.p, 0, 8 * sizeof(unsigned long long), d.stream) != cudaSuccess) {
            cudaGetLastError();
            rc = fail(c, DXB_ECUDA, "calibration: clearing the phantom tallies failed");
            break;
        }
        rc = uploadBeam(c, d, pb);
        if (rc == DXB_OK)
            rc = runOnDevice(c, d, w, pb, mode, true, 2, shardBase + i, shardWorld, nullptr, true, &tr[i]);
        if (rc == DXB_OK) {
            launchHoleSums(w.tally.p, d.ctdiHole.p, w.nvox, d.holeSums.p, d.stream);
            if (cudaGetLastError() != cudaSuccess)
                rc = fail(c, DXB_ECUDA, "calibration: hole sums");
        }
    }
    unsigned long long sums[5] = { 0, 0, 0, 0, 0 };
    double msMax = 0;
    for (uint64_t i = 0; i < nDev && rc == DXB_OK; ++i) {
        DeviceState& d = *c->devs[i];
        rc = collectStats(c, d, tr[i]);
        if (rc != DXB_OK)
            break;
        unsigned long long h[5];
        if (overRanks) {
            rc = mgShareHoleSums(c, d.holeSums.p, h);
            if (rc != DXB_OK)
                break;
        } else if (cudaMemcpy(h, d.holeSums.p, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) {
            cudaGetLastError();
            rc = fail(c, DXB_ECUDA, "calibration: reading the hole sums failed");
            break;
        }
        for (int k = 0; k < 5; ++k)
            sums[k] += h[k];
        msMax = std::max(msMax, tr[i].ms);
    }
    const double invE = 1.0 / c->scaleE;
    c->scaleE = saveE;
    c->scaleE2 = saveE2;
    c->runKey = saveKey;
    cudaSetDevice(c->devs[0]->device);
    if (rc != DXB_OK)
        return rc;
    if (msOut)
        *msOut = msMax;
    const double vol = ph.spacing[0] * ph.spacing[1] * ph.spacing[2];
    double kerma[5];
    for (int h = 0; h < 5; ++h)
        kerma[h] = ph.holeCount[h] ? static_cast<double>(sums[h]) * invE / (static_cast<double>(ph.holeCount[h]) * vol) : 0.0; // keV/g, mean over the 10 cm chamber length
    // CTDI100 = (1/NT) * integral_{-5}^{5} K dz = Kmean * 10 cm / collimation
    const double ctdi100c = kerma[0] * 10.0 / b.collimation;
    const double ctdi100p = 0.25 * (kerma[1] + kerma[2] + kerma[3] + kerma[4]) * 10.0 / b.collimation;
    const double ctdiwSim = ctdi100c / 3.0 + 2.0 * ctdi100p / 3.0; // keV/g for pb.nTotal histories in ONE rotation
    if (!(ctdiwSim > 0))
        return fail(c, DXB_EINVAL, "CTDI calibration run scored nothing");
    const double perHistory = ctdiwSim / static_cast<double>(pb.nTotal);
    // histories the main beam spends per rotation
    const double step = std::fabs(b.step_angle) > 0 ? std::fabs(b.step_angle) : kPi / 180.0;
    double perRot = static_cast<double>(b.particles_per_exposure) * (2.0 * kPi / step);
    if (b.type == DXB_BEAM_CT_SPIRAL_DUAL)
        perRot *= 2.0;
    const double target = b.type == DXB_BEAM_CT_SEQUENTIAL ? b.ctdi : b.ctdi * b.pitch; // CTDIw [mGy] per rotation
    *factorOut = target / (perHistory * perRot);
    return DXB_OK;
}

int initDevice(dxb_ctx* c, DeviceState& d)
{
    CUDA_TRY(c, cudaSetDevice(d.device));
    CUDA_TRY(c, cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
    CUDA_TRY(c, cudaEventCreate(&d.evStart));
    CUDA_TRY(c, cudaEventCreate(&d.evTransport));
    CUDA_TRY(c, cudaEventCreate(&d.evEnd));
    CUDA_TRY(c, d.counters.alloc(32, d.device));
    CUDA_TRY(c, cudaMemset(d.counters.p, 0, 32 * sizeof(unsigned long long)));
    d.world.device = d.device;
    return DXB_OK;
}

} // namespace dxb

// ============================================================================ C ABI
extern "C" {

int dxb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int dxb_create(dxb_ctx** out, const int* cuda_devices, int n_devices)
{
    if (!out)
        return DXB_EINVAL;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return DXB_ECUDA; // no CPU fallback
    }
    auto c = std::make_unique<dxb_ctx>();
    std::vector<int> devices;
    if (!cuda_devices || n_devices <= 0) {
        int cur = 0;
        cudaGetDevice(&cur);
        devices.push_back(cur);
    } else {
        for (int i = 0; i < n_devices; ++i) {
            if (cuda_devices[i] < 0 || cuda_devices[i] >= count)
                return DXB_EINVAL;
            devices.push_back(cuda_devices[i]);
        }
    }
    for (size_t i = 0; i < devices.size(); ++i)
        for (size_t j = 0; j < i; ++j)
            if (devices[i] == devices[j])
                return DXB_EINVAL;
    for (int dev : devices) {
        auto d = std::make_unique<DeviceState>();
        d->device = dev;
        if (initDevice(c.get(), *d) != DXB_OK)
            return DXB_ECUDA;
        c->devs.push_back(std::move(d));
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, devices[0]) == cudaSuccess)
        c->smCount = prop.multiProcessorCount;
    setLaunchSmCount(c->smCount);
    // several devices: peer access all-to-all, exchange streams and events (exchange.cu)
    if (c->devs.size() > 1 && mgInit(c.get()) != DXB_OK) {
        mgDestroy(c.get());
        return DXB_ECUDA;
    }
    cudaSetDevice(devices[0]);
    *out = c.release();
    return DXB_OK;
}

void dxb_destroy(dxb_ctx* c)
{
    if (!c)
        return;
    c->workers.reset();
    mgDestroy(c);
    for (auto& d : c->devs) {
        cudaSetDevice(d->device);
        cudaStreamSynchronize(d->stream);
        if (d->evStart)
            cudaEventDestroy(d->evStart);
        if (d->evTransport)
            cudaEventDestroy(d->evTransport);
        if (d->evEnd)
            cudaEventDestroy(d->evEnd);
        if (d->ownStream && d->stream)
            cudaStreamDestroy(d->stream);
    }
    delete c;
}

const char* dxb_last_error(const dxb_ctx* c) { return c ? c->error.c_str() : "null context"; }

int dxb_set_materials(dxb_ctx* c, uint32_t n, const dxb_material* const* materials)
{
    if (!c || n == 0 || n > 255 || !materials)
        return fail(c, DXB_EINVAL, "set_materials: need 1..255 materials");
    std::vector<std::shared_ptr<Material>> mats;
    for (uint32_t i = 0; i < n; ++i) {
        if (!materials[i] || !materials[i]->m)
            return fail(c, DXB_EMATERIAL, "set_materials: null material");
        mats.push_back(materials[i]->m);
    }
    c->materials = mats;
    for (auto& d : c->devs) {
        CUDA_TRY(c, cudaSetDevice(d->device));
        int rc = uploadTables(c, d->world, mats, d->stream);
        if (rc != DXB_OK)
            return rc;
        d->world.hasGrid = false; // majorant depends on the tables
    }
    return DXB_OK;
}

int dxb_set_grid(dxb_ctx* c, const uint64_t dim[3], const double spacing_cm[3], const double* density, const uint8_t* material)
{
    if (!c || !dim || !spacing_cm || !density || !material)
        return fail(c, DXB_EINVAL, "set_grid: null argument");
    if (c->materials.empty())
        return fail(c, DXB_ESTATE, "set_grid: call dxb_set_materials first");
    const uint64_t n = dim[0] * dim[1] * dim[2];
    if (n == 0 || n >= (1ull << 32) || dim[0] >= (1u << 20) || dim[1] >= (1u << 20) || dim[2] >= (1u << 20))
        return fail(c, DXB_EINVAL, "set_grid: bad dimensions");
    for (int i = 0; i < 3; ++i)
        if (!(spacing_cm[i] > 0))
            return fail(c, DXB_EINVAL, "set_grid: spacing must be positive");
    c->tallyValid = false;
    if (c->devs.size() > 1) // every device uploads its slab; the dose score is distributed the same way (exchange.cu)
        return mgSetGrid(c, dim, spacing_cm, density, material);
    int rc = mgFlush(c);
    if (rc != DXB_OK)
        return rc;
    DeviceState& d0 = *c->devs[0];
    CUDA_TRY(c, cudaSetDevice(d0.device));
    rc = uploadGrid(c, d0.world, dim, spacing_cm, density, material, d0.stream, 0, n, true);
    if (rc != DXB_OK)
        return rc;
    if (c->exchanging && (rc = mgPrepareExchange(c)) != DXB_OK)
        return rc;
    CUDA_TRY(c, cudaSetDevice(d0.device));
    CUDA_TRY(c, d0.dose.alloc(n, d0.device));
    CUDA_TRY(c, d0.variance.alloc(n, d0.device));
    CUDA_TRY(c, d0.events.alloc(n, d0.device));
    c->tallyValid = false;
    return dxb_clear_dose(c);
}

int dxb_set_grid_sharded(dxb_ctx* c, const uint64_t dim[3], const double spacing_cm[3], const double* density, const uint8_t* material,
    uint64_t voxel_begin, uint64_t voxel_end)
{
    if (!c || !dim || !spacing_cm || !density || !material)
        return fail(c, DXB_EINVAL, "set_grid_sharded: null argument");
    if (c->materials.empty())
        return fail(c, DXB_ESTATE, "set_grid_sharded: call dxb_set_materials first");
    if (c->devs.size() != 1)
        return fail(c, DXB_ESTATE, "set_grid_sharded: one device per context (one process per GPU)");
    const uint64_t n = dim[0] * dim[1] * dim[2];
    if (n == 0 || n >= (1ull << 32) || dim[0] >= (1u << 20) || dim[1] >= (1u << 20) || dim[2] >= (1u << 20))
        return fail(c, DXB_EINVAL, "set_grid_sharded: bad dimensions");
    for (int i = 0; i < 3; ++i)
        if (!(spacing_cm[i] > 0))
            return fail(c, DXB_EINVAL, "set_grid_sharded: spacing must be positive");
    if (voxel_begin > voxel_end || voxel_end > n)
        return fail(c, DXB_EINVAL, "set_grid_sharded: voxel range outside the grid");
    DeviceState& d0 = *c->devs[0];
    CUDA_TRY(c, cudaSetDevice(d0.device));
    int rc = mgFlush(c);
    if (rc != DXB_OK)
        return rc;
    rc = uploadGrid(c, d0.world, dim, spacing_cm, density, material, d0.stream, voxel_begin, voxel_end, false);
    if (rc != DXB_OK)
        return rc;
    if (c->exchanging && (rc = mgPrepareExchange(c)) != DXB_OK)
        return rc;
    CUDA_TRY(c, d0.dose.alloc(n, d0.device));
    CUDA_TRY(c, d0.variance.alloc(n, d0.device));
    CUDA_TRY(c, d0.events.alloc(n, d0.device));
    c->tallyValid = false;
    const size_t nn = n;
    CUDA_TRY(c, cudaMemsetAsync(d0.dose.p, 0, nn * sizeof(double), d0.stream));
    CUDA_TRY(c, cudaMemsetAsync(d0.variance.p, 0, nn * sizeof(double), d0.stream));
    CUDA_TRY(c, cudaMemsetAsync(d0.events.p, 0, nn * sizeof(unsigned long long), d0.stream));
    return DXB_OK;
}

int dxb_grid_buffers(dxb_ctx* c, void** voxels, void** max_density_bits, uint64_t* n_voxels)
{
    if (!c || c->devs.empty() || c->devs[0]->world.nvox == 0 || !c->devs[0]->world.voxels.p)
        return fail(c, DXB_ESTATE, "grid_buffers: no grid");
    World& w = c->devs[0]->world;
    if (voxels)
        *voxels = w.voxels.p;
    if (max_density_bits)
        *max_density_bits = w.maxDensityBits.p;
    if (n_voxels)
        *n_voxels = w.nvox;
    return DXB_OK;
}

int dxb_finish_grid(dxb_ctx* c)
{
    if (!c || c->devs.size() != 1 || c->devs[0]->world.nvox == 0 || !c->devs[0]->world.voxels.p)
        return fail(c, DXB_ESTATE, "finish_grid: call dxb_set_grid_sharded first");
    DeviceState& d0 = *c->devs[0];
    CUDA_TRY(c, cudaSetDevice(d0.device));
    return finishGrid(c, d0.world, d0.stream);
}

int dxb_set_grid_center(dxb_ctx* c, const double center_cm[3])
{
    if (!c || !center_cm)
        return DXB_EINVAL;
    for (auto& d : c->devs)
        for (int i = 0; i < 3; ++i)
            d->world.center[i] = center_cm[i];
    return DXB_OK;
}

int dxb_clear_dose(dxb_ctx* c)
{
    if (!c || c->devs.empty() || !c->devs[0]->world.hasGrid)
        return fail(c, DXB_ESTATE, "clear_dose: no grid");
    int rc = mgFlush(c);
    if (rc != DXB_OK)
        return rc;
    const size_t n = c->devs[0]->world.nvox;
    for (auto& d : c->devs) {
        if (!d->dose.p)
            continue;
        CUDA_TRY(c, cudaSetDevice(d->device));
        CUDA_TRY(c, cudaMemsetAsync(d->dose.p, 0, n * sizeof(double), d->stream));
        CUDA_TRY(c, cudaMemsetAsync(d->variance.p, 0, n * sizeof(double), d->stream));
        CUDA_TRY(c, cudaMemsetAsync(d->events.p, 0, n * sizeof(unsigned long long), d->stream));
    }
    for (auto& d : c->devs) {
        CUDA_TRY(c, cudaSetDevice(d->device));
        CUDA_TRY(c, cudaStreamSynchronize(d->stream));
    }
    CUDA_TRY(c, cudaSetDevice(c->devs[0]->device));
    return DXB_OK;
}

int dxb_set_seed(dxb_ctx* c, uint64_t seed)
{
    if (!c)
        return DXB_EINVAL;
    c->seed = seed;
    c->beamCounter = 0; // the next beam runs on key `seed` itself
    return DXB_OK;
}

uint64_t dxb_last_beam_key(const dxb_ctx* c) { return c ? c->beamKey : 0; }

int dxb_set_history_range(dxb_ctx* c, uint64_t rank, uint64_t world)
{
    if (!c || world == 0 || rank >= world || world > 65535)
        return fail(c, DXB_EINVAL, "set_history_range: need rank < world");
    c->rank = rank;
    c->world = world;
    return DXB_OK;
}

uint64_t dxb_shard_local_count(uint64_t n_total, uint64_t rank, uint64_t world)
{
    return world == 0 ? 0 : localCount(n_total, rank, world);
}

uint64_t dxb_shard_history_id(uint64_t local_index, uint64_t rank, uint64_t world)
{
    // the same arithmetic as the refill phase of transportKernel
    return ((local_index / kShardBlock) * world + rank) * kShardBlock + (local_index % kShardBlock);
}

int dxb_set_calibration_histories(dxb_ctx* c, uint64_t n)
{
    if (!c || n == 0)
        return DXB_EINVAL;
    c->calibHistories = n;
    return DXB_OK;
}

int dxb_set_stream(dxb_ctx* c, void* cuda_stream)
{
    if (!c || c->devs.empty())
        return DXB_EINVAL;
    DeviceState& d = *c->devs[0];
    cudaSetDevice(d.device);
    cudaStreamSynchronize(d.stream);
    if (d.ownStream && d.stream)
        cudaStreamDestroy(d.stream);
    if (cuda_stream) {
        d.stream = static_cast<cudaStream_t>(cuda_stream);
        d.ownStream = false;
    } else {
        if (cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking) != cudaSuccess)
            return fail(c, DXB_ECUDA, "stream create failed");
        d.ownStream = true;
    }
    return DXB_OK;
}

int dxb_set_option(dxb_ctx* c, const char* key, double value)
{
    if (!c || !key)
        return DXB_EINVAL;
    const std::string k(key);
    if (k == "batch_histories") {
        if (value < kShardBlock)
            return DXB_EINVAL;
        c->opt.batch = static_cast<uint64_t>(value) / kShardBlock * kShardBlock; // launches start on shard-block boundaries
    } else if (k == "threads_per_block") {
        const int t = static_cast<int>(value);
        if (t < 32 || t > 256 || (t % 32))
            return DXB_EINVAL;
        c->opt.threads = t;
    } else if (k == "blocks_per_sm") {
        c->opt.blocksPerSm = static_cast<int>(value);
    } else if (k == "table_in_smem") {
        c->opt.tableInSmem = value != 0;
    } else if (k == "slots_per_lane") {
        const int v = static_cast<int>(value);
        if (v != 0 && v != 2 && v != 3 && v != 4 && v != 6)
            return fail(c, DXB_EINVAL, "slots_per_lane must be 0 (register kernel), 2, 3, 4 or 6");
        c->opt.slots = v;
    } else if (k == "smem_pad_kb") {
        c->opt.smemPadKb = static_cast<int>(value);
    } else if (k == "pool_slots") {
        const int v = static_cast<int>(value);
        if (v != 0 && v != 6 && v != 8 && v != 12 && v != 16)
            return fail(c, DXB_EINVAL, "pool_slots must be 0 (off), 6, 8, 12 or 16");
        c->opt.poolSlots = v;
    } else if (k == "step_pairs") {
        c->opt.stepPairs = static_cast<int>(value);
    } else if (k == "pool_threads") {
        c->opt.poolThreads = static_cast<int>(value);
    } else if (k == "pool_min_blocks") {
        c->opt.poolMinBlocks = static_cast<int>(value);
    } else if (k == "step_quad") {
        c->opt.stepQuad = static_cast<int>(value);
    } else if (k == "brick_filter") {
        c->opt.brickFilter = value != 0; // (the table is built with the grid; switching the filter off takes effect at once)
    } else if (k == "brick_voxels") {
        const int v = static_cast<int>(value);
        if (v < 2 || v > 256 || (v & (v - 1)))
            return fail(c, DXB_EINVAL, "brick_voxels must be a power of two in 2..256");
        c->opt.brickVoxels = v;
    } else if (k == "dense_box") {
        const int v = static_cast<int>(value);
        if (v < -1 || v > 1)
            return fail(c, DXB_EINVAL, "dense_box must be -1 (auto), 0 (off) or 1 (on)");
        c->opt.denseBox = v; // the box is built with the grid unless the option is 0 then; switching it off takes effect at once
    } else if (k == "dense_theta") {
        if (!(value > 0 && value < 1))
            return fail(c, DXB_EINVAL, "dense_theta must be in (0, 1)");
        c->opt.denseTheta = value; // takes effect with the next dxb_set_grid
    } else if (k == "local_majorant") {
        const int v = static_cast<int>(value);
        if (v < -1 || v > 1)
            return fail(c, DXB_EINVAL, "local_majorant must be -1 (auto), 0 (off) or 1 (on)");
        c->opt.localMajorant = v; // takes effect with the next dxb_set_grid (the table is built with the grid)
    } else if (k == "slab_cm") {
        if (!(value > 0))
            return DXB_EINVAL;
        c->opt.slabCm = value;
    } else if (k == "diag") {
        c->opt.diag = static_cast<int>(value);
    } else if (k == "service_warps") {
        c->opt.serviceWarps = static_cast<int>(value);
    } else if (k == "interact_bias") {
        c->opt.interactBias = static_cast<int>(value);
    } else if (k == "refill_threshold") {
        c->opt.refillThreshold = static_cast<int>(value);
    } else if (k == "interact_threshold") {
        c->opt.interactThreshold = static_cast<int>(value);
    } else if (k == "rayleigh_threshold") {
        c->opt.rayleighThreshold = static_cast<int>(value);
    } else {
        return fail(c, DXB_EINVAL, "unknown option " + k);
    }
    return DXB_OK;
}

int dxb_run_transport(dxb_ctx* c, const dxb_beam_desc* beam, int physics_mode, dxb_progress* progress)
{
    if (!c || !beam)
        return DXB_EINVAL;
    if (physics_mode < 0 || physics_mode > 2)
        return fail(c, DXB_EINVAL, "physics_mode must be 0, 1 or 2");
    if (c->devs.empty() || !c->devs[0]->world.hasGrid || !c->devs[0]->world.hasTables)
        return fail(c, DXB_ESTATE, "run: set materials and grid first");
    const auto hostEntry = std::chrono::steady_clock::now();
    // the same beam as last time (a job usually repeats a beam, or runs it again after a parameter study): the exposure
    // expansion (thousands of poses), the alias tables and their upload are reused
    const uint64_t bh = beamHash(*beam);
    int rc = DXB_OK;
    if (!c->lastBeam || c->lastBeam->hash != bh) {
        auto fresh = std::make_unique<PreparedBeam>();
        rc = prepareBeam(c, *beam, *fresh);
        if (rc != DXB_OK)
            return rc;
        fresh->hash = bh;
        c->lastBeam = std::move(fresh);
    }
    const PreparedBeam& pb = *c->lastBeam;
    chooseScales(c, pb, false);
    const uint64_t nDev = c->devs.size();
    const uint64_t effWorld = c->world * nDev;
    if (progress) {
        uint64_t mine = 0;
        for (uint64_t i = 0; i < nDev; ++i)
            mine += std::min(localCount(pb.nTotal, c->rank * nDev + i, effWorld), pb.nTotal);
        progress->total.store(std::min(mine, pb.nTotal));
        progress->done.store(0);
        progress->start_ns.store(std::chrono::duration_cast<std::chrono::nanoseconds>(
            std::chrono::steady_clock::now().time_since_epoch()).count());
    }
    c->stats = dxb_run_stats {};
    c->tallyValid = false;
    // Every beam gets its own Philox key (the reference's driver calls transport() once per beam on one world,
    // R:src/libopendxmc/simulationpipeline.cpp:161-167: identical streams would make identical beams add no information
    // while their variances are summed as if independent).  All ranks / devices of a job count the same beams, so the
    // key - and with it the result - stays independent of the GPU count.
    c->beamKey = c->seed + c->beamCounter * DXB_BEAM_KEY_STRIDE;
    c->runKey = c->beamKey;
    ++c->beamCounter;
    std::vector<TransportResult> results(nDev);
    // Hand over the tally buffer, upload the beam and launch - one host thread per device, so that device 7 does not start
    // a launch sequence later than device 0 (the sequence is ~40 driver calls per device).
    auto launchOn = [&](size_t i) -> int {
        DeviceState& d = *c->devs[i];
        CUDA_TRY(c, cudaSetDevice(d.device));
        if (c->exchanging) {
            // double-buffered tallies: the buffer of this beam is cleared by the exchange of the beam before
            if (!d.needsClear[d.world.cur])
                CUDA_TRY(c, cudaStreamWaitEvent(d.stream, d.evBufReady[d.world.cur], 0));
            else // the last beam was never finished (its tallies are dropped, as without an exchange)
                CUDA_TRY(c, cudaMemsetAsync(d.world.tallyCur(), 0, d.world.nvox * 4 * sizeof(unsigned long long), d.stream));
        } else {
            CUDA_TRY(c, cudaMemsetAsync(d.world.tallyCur(), 0, d.world.nvox * 4 * sizeof(unsigned long long), d.stream));
        }
        int r = uploadBeam(c, d, pb);
        if (r != DXB_OK)
            return r;
        r = runOnDevice(c, d, d.world, pb, physics_mode, false, -1, c->rank * nDev + i, effWorld, progress, nDev > 1, &results[i]);
        if (r != DXB_OK)
            return r;
        if (c->exchanging) {
            d.needsClear[d.world.cur] = true; // scored into, not yet exchanged
            CUDA_TRY(c, cudaEventRecord(d.evTransportDone[d.world.cur], d.stream));
        }
        return DXB_OK;
    };
    c->exchanged = false;
    const auto hostT0 = std::chrono::steady_clock::now();
    rc = nDev > 1 ? overDevices(c, launchOn) : launchOn(0);
    if (rc != DXB_OK)
        return rc;
    const auto hostT1 = std::chrono::steady_clock::now();
    auto collectOn = [&](size_t i) -> int { return collectStats(c, *c->devs[i], results[i]); };
    rc = nDev > 1 ? overDevices(c, collectOn) : collectOn(0);
    if (rc != DXB_OK)
        return rc;
    bool cancelled = false;
    double msMax = 0;
    for (uint64_t i = 0; i < nDev; ++i) {
        cancelled = cancelled || results[i].cancelled;
        msMax = std::max(msMax, results[i].ms);
        c->stats.steps += results[i].stats[0];
        c->stats.interactions += results[i].stats[1];
        c->stats.deposits += results[i].stats[2];
        c->stats.energy_emitted_kev += static_cast<double>(results[i].stats[3]) / 65536.0;
        c->stats.histories += results[i].stats[4];
        c->stats.hops += results[i].stats[5];
        c->stats.voxel_fetches += results[i].stats[6];
        c->stats.kernel_launches += results[i].launches;
    }
    c->stats.transport_ms = msMax;
    if (traceHost()) {
        const auto hostT2 = std::chrono::steady_clock::now();
        const auto us = [](auto a, auto b) { return std::chrono::duration_cast<std::chrono::microseconds>(b - a).count(); };
        std::fprintf(stderr, "[dxb host] run_transport: since last return %lld us, prepare %lld us, launch %lld us, wait+collect %lld us (kernel %.3f ms)\n",
            static_cast<long long>(c->lastReturn.time_since_epoch().count() ? us(c->lastReturn, hostEntry) : -1),
            static_cast<long long>(us(hostEntry, hostT0)), static_cast<long long>(us(hostT0, hostT1)),
            static_cast<long long>(us(hostT1, hostT2)), msMax);
    }
    c->lastReturn = std::chrono::steady_clock::now();
    if (progress)
        progress->done.store(progress->total.load());
    if (cancelled)
        return fail(c, DXB_ECANCELLED, "cancelled");
    if (c->exchanging) {
        // before the peers may clear the buffer of the PREVIOUS beam (at the next exchange; with one process per GPU the
        // caller's barrier tells them), this participant's pulls from it must have finished - they ran under the
        // transport kernels just awaited
        for (auto& dp : c->devs) {
            DeviceState& d = *dp;
            const int prev = d.world.cur ^ 1;
            if (d.pullsPending[prev]) {
                CUDA_TRY(c, cudaSetDevice(d.device));
                CUDA_TRY(c, cudaEventSynchronize(d.evPullsDone[prev]));
                d.pullsPending[prev] = false;
            }
        }
        CUDA_TRY(c, cudaSetDevice(c->devs[0]->device));
    }
    c->tallyValid = true;
    return DXB_OK;
}

int dxb_finish_beam(dxb_ctx* c, const dxb_beam_desc* beam, int physics_mode, int use_beam_calibration, double* factor_out)
{
    if (!c || !beam)
        return DXB_EINVAL;
    if (!c->tallyValid)
        return fail(c, DXB_ESTATE, "finish_beam: no tallies (call dxb_run_transport first; a beam is finished once)");
    DeviceState& d0 = *c->devs[0];
    double factor = kKeVperGramToMilliGray;
    double calibMs = 0;
    if (use_beam_calibration) {
        if (isCtBeam(beam->type)) {
            int rc = ctCalibration(c, *beam, physics_mode, &factor, &calibMs);
            if (rc != DXB_OK)
                return rc;
        } else {
            factor = beamAnalyticCalibration(*beam);
        }
    }
    CUDA_TRY(c, cudaSetDevice(d0.device));
    if (c->exchanging) {
        // several participants: the exchange (peer pulls, slab reduce -> dose, clear) is enqueued on the exchange streams and
        // runs underneath the next beam's transport; read-out calls wait for it (exchange.cu)
        const int rc = mgEnqueueExchange(c, factor);
        if (rc != DXB_OK)
            return rc;
    } else {
        World& w = d0.world;
        const double vol = w.spacing[0] * w.spacing[1] * w.spacing[2];
        launchEnergyToDose(w.tallyCur(), w.voxels.p, d0.dose.p, d0.variance.p, d0.events.p, w.nvox, 1.0 / c->scaleE, 1.0 / c->scaleE2,
            factor, vol, d0.stream);
        CUDA_TRY(c, cudaGetLastError());
        CUDA_TRY(c, cudaEventRecord(d0.evEnd, d0.stream));
        CUDA_TRY(c, cudaStreamSynchronize(d0.stream));
    }
    // a second finish of the same beam would add its dose, variance and events once more
    c->tallyValid = false;
    c->stats.calibration_factor = factor;
    c->stats.calibration_ms = calibMs;
    if (factor_out)
        *factor_out = factor;
    return DXB_OK;
}

int dxb_set_tally_storage(dxb_ctx* c, void* device_ptr, uint64_t n_words)
{
    if (!c || c->devs.empty() || !c->devs[0]->world.hasGrid)
        return fail(c, DXB_ESTATE, "set_tally_storage: set the grid first");
    if (c->devs.size() != 1 || c->exchanging)
        return fail(c, DXB_ESTATE, "set_tally_storage: single-device contexts without a library-managed exchange only");
    DeviceState& d0 = *c->devs[0];
    World& w = d0.world;
    CUDA_TRY(c, cudaSetDevice(d0.device));
    CUDA_TRY(c, cudaStreamSynchronize(d0.stream));
    if (!device_ptr) {
        if (!w.tally.owned) {
            w.tally.release();
            CUDA_TRY(c, w.tally.alloc(w.nvox * 4, w.device));
        }
    } else {
        if (n_words != w.nvox * 4)
            return fail(c, DXB_EINVAL, "set_tally_storage: the buffer must hold 4 x 64-bit words per voxel");
        if (reinterpret_cast<uintptr_t>(device_ptr) % 32)
            return fail(c, DXB_EINVAL, "set_tally_storage: the buffer must be 32-byte aligned");
        w.tally.adopt(static_cast<unsigned long long*>(device_ptr), w.nvox * 4, w.device);
    }
    CUDA_TRY(c, cudaMemsetAsync(w.tally.p, 0, w.nvox * 4 * sizeof(unsigned long long), d0.stream));
    CUDA_TRY(c, cudaStreamSynchronize(d0.stream));
    c->tallyValid = false;
    return DXB_OK;
}

int dxb_finish_beam_sharded(dxb_ctx* c, const dxb_beam_desc* beam, int physics_mode, int use_beam_calibration, const void* multicast_tally,
    const void* const* peer_tallies, int n_peers, uint64_t voxel_begin, uint64_t voxel_end, double* factor_out)
{
    if (!c || !beam)
        return DXB_EINVAL;
    if (!c->tallyValid)
        return fail(c, DXB_ESTATE, "finish_beam_sharded: no tallies (call dxb_run_transport first)");
    if (c->devs.size() != 1 || c->exchanging)
        return fail(c, DXB_ESTATE, "finish_beam_sharded: single-device contexts without a library-managed exchange only");
    DeviceState& d0 = *c->devs[0];
    World& w = d0.world;
    if (voxel_begin > voxel_end || voxel_end > w.nvox || n_peers < 0 || n_peers > 63 || (n_peers > 0 && !peer_tallies && !multicast_tally))
        return fail(c, DXB_EINVAL, "finish_beam_sharded: bad slab or peer list");
    double factor = kKeVperGramToMilliGray;
    double calibMs = 0;
    if (use_beam_calibration) {
        // every rank derives the factor itself: the nested CTDI run is deterministic (Philox + integer tallies), so all
        // ranks get the same bits without a broadcast
        if (isCtBeam(beam->type)) {
            int rc = ctCalibration(c, *beam, physics_mode, &factor, &calibMs);
            if (rc != DXB_OK)
                return rc;
        } else {
            factor = beamAnalyticCalibration(*beam);
        }
    }
    CUDA_TRY(c, cudaSetDevice(d0.device));
    DevBuf<const unsigned long long*> dPeers;
    if (!multicast_tally && n_peers > 0) {
        std::vector<const unsigned long long*> peers(n_peers);
        for (int i = 0; i < n_peers; ++i)
            peers[i] = static_cast<const unsigned long long*>(peer_tallies[i]);
        CUDA_TRY(c, dPeers.upload(peers, d0.device, d0.stream));
    }
    const double vol = w.spacing[0] * w.spacing[1] * w.spacing[2];
    const unsigned long long* src = multicast_tally ? static_cast<const unsigned long long*>(multicast_tally) : w.tally.p;
    launchFusedReduceToDose(src, multicast_tally != nullptr, dPeers.p, multicast_tally ? 0 : n_peers,
        n_peers > 0 ? static_cast<int>(c->rank % static_cast<uint64_t>(n_peers)) : 0, w.voxels.p, d0.dose.p, d0.variance.p,
        d0.events.p, voxel_begin, voxel_end, 1.0 / c->scaleE, 1.0 / c->scaleE2, factor, vol, d0.stream);
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaEventRecord(d0.evEnd, d0.stream));
    CUDA_TRY(c, cudaStreamSynchronize(d0.stream));
    c->tallyValid = false; // a beam is finished once
    c->stats.calibration_factor = factor;
    c->stats.calibration_ms = calibMs;
    if (factor_out)
        *factor_out = factor;
    return DXB_OK;
}

int dxb_dose_buffers(dxb_ctx* c, void** dose, void** variance, void** n_events, uint64_t* n_voxels)
{
    if (!c || c->devs.empty() || !c->devs[0]->world.hasGrid)
        return fail(c, DXB_ESTATE, "dose_buffers: no grid");
    const int grc = mgGatherDose(c);
    if (grc != DXB_OK)
        return grc;
    DeviceState& d0 = *c->devs[0];
    if (dose)
        *dose = d0.dose.p;
    if (variance)
        *variance = d0.variance.p;
    if (n_events)
        *n_events = d0.events.p;
    if (n_voxels)
        *n_voxels = d0.world.nvox;
    return DXB_OK;
}

int dxb_run(dxb_ctx* c, const dxb_beam_desc* beam, int physics_mode, int use_beam_calibration, dxb_progress* progress)
{
    const auto t0 = std::chrono::steady_clock::now();
    int rc = dxb_run_transport(c, beam, physics_mode, progress);
    if (rc != DXB_OK)
        return rc;
    if (c->world > 1)
        return fail(c, DXB_ESTATE, "dxb_run on a context that is one shard of a multi-process job: use dxb_run_transport, a barrier over the ranks, dxb_finish_beam");
    rc = dxb_finish_beam(c, beam, physics_mode, use_beam_calibration, nullptr);
    if (rc != DXB_OK)
        return rc;
    c->stats.total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    return DXB_OK;
}

int dxb_tally_buffer(dxb_ctx* c, void** device_ptr, uint64_t* n_words)
{
    if (!c || c->devs.empty() || !c->devs[0]->world.hasGrid)
        return fail(c, DXB_ESTATE, "tally_buffer: no grid");
    if (c->exchanging)
        return fail(c, DXB_ESTATE, "tally_buffer: the library exchanges the tallies of this context itself");
    if (device_ptr)
        *device_ptr = c->devs[0]->world.tally.p;
    if (n_words)
        *n_words = c->devs[0]->world.nvox * 4;
    return DXB_OK;
}

int dxb_get_dose(dxb_ctx* c, double* dose, double* variance, uint64_t* n_events)
{
    if (!c || c->devs.empty() || !c->devs[0]->world.hasGrid)
        return fail(c, DXB_ESTATE, "get_dose: no grid");
    DeviceState& d0 = *c->devs[0];
    const size_t n = d0.world.nvox;
    if (c->devs.size() > 1) // every device copies its slab of the dose score over its own PCIe link
        return mgGetDose(c, 0, n, dose, variance, n_events);
    int rc = mgFlush(c);
    if (rc != DXB_OK)
        return rc;
    CUDA_TRY(c, cudaSetDevice(d0.device));
    if (dose)
        CUDA_TRY(c, cudaMemcpyAsync(dose, d0.dose.p, n * sizeof(double), cudaMemcpyDeviceToHost, d0.stream));
    if (variance)
        CUDA_TRY(c, cudaMemcpyAsync(variance, d0.variance.p, n * sizeof(double), cudaMemcpyDeviceToHost, d0.stream));
    if (n_events)
        CUDA_TRY(c, cudaMemcpyAsync(n_events, d0.events.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, d0.stream));
    CUDA_TRY(c, cudaStreamSynchronize(d0.stream));
    return DXB_OK;
}

int dxb_get_dose_range(dxb_ctx* c, uint64_t voxel_begin, uint64_t voxel_end, double* dose, double* variance, uint64_t* n_events)
{
    if (!c || c->devs.empty() || !c->devs[0]->world.hasGrid)
        return fail(c, DXB_ESTATE, "get_dose_range: no grid");
    DeviceState& d0 = *c->devs[0];
    const size_t n = d0.world.nvox;
    if (voxel_begin > voxel_end || voxel_end > n)
        return fail(c, DXB_EINVAL, "get_dose_range: voxel range outside the grid");
    const size_t b = voxel_begin, m = voxel_end - voxel_begin;
    if (c->devs.size() > 1)
        return mgGetDose(c, voxel_begin, voxel_end, dose, variance, n_events);
    int rc = mgFlush(c);
    if (rc != DXB_OK)
        return rc;
    CUDA_TRY(c, cudaSetDevice(d0.device));
    if (m > 0) {
        if (dose)
            CUDA_TRY(c, cudaMemcpyAsync(dose + b, d0.dose.p + b, m * sizeof(double), cudaMemcpyDeviceToHost, d0.stream));
        if (variance)
            CUDA_TRY(c, cudaMemcpyAsync(variance + b, d0.variance.p + b, m * sizeof(double), cudaMemcpyDeviceToHost, d0.stream));
        if (n_events)
            CUDA_TRY(c, cudaMemcpyAsync(n_events + b, d0.events.p + b, m * sizeof(uint64_t), cudaMemcpyDeviceToHost, d0.stream));
    }
    CUDA_TRY(c, cudaStreamSynchronize(d0.stream));
    return DXB_OK;
}

int dxb_get_energy_scored(dxb_ctx* c, double* energy, double* energy_sq, uint64_t* n_events)
{
    if (!c || c->devs.empty() || !c->devs[0]->world.hasGrid)
        return fail(c, DXB_ESTATE, "get_energy_scored: no grid");
    DeviceState& d0 = *c->devs[0];
    World& w = d0.world;
    const size_t n = w.nvox;
    if (c->exchanging && c->exchanged)
        return fail(c, DXB_ESTATE, "get_energy_scored: the tallies of the last beam have been handed to the exchange (read them before dxb_finish_beam)");
    CUDA_TRY(c, cudaSetDevice(d0.device));
    DevBuf<unsigned long long> summed; // several devices: sum of their buffers (integers: identical to one device)
    const unsigned long long* tally = w.tallyCur();
    if (c->devs.size() > 1) {
        const int rc = mgSumTallies(c, summed);
        if (rc != DXB_OK)
            return rc;
        tally = summed.p;
    }
    DevBuf<double> e, e2;
    DevBuf<unsigned long long> cnt;
    if (energy)
        CUDA_TRY(c, e.alloc(n, d0.device));
    if (energy_sq)
        CUDA_TRY(c, e2.alloc(n, d0.device));
    if (n_events)
        CUDA_TRY(c, cnt.alloc(n, d0.device));
    launchTallyToEnergy(tally, e.p, e2.p, cnt.p, n, 1.0 / c->scaleE, 1.0 / c->scaleE2, d0.stream);
    CUDA_TRY(c, cudaGetLastError());
    if (energy)
        CUDA_TRY(c, cudaMemcpyAsync(energy, e.p, n * sizeof(double), cudaMemcpyDeviceToHost, d0.stream));
    if (energy_sq)
        CUDA_TRY(c, cudaMemcpyAsync(energy_sq, e2.p, n * sizeof(double), cudaMemcpyDeviceToHost, d0.stream));
    if (n_events)
        CUDA_TRY(c, cudaMemcpyAsync(n_events, cnt.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, d0.stream));
    CUDA_TRY(c, cudaStreamSynchronize(d0.stream));
    return DXB_OK;
}

int dxb_get_dose_postprocessed(dxb_ctx* c, int delete_air_dose, double* dose, double* variance, double* n_events, char units_out[4])
{
    if (!c || c->devs.empty() || !c->devs[0]->world.hasGrid)
        return fail(c, DXB_ESTATE, "get_dose_postprocessed: no grid");
    DeviceState& d0 = *c->devs[0];
    World& w = d0.world;
    const size_t n = w.nvox;
    const int grc = mgGatherDose(c); // waits for pending exchanges; several devices: all slabs of the dose score -> device 0
    if (grc != DXB_OK)
        return grc;
    CUDA_TRY(c, cudaSetDevice(d0.device));
    // max(dose) < 1 mGy -> report uGy (dose x 1e3, variance x 1e6), R:src/libopendxmc/simulationpipeline.cpp:187-195,227-229
    DevBuf<unsigned long long> dMax;
    CUDA_TRY(c, dMax.alloc(1, d0.device));
    CUDA_TRY(c, cudaMemsetAsync(dMax.p, 0, sizeof(unsigned long long), d0.stream));
    launchMax(d0.dose.p, w.voxels.p, n, delete_air_dose, dMax.p, d0.stream);
    unsigned long long bits = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&bits, dMax.p, sizeof(bits), cudaMemcpyDeviceToHost, d0.stream));
    CUDA_TRY(c, cudaStreamSynchronize(d0.stream));
    double maxDose;
    std::memcpy(&maxDose, &bits, sizeof(double));
    const bool micro = maxDose < 1.0;
    if (units_out)
        std::memcpy(units_out, micro ? "uGy" : "mGy", 4);
    DevBuf<double> tmp;
    CUDA_TRY(c, tmp.alloc(n, d0.device));
    if (dose) {
        launchPostprocess(d0.dose.p, w.voxels.p, tmp.p, n, delete_air_dose, micro ? 1e3 : 1.0, d0.stream);
        CUDA_TRY(c, cudaMemcpyAsync(dose, tmp.p, n * sizeof(double), cudaMemcpyDeviceToHost, d0.stream));
        CUDA_TRY(c, cudaStreamSynchronize(d0.stream));
    }
    if (n_events) {
        launchU64ToDouble(d0.events.p, w.voxels.p, tmp.p, n, delete_air_dose, d0.stream);
        CUDA_TRY(c, cudaMemcpyAsync(n_events, tmp.p, n * sizeof(double), cudaMemcpyDeviceToHost, d0.stream));
        CUDA_TRY(c, cudaStreamSynchronize(d0.stream));
    }
    if (variance) {
        launchPostprocess(d0.variance.p, w.voxels.p, tmp.p, n, delete_air_dose, micro ? 1e6 : 1.0, d0.stream);
        CUDA_TRY(c, cudaMemcpyAsync(variance, tmp.p, n * sizeof(double), cudaMemcpyDeviceToHost, d0.stream));
        CUDA_TRY(c, cudaStreamSynchronize(d0.stream));
    }
    return DXB_OK;
}

int dxb_organ_dose(dxb_ctx* c, const uint8_t* organ, uint32_t n_organs, double* dose_out, double* mass_out,
    uint64_t* n_voxels_out, double* variance_out)
{
    if (!c || !organ || n_organs == 0 || n_organs > 256)
        return fail(c, DXB_EINVAL, "organ_dose: bad arguments");
    if (c->devs.empty() || !c->devs[0]->world.hasGrid)
        return fail(c, DXB_ESTATE, "organ_dose: no grid");
    DeviceState& d0 = *c->devs[0];
    World& w = d0.world;
    const size_t n = w.nvox;
    const int grc = mgGatherDose(c);
    if (grc != DXB_OK)
        return grc;
    CUDA_TRY(c, cudaSetDevice(d0.device));
    DevBuf<unsigned char> dOrg;
    DevBuf<double> acc; // energy[256], mass[256], var[256]
    DevBuf<unsigned long long> cnt;
    CUDA_TRY(c, dOrg.alloc(n, d0.device));
    CUDA_TRY(c, acc.alloc(768, d0.device));
    CUDA_TRY(c, cnt.alloc(256, d0.device));
    CUDA_TRY(c, cudaMemcpyAsync(dOrg.p, organ, n, cudaMemcpyHostToDevice, d0.stream));
    CUDA_TRY(c, cudaMemsetAsync(acc.p, 0, 768 * sizeof(double), d0.stream));
    CUDA_TRY(c, cudaMemsetAsync(cnt.p, 0, 256 * sizeof(unsigned long long), d0.stream));
    const double vol = w.spacing[0] * w.spacing[1] * w.spacing[2];
    launchOrganDose(d0.dose.p, d0.variance.p, w.voxels.p, dOrg.p, n, vol, acc.p, acc.p + 256, cnt.p, acc.p + 512, d0.stream);
    CUDA_TRY(c, cudaGetLastError());
    double h[768];
    unsigned long long hc[256];
    CUDA_TRY(c, cudaMemcpyAsync(h, acc.p, sizeof(h), cudaMemcpyDeviceToHost, d0.stream));
    CUDA_TRY(c, cudaMemcpyAsync(hc, cnt.p, sizeof(hc), cudaMemcpyDeviceToHost, d0.stream));
    CUDA_TRY(c, cudaStreamSynchronize(d0.stream));
    for (uint32_t o = 0; o < n_organs; ++o) {
        const double mass = h[256 + o];
        if (dose_out)
            dose_out[o] = mass > 0 ? h[o] / mass : 0.0;
        if (mass_out)
            mass_out[o] = mass;
        if (n_voxels_out)
            n_voxels_out[o] = hc[o];
        if (variance_out)
            variance_out[o] = mass > 0 ? h[512 + o] / (mass * mass) : 0.0;
    }
    return DXB_OK;
}

int dxb_get_run_stats(const dxb_ctx* c, dxb_run_stats* out)
{
    if (!c || !out)
        return DXB_EINVAL;
    *out = c->stats;
    return DXB_OK;
}

int dxb_device_attenuation(dxb_ctx* c, uint32_t material_index, int physics_mode, const double* energy_kev, uint32_t n, float* out4)
{
    (void)physics_mode;
    if (!c || !energy_kev || !out4 || n == 0)
        return DXB_EINVAL;
    if (c->devs.empty() || !c->devs[0]->world.hasTables || material_index >= static_cast<uint32_t>(c->devs[0]->world.n_mat))
        return fail(c, DXB_ESTATE, "device_attenuation: no tables / bad material");
    DeviceState& d0 = *c->devs[0];
    CUDA_TRY(c, cudaSetDevice(d0.device));
    std::vector<float> e(n);
    for (uint32_t i = 0; i < n; ++i)
        e[i] = static_cast<float>(energy_kev[i]);
    DevBuf<float> dE, dOut;
    CUDA_TRY(c, dE.upload(e, d0.device, d0.stream));
    CUDA_TRY(c, dOut.alloc(static_cast<size_t>(n) * 4, d0.device));
    launchAttenuationProbe(d0.world.tablesDev(), static_cast<int>(material_index), dE.p, static_cast<int>(n), dOut.p, d0.stream);
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaMemcpyAsync(out4, dOut.p, static_cast<size_t>(n) * 4 * sizeof(float), cudaMemcpyDeviceToHost, d0.stream));
    CUDA_TRY(c, cudaStreamSynchronize(d0.stream));
    return DXB_OK;
}

int dxb_get_dense_box(dxb_ctx* c, int* built, int* useful, int box[6], float faces[6], float ratio[16])
{
    if (!c || c->devs.empty() || !c->devs[0]->world.hasGrid)
        return fail(c, DXB_ESTATE, "get_dense_box: no grid");
    const World& w = c->devs[0]->world;
    if (built)
        *built = w.dbBuilt ? 1 : 0;
    if (useful)
        *useful = w.dbUseful ? 1 : 0;
    for (int a = 0; a < 6; ++a) {
        if (box)
            box[a] = w.dbBuilt ? w.dbBox[a] : 0;
        if (faces)
            faces[a] = w.dbBuilt ? w.dbFaces[a] : 0.0f;
    }
    if (ratio)
        for (int b = 0; b < 16; ++b)
            ratio[b] = w.dbBuilt ? w.dbRatio[b] : 1.0f;
    return DXB_OK;
}

int dxb_get_local_majorant(dxb_ctx* c, int* n_slabs, int* shift, int* useful, float* ratio)
{
    if (!c || c->devs.empty() || !c->devs[0]->world.hasGrid)
        return fail(c, DXB_ESTATE, "get_local_majorant: no grid");
    const World& w = c->devs[0]->world;
    if (n_slabs)
        *n_slabs = w.lmSlabs;
    if (shift)
        *shift = w.lmShift;
    if (useful)
        *useful = w.lmUseful ? 1 : 0;
    if (ratio && w.lmSlabs > 0)
        std::memcpy(ratio, w.lmHost.data(), static_cast<size_t>(w.lmSlabs) * kLmBands * sizeof(float));
    return DXB_OK;
}

int dxb_device_majorant(dxb_ctx* c, const double* energy_kev, uint32_t n, float* out)
{
    if (!c || !energy_kev || !out || n == 0)
        return DXB_EINVAL;
    if (c->devs.empty() || !c->devs[0]->world.hasGrid)
        return fail(c, DXB_ESTATE, "device_majorant: no grid");
    DeviceState& d0 = *c->devs[0];
    CUDA_TRY(c, cudaSetDevice(d0.device));
    std::vector<float> e(n);
    for (uint32_t i = 0; i < n; ++i)
        e[i] = static_cast<float>(energy_kev[i]);
    DevBuf<float> dE, dOut;
    CUDA_TRY(c, dE.upload(e, d0.device, d0.stream));
    CUDA_TRY(c, dOut.alloc(n, d0.device));
    launchMajorantProbe(d0.world.majorant.p, dE.p, static_cast<int>(n), dOut.p, d0.stream);
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaMemcpyAsync(out, dOut.p, n * sizeof(float), cudaMemcpyDeviceToHost, d0.stream));
    CUDA_TRY(c, cudaStreamSynchronize(d0.stream));
    return DXB_OK;
}

// CT segmentation, SURVEY §8f-1 — R:src/libopendxmc/ctsegmentationpipeline.cpp:61-169
int dxb_segment_ct(dxb_ctx* c, const double* hu, uint64_t n, const dxb_tube_desc* tube, uint8_t* material_out,
    double* density_out, dxb_material** materials_out)
{
    if (!c || !hu || n == 0 || !tube || !material_out || !density_out)
        return fail(c, DXB_EINVAL, "segment_ct: null argument");
    static const char* names[5] = { "Air, Dry (near sea level)", "Adipose Tissue (ICRP)", "Tissue, Soft (ICRP)", "Muscle, Skeletal",
        "Bone, Cortical (ICRP)" };
    std::vector<std::shared_ptr<Material>> mats;
    std::vector<double> dens;
    for (const char* nm : names) {
        auto m = Material::byNistName(nm);
        if (!m)
            return fail(c, DXB_EMATERIAL, "segment_ct: material");
        mats.push_back(m);
        dens.push_back(nistFind(nm)->density);
    }
    dens.back() = 1.09; // "Density for bone is to high", R:...ctsegmentationpipeline.cpp:126-127
    const auto en = tubeEnergies(*tube);
    const auto sw = tubeSpectrum(*tube, en, true);
    auto air = mats[0];
    auto water = Material::byNistName("Water, Liquid");
    const double airD = nistFind("Air, Dry (near sea level)")->density, waterD = 1.0;
    std::vector<double> HU(5, 0.0), att(5, 0.0);
    double attW = 0, attA = 0;
    for (size_t k = 0; k < en.size(); ++k) {
        const double e = std::clamp(en[k], kEMin, kEMax);
        const double uw = water->total(e), ua = air->total(e);
        attW += sw[k] * uw;
        attA += sw[k] * ua;
        for (int i = 0; i < 5; ++i) {
            const double um = mats[i]->total(e);
            att[i] += sw[k] * um;
            HU[i] += 1000.0 * sw[k] * (um * dens[i] - uw * waterD) / (uw * waterD - ua * airD);
        }
    }
    std::vector<double> sep;
    for (int i = 0; i < 4; ++i)
        sep.push_back(0.5 * (HU[i] + HU[i + 1]));
    DeviceState& d0 = *c->devs[0];
    CUDA_TRY(c, cudaSetDevice(d0.device));
    DevBuf<double> dHu, dSep, dAtt, dDens;
    DevBuf<unsigned char> dMat;
    CUDA_TRY(c, dHu.alloc(n, d0.device));
    CUDA_TRY(c, dDens.alloc(n, d0.device));
    CUDA_TRY(c, dMat.alloc(n, d0.device));
    CUDA_TRY(c, dSep.upload(sep, d0.device, d0.stream));
    CUDA_TRY(c, dAtt.upload(att, d0.device, d0.stream));
    CUDA_TRY(c, cudaMemcpyAsync(dHu.p, hu, n * sizeof(double), cudaMemcpyHostToDevice, d0.stream));
    launchSegment(dHu.p, n, dSep.p, 4, dAtt.p, attW * waterD, attA * airD, dMat.p, dDens.p, d0.stream);
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaMemcpyAsync(material_out, dMat.p, n, cudaMemcpyDeviceToHost, d0.stream));
    CUDA_TRY(c, cudaMemcpyAsync(density_out, dDens.p, n * sizeof(double), cudaMemcpyDeviceToHost, d0.stream));
    CUDA_TRY(c, cudaStreamSynchronize(d0.stream));
    if (materials_out)
        for (int i = 0; i < 5; ++i)
            materials_out[i] = new dxb_material { mats[i] };
    return DXB_OK;
}

// -------------------------------------------------------------------------- progress
dxb_progress* dxb_progress_create(void) { return new dxb_progress(); }
void dxb_progress_destroy(dxb_progress* p) { delete p; }
void dxb_progress_read(const dxb_progress* p, uint64_t* done, uint64_t* total)
{
    if (!p)
        return;
    if (done)
        *done = p->done.load(std::memory_order_relaxed);
    if (total)
        *total = p->total.load(std::memory_order_relaxed);
}
void dxb_progress_stop(dxb_progress* p)
{
    if (p)
        p->stop.store(1, std::memory_order_relaxed);
}
int dxb_progress_continue(const dxb_progress* p) { return p ? !p->stop.load(std::memory_order_relaxed) : 1; }
void dxb_progress_reset(dxb_progress* p)
{
    if (!p)
        return;
    p->stop.store(0);
    p->done.store(0);
    p->total.store(0);
}
int dxb_progress_message(const dxb_progress* p, char* buf, int cap)
{
    if (!p || !buf || cap <= 0)
        return DXB_EINVAL;
    const uint64_t done = p->done.load(), total = p->total.load();
    const int64_t now = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
    const double elapsed = (now - p->start_ns.load()) * 1e-9;
    if (done == 0 || total == 0) {
        std::snprintf(buf, cap, "Starting simulation");
    } else {
        const double remaining = elapsed * static_cast<double>(total - done) / static_cast<double>(done);
        const int s = static_cast<int>(remaining);
        std::snprintf(buf, cap, "Remaining time %02d:%02d:%02d", s / 3600, (s / 60) % 60, s % 60);
    }
    return DXB_OK;
}

} // extern "C"
