// device_types.cuh — PODs shared between the host context and the sm_100a kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace dxb {

// table geometry, mirrors physics.hpp (static_asserted in context.cu)
constexpr int kDevNE = 465;
constexpr int kDevEPerOctave = 64;
constexpr int kDevNX = 353;
constexpr int kDevXPerOctave = 32;
constexpr float kDevXMinInv = 128.0f;
constexpr int kMaxShells = 5;
constexpr int kSourceBufWords = 12; // words per entry of the per-warp source buffer (transport.cu)

// HBM layout of the voxel grid (x fastest, like the reference: i + j*nx + k*nx*ny,
// R:src/libopendxmc/otherphantomimportpipeline.cpp:44).
//   voxels : u32 {top 24 bits of the f32 density (round to nearest, 2^-17 relative), material index}   4 B / voxel
//            (the 3.84 cm primary-beam slab of config C2 is then 40 MB and stays resident in the 126 MB L2)
//   tally  : 4 x u64 {sum E, sum E^2, events, pad}           32 B / voxel = one DRAM sector,
//            64-bit fixed point so that sums are order independent
//   dose   : 3 x f64 {dose, variance, events} accumulated over beams (DoseScore)
struct GridDev {
    int nx, ny, nz;
    float x0, y0, z0; // min corner [cm]
    float x1, y1, z1; // max corner
    float inv_dx, inv_dy, inv_dz;
    float offx, offy, offz; // voxel coordinate = p * inv_d + off
    const unsigned int* __restrict__ voxels;
    unsigned long long* __restrict__ tally;
};

// packed voxel helpers
__host__ __device__ inline unsigned int quantizeDensityBits(unsigned int f32bits) { return (f32bits + 0x80u) & 0xFFFFFF00u; }
#ifdef __CUDACC__
__device__ __forceinline__ float voxelDensity(unsigned int v) { return __uint_as_float(v & 0xFFFFFF00u); }
__device__ __forceinline__ int voxelMaterial(unsigned int v) { return static_cast<int>(v & 0xFFu); }
#endif

struct ShellDev {
    float binding, nel_fraction, photo_fraction, fluor_yield, fluor_energy, j0, pad0, pad1;
};

struct TablesDev {
    int n_mat;
    const float4* __restrict__ att;   // [n_mat*NE] {photo, incoh, coh, total} cm2/g
    const float* __restrict__ tot;    // [n_mat*NE] total, staged in shared memory by the kernel
    const float* __restrict__ etr;    // [n_mat*NE] mass energy-transfer (kerma estimator)
    const float* __restrict__ majorant; // [NE] mu_max [1/cm]
    const float* __restrict__ ffcdf;  // [n_mat*NX]
    const float* __restrict__ sf;     // [n_mat*NX]
    const ShellDev* __restrict__ shells; // [n_mat*kMaxShells]
    const int* __restrict__ n_shells; // [n_mat]
    const float* __restrict__ rest_j0; // [n_mat] J(0) of the electrons outside the shell table (0: at rest)
};

struct SpectrumDev {
    int n;             // 1: mono-energetic
    float e0, step;    // energy[i] = e0 + i*step
    const float* __restrict__ prob;           // alias acceptance
    const unsigned short* __restrict__ alias;
};

struct BowtieDev {
    int n; // 0: none
    const float* __restrict__ angle;
    const float* __restrict__ weight;
};

struct ExposureDev { // 64 B
    float pos[3];
    float c0[3];
    float c1[3];
    float dir[3];
    float hx, hy;
    float weight;
    int tube;
};

struct RunParams {
    GridDev grid;
    TablesDev tab;
    SpectrumDev spec[2];
    BowtieDev bow[2];
    const ExposureDev* __restrict__ exposures;
    unsigned long long n_exposures;
    unsigned long long ppe;         // particles per exposure
    unsigned long long n_total;     // histories of the whole beam (all ranks)
    unsigned long long local_begin; // this launch covers local indices [local_begin, local_end)
    unsigned long long local_end;
    unsigned int world, rank;       // history sharding: 65536-history blocks dealt round-robin
    unsigned int seed_lo, seed_hi;  // Philox key
    unsigned int round_key[10][2];  // Philox round keys (key + r * Weyl constants), read from the constant bank
    float tally_scale_e, tally_scale_e2;
    int score_material;             // calibration: kerma collision estimator in this material (-1: off)
    int refill_threshold;           // dead lanes per warp that trigger a refill phase
    int interact_threshold;         // waiting lanes per warp that trigger an interaction phase
    int rayleigh_threshold;         // lanes waiting for a Rayleigh try that trigger a Rayleigh phase
    int step_pairs;                 // mux kernel: step pairs per step phase (>= 1)
    int step_quad;                  // pool kernel: two step pairs per phase with all four gathers in flight
    int diag;                       // pool kernel: stats[8..15] = executions / claimed lanes per phase
    int service_warps;              // pool kernel: warps per block that prefer interaction / refill phases
    int interact_bias;              // mux kernel: an interaction phase runs when waiting lanes + bias >= stepping lanes;
                                    // pool kernel: stepper warps keep stepping while at least this many lanes can
    // slab-local majorants (pool kernel, LM builds): the grid is cut into slabs of 2^lm_shift voxel layers along z; inside
    // slab s and energy band b (32 energy nodes) the tracking majorant is mu_max(E) * lm_ratio[s * kLmBands + b]
    const float* __restrict__ lm_ratio; // [lm_slabs * kLmBands], in (0, 1]
    int lm_slabs, lm_shift;
    float lm_thickness;              // slab thickness [cm] = 2^lm_shift * dz
    // brick pre-filter (pool kernel, quad step): per brick of 2^brick_shift voxels per edge and per energy octave (energy node
    // index >> 6) an upper bound (q + 1) / 256 of mu(voxel) / mu_max over the brick; a tentative collision whose uniform number
    // is not below the bound is virtual whatever the voxel holds, so its gather is skipped.  nullptr: no filter.
    const unsigned char* __restrict__ brick; // [bricks * 8]
    int brick_shift, brick_nx, brick_ny;
    // dense box (pool kernel, DB builds): every voxel that is not thin (air) lies inside the box; inside it the photons are
    // tracked with mu_max(E), in the rest of the grid with mu_max(E) * db_ratio[band] (flights, transport_pool.cu)
    float db_lo[3], db_hi[3];  // faces of the box [cm]
    int db_i0[3], db_n[3];     // first voxel index and extent per axis
    float db_ratio[16];        // kLmBands energy bands (energy node index >> 5)
    float db_diag;             // length of the grid's diagonal [cm], rounded up: no path through the grid is longer
    unsigned int hbase_lo, hbase_hi; // mux kernel: global id of the first history of this launch (ids of one launch span < 2^32)
    unsigned long long* __restrict__ work_counter; // global cursor into [local_begin, local_end)
    unsigned long long* __restrict__ stats;        // [5]: steps, interactions, deposits, emitted (2^-16 keV), histories
};

constexpr unsigned int kShardBlock = 65536; // histories per sharding block
constexpr int kLmBands = 16;                // energy bands of the slab-local majorant table: band = energy node index >> 5
constexpr int kLmMaxSlabs = 256;

// dynamic shared memory of the transport kernel:
//   [per warp: 4 x u64 history-pool words][per warp: source buffer, kSourceBufWords x 32 words]
//   [per thread: kLaneCounters words][optional: total-attenuation table, n_mat x kDevNE floats]
constexpr int kLaneCounters = 5; // interactions, deposits, histories, emitted lo, emitted hi
__host__ __device__ inline size_t transportSmemBytes(int threads, int table_floats)
{
    const size_t warps = static_cast<size_t>(threads) / 32;
    return warps * 32 + warps * kSourceBufWords * 32 * 4 + static_cast<size_t>(threads) * kLaneCounters * 4 + static_cast<size_t>(table_floats) * 4;
}

// the lane-multiplexed kernel (transport_mux.cu): every lane owns `slots` photons in shared memory
constexpr int kSlotWords = 10;  // px py pz dx dy dz E w hlo meta
constexpr int kMuxBufWords = 9; // px py pz dx dy dz E w lane-offset
__host__ __device__ inline size_t muxSmemBytes(int threads, int slots, int table_floats)
{
    const size_t warps = static_cast<size_t>(threads) / 32;
    return (static_cast<size_t>(table_floats) + kDevNE) * 4 + warps * (static_cast<size_t>(kSlotWords) * slots + kMuxBufWords) * 32 * 4;
}

// the block-pooled kernel (transport_pool.cu): 32 classes x `slots` photons per block + 2 status words per class
__host__ __device__ inline size_t poolSmemBytes(int slots, int table_floats, int lm_slabs = 0)
{
    return (static_cast<size_t>(table_floats) + kDevNE + 64 + static_cast<size_t>(kSlotWords) * slots * 32 + static_cast<size_t>(lm_slabs) * kLmBands) * 4;
}

// launch wrappers (transport.cu, transport_mux.cu, transport_pool.cu)
struct LaunchConfig {
    int blocks, threads;
    size_t smem;
    bool table_in_smem;
    int slots; // 0: one photon per lane in registers (transport.cu); >= 2: slots per lane (mux) / per class (pool)
    bool pool; // block-pooled kernel instead of the lane-multiplexed one
    int min_blocks; // pool kernel: 5 / 6 = the 48 / 40-register builds (5 / 6 blocks of <= 256 threads per SM), else 64 registers
    bool local_majorant; // pool kernel: the slab-local majorant build
    bool brick_filter;   // pool kernel: the brick pre-filter build of the quad step
    bool dense_box = false; // pool kernel: the dense-box build of the quad step
};

} // namespace dxb
