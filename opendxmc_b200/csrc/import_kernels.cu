// import_kernels.cu — the O(N) device passes of the input producers next to the hot path:
// ICRP phantom import (organ -> medium -> density remap, R:src/libopendxmc/icrpphantomimportpipeline.cpp:209-351).
// Both kernels stream the u8 organ array once (HBM-bound: 1 B in; 1 + 1 + 8 B out per voxel).
#include "context_types.hpp"
#include "icrp.hpp"

namespace dxb {
namespace {

// which of the 256 organ values occur (pruneOrganArray's `id_exists` scans, :218-221, for all ids at once)
__global__ void organPresenceKernel(const unsigned char* __restrict__ organ, size_t n, unsigned int* __restrict__ present /*[256]*/)
{
    __shared__ unsigned int s_seen[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        s_seen[i] = 0u;
    __syncthreads();
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x * 16;
    for (size_t base = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 16; base < n; base += stride) {
        if (base + 16 <= n) {
            const uint4 v = *reinterpret_cast<const uint4*>(organ + base);
            const unsigned int w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    s_seen[(w[k] >> (8 * b)) & 0xffu] = 1u; // benign race: every writer stores 1
        } else {
            for (size_t i = base; i < n; ++i)
                s_seen[organ[i]] = 1u;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        if (s_seen[i])
            present[i] = 1u;
}

// organ value -> (new organ value, medium, density): three 256-entry tables in shared memory, 16 voxels per thread and step
__global__ void icrpRemapKernel(const unsigned char* __restrict__ in, size_t n, const unsigned char* __restrict__ organLut,
    const unsigned char* __restrict__ materialLut, const double* __restrict__ densityLut, unsigned char* __restrict__ organOut,
    unsigned char* __restrict__ materialOut, double* __restrict__ densityOut)
{
    __shared__ unsigned char s_org[256], s_mat[256];
    __shared__ double s_den[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        s_org[i] = organLut[i];
        s_mat[i] = materialLut[i];
        s_den[i] = densityLut[i];
    }
    __syncthreads();
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x * 16;
    for (size_t base = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 16; base < n; base += stride) {
        if (base + 16 <= n) {
            const uint4 v = *reinterpret_cast<const uint4*>(in + base);
            const unsigned int w[4] = { v.x, v.y, v.z, v.w };
            unsigned int o[4], m[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                o[k] = 0u;
                m[k] = 0u;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const unsigned int val = (w[k] >> (8 * b)) & 0xffu;
                    o[k] |= static_cast<unsigned int>(s_org[val]) << (8 * b);
                    m[k] |= static_cast<unsigned int>(s_mat[val]) << (8 * b);
                    densityOut[base + 4 * k + b] = s_den[val];
                }
            }
            *reinterpret_cast<uint4*>(organOut + base) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4*>(materialOut + base) = make_uint4(m[0], m[1], m[2], m[3]);
        } else {
            for (size_t i = base; i < n; ++i) {
                const unsigned int val = in[i];
                organOut[i] = s_org[val];
                materialOut[i] = s_mat[val];
                densityOut[i] = s_den[val];
            }
        }
    }
}

} // namespace
} // namespace dxb

using namespace dxb;

extern "C" {

int dxb_icrp_plan(dxb_icrp** out, const char* organs_dat, const char* media_dat, int remove_arms, const uint8_t present[256])
{
    if (!out || !organs_dat || !media_dat || !present)
        return DXB_EINVAL;
    *out = nullptr;
    auto p = std::make_unique<dxb_icrp>();
    if (!icrpPlan(organs_dat, media_dat, remove_arms != 0, present, p->plan))
        return DXB_EINVAL; // a table without a single valid line: the reference returns without a result (:223-224, :250-251)
    *out = p.release();
    return DXB_OK;
}

void dxb_icrp_destroy(dxb_icrp* p) { delete p; }

int dxb_icrp_luts(const dxb_icrp* p, uint8_t organ_lut[256], uint8_t material_lut[256], double density_lut[256])
{
    if (!p)
        return DXB_EINVAL;
    for (int v = 0; v < 256; ++v) {
        if (organ_lut)
            organ_lut[v] = p->plan.organLut[v];
        if (material_lut)
            material_lut[v] = p->plan.materialLut[v];
        if (density_lut)
            density_lut[v] = p->plan.densityLut[v];
    }
    return DXB_OK;
}

uint32_t dxb_icrp_n_organs(const dxb_icrp* p) { return p ? static_cast<uint32_t>(p->plan.organNames.size()) : 0; }
const char* dxb_icrp_organ_name(const dxb_icrp* p, uint32_t i) { return (p && i < p->plan.organNames.size()) ? p->plan.organNames[i].c_str() : ""; }
double dxb_icrp_organ_density(const dxb_icrp* p, uint32_t i) { return (p && i < p->plan.organDensity.size()) ? p->plan.organDensity[i] : 0.0; }
uint32_t dxb_icrp_organ_medium(const dxb_icrp* p, uint32_t i) { return (p && i < p->plan.organMedium.size()) ? p->plan.organMedium[i] : 0; }
uint32_t dxb_icrp_n_media(const dxb_icrp* p) { return p ? static_cast<uint32_t>(p->plan.mediaNames.size()) : 0; }
const char* dxb_icrp_medium_name(const dxb_icrp* p, uint32_t i) { return (p && i < p->plan.mediaNames.size()) ? p->plan.mediaNames[i].c_str() : ""; }
int dxb_icrp_medium_composition(const dxb_icrp* p, uint32_t i, uint32_t* Z, double* weight, int cap)
{
    if (!p || i >= p->plan.mediaComposition.size())
        return 0;
    const auto& c = p->plan.mediaComposition[i];
    for (int k = 0; k < cap && k < static_cast<int>(c.size()); ++k) {
        if (Z)
            Z[k] = c[k].first;
        if (weight)
            weight[k] = c[k].second;
    }
    return static_cast<int>(c.size());
}

int dxb_icrp_import(dxb_ctx* c, const uint8_t* organ_in, uint64_t n, const char* organs_dat, const char* media_dat, int remove_arms,
    uint8_t* organ_out, uint8_t* material_out, double* density_out, dxb_icrp** plan_out)
{
    if (!c || c->devs.empty() || !organ_in || n == 0 || !organs_dat || !media_dat || !organ_out || !material_out || !density_out || !plan_out)
        return fail(c, DXB_EINVAL, "icrp_import: null argument");
    *plan_out = nullptr;
    DeviceState& d0 = *c->devs[0];
    CUDA_TRY(c, cudaSetDevice(d0.device));
    const int blocks = c->smCount * 8;
    DevBuf<unsigned char> dIn, dOrg, dMat, dLut;
    DevBuf<double> dDen, dDenLut;
    DevBuf<unsigned int> dPresent;
    CUDA_TRY(c, dIn.alloc(n + 16, d0.device));
    CUDA_TRY(c, dPresent.alloc(256, d0.device));
    CUDA_TRY(c, cudaMemcpyAsync(dIn.p, organ_in, n, cudaMemcpyHostToDevice, d0.stream));
    CUDA_TRY(c, cudaMemsetAsync(dPresent.p, 0, 256 * sizeof(unsigned int), d0.stream));
    organPresenceKernel<<<blocks, 256, 0, d0.stream>>>(dIn.p, n, dPresent.p);
    CUDA_TRY(c, cudaGetLastError());
    unsigned int hp[256];
    CUDA_TRY(c, cudaMemcpyAsync(hp, dPresent.p, sizeof(hp), cudaMemcpyDeviceToHost, d0.stream));
    // the output buffers while the presence scan runs
    CUDA_TRY(c, dOrg.alloc(n + 16, d0.device));
    CUDA_TRY(c, dMat.alloc(n + 16, d0.device));
    CUDA_TRY(c, dDen.alloc(n, d0.device));
    CUDA_TRY(c, cudaStreamSynchronize(d0.stream));
    uint8_t present[256];
    for (int v = 0; v < 256; ++v)
        present[v] = hp[v] ? 1 : 0;
    dxb_icrp* plan = nullptr;
    const int rc = dxb_icrp_plan(&plan, organs_dat, media_dat, remove_arms, present);
    if (rc != DXB_OK)
        return fail(c, rc, "icrp_import: the organ or the media table holds no valid line");
    std::vector<unsigned char> luts(512);
    std::vector<double> den(256);
    for (int v = 0; v < 256; ++v) {
        luts[v] = plan->plan.organLut[v];
        luts[256 + v] = plan->plan.materialLut[v];
        den[v] = plan->plan.densityLut[v];
    }
    cudaError_t e = dLut.upload(luts, d0.device, d0.stream);
    if (e == cudaSuccess)
        e = dDenLut.upload(den, d0.device, d0.stream);
    if (e == cudaSuccess) {
        icrpRemapKernel<<<blocks, 256, 0, d0.stream>>>(dIn.p, n, dLut.p, dLut.p + 256, dDenLut.p, dOrg.p, dMat.p, dDen.p);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(organ_out, dOrg.p, n, cudaMemcpyDeviceToHost, d0.stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(material_out, dMat.p, n, cudaMemcpyDeviceToHost, d0.stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(density_out, dDen.p, n * sizeof(double), cudaMemcpyDeviceToHost, d0.stream);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(d0.stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        dxb_icrp_destroy(plan);
        return fail(c, DXB_ECUDA, std::string("icrp_import: ") + cudaGetErrorString(e));
    }
    *plan_out = plan;
    return DXB_OK;
}

} // extern "C"
