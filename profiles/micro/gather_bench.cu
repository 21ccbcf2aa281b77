// gather_bench.cu — measures what B200 sustains for the access pattern of the Woodcock step: independent random
// 4-byte gathers (ld.global.cg, one 32-byte sector each) over an array much larger than L2, with K loads in flight per
// thread.  This is the ceiling the transport kernel's voxel fetches are compared with in DESIGN.md (sector roofline).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/micro/gather_bench profiles/micro/gather_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned int mix(unsigned int x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

template <int K>
__global__ void __launch_bounds__(256) gather(const unsigned int* __restrict__ a, unsigned int n, int iters, unsigned int* out)
{
    unsigned int s = mix(blockIdx.x * blockDim.x + threadIdx.x + 1u);
    unsigned int acc = 0;
    for (int it = 0; it < iters; ++it) {
        unsigned int idx[K], v[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            s = mix(s + 0x9e3779b9u);
            idx[k] = static_cast<unsigned int>((static_cast<unsigned long long>(s) * n) >> 32);
        }
#pragma unroll
        for (int k = 0; k < K; ++k)
            v[k] = __ldcg(a + idx[k]);
#pragma unroll
        for (int k = 0; k < K; ++k)
            acc += v[k];
    }
    if (acc == 0x12345678u)
        out[0] = acc;
}

template <int K>
void run(const unsigned int* a, unsigned int n, unsigned int* out, int blocksPerSm, const char* tag)
{
    const int iters = 2048 / K;
    const int blocks = 148 * blocksPerSm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    gather<K><<<blocks, 256>>>(a, n, iters, out);
    cudaEventRecord(e0);
    for (int r = 0; r < 3; ++r)
        gather<K><<<blocks, 256>>>(a, n, iters, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double loads = 3.0 * blocks * 256.0 * iters * K;
    printf("%s array=%7.1f MB K=%d blocks/SM=%d: %.1f G gathers/s, %.0f GB/s of 32-byte sectors\n", tag, n * 4.0 / 1e6, K, blocksPerSm,
        loads / ms / 1e6, loads * 32.0 / ms / 1e6);
}

int main(int argc, char** argv)
{
    if (argc > 1) {
        // L2 fetch granularity hint (bytes: 32, 64 or 128)
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, static_cast<size_t>(atoi(argv[1])));
        size_t g = 0;
        cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
        printf("cudaLimitMaxL2FetchGranularity requested %s -> %zu (%s)\n", argv[1], g, cudaGetErrorString(e));
    }
    const unsigned int nBig = 78643200u; // C2: 512 x 512 x 300 voxels x 4 B = 315 MB
    unsigned int *a, *out;
    cudaMalloc(&a, static_cast<size_t>(nBig) * 4 * 8);
    cudaMalloc(&out, 4);
    cudaMemset(a, 1, static_cast<size_t>(nBig) * 4 * 8);
    for (int bps = 4; bps <= 8; bps += 4) {
        run<1>(a, nBig, out, bps, "C2 voxels ");
        run<4>(a, nBig, out, bps, "C2 voxels ");
    }
    run<4>(a, nBig * 8u, out, 8, "2.5 GB     ");
    run<4>(a, nBig / 8u, out, 8, "L2 resident");
    run<4>(a, nBig / 2u, out, 8, "157 MB     ");
    run<4>(a, nBig / 4u, out, 8, "79 MB      ");
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
