// kernels.hpp — host-callable launchers implemented in transport.cu
#pragma once
#include "device_types.cuh"

namespace dxb {
cudaError_t launchTransport(const RunParams& p, int mode, bool calib, const LaunchConfig& cfg, cudaStream_t stream);
int transportOccupancy(int mode, bool calib, bool smemTable, int threads, size_t smem);
// lane-multiplexed kernel (transport_mux.cu), cfg.slots photons per lane
cudaError_t launchTransportMux(const RunParams& p, int mode, bool calib, const LaunchConfig& cfg, cudaStream_t stream);
int transportMuxSlots(int mode, bool calib, bool smemTable, int slots); // slot count of the variant that will run
int transportMuxOccupancy(int mode, bool calib, bool smemTable, int slots, int threads, size_t smem);
// block-pooled kernel (transport_pool.cu), cfg.slots photon slots per lane class
cudaError_t launchTransportPool(const RunParams& p, int mode, bool calib, const LaunchConfig& cfg, cudaStream_t stream);
int transportPoolSlots(int mode, bool calib, bool smemTable, int slots, bool localMajorant = false, bool brickFilter = false, bool denseBox = false);
int transportPoolOccupancy(int mode, bool calib, bool smemTable, int slots, int threads, size_t smem, int minBlocks, bool localMajorant = false, bool brickFilter = false, bool denseBox = false);
void setLaunchSmCount(int sms);
void launchHoleSums(const unsigned long long* tally, const signed char* hole, size_t n, unsigned long long* sums, cudaStream_t s);
void launchDenseBox(const unsigned int* voxels, int nx, int ny, int nz, const unsigned int* thin, int* box, cudaStream_t s);
void launchOutsideMax(const unsigned int* voxels, int nx, int ny, int nz, const int* box, unsigned int* out, cudaStream_t s);
void launchSlabMax(const unsigned int* voxels, size_t layerSize, int nz, int shift, int nslabs, unsigned int* out, cudaStream_t s);
void launchBrickBound(const unsigned int* voxels, int nx, int ny, int nz, int shift, int nbx, int nby, int nbz, const float* tot,
    const float* majorant, int n_mat, unsigned char* out, cudaStream_t s);
void launchPackVoxels(const double* density, const unsigned char* material, unsigned int* out, size_t n, unsigned int* maxBits, cudaStream_t s);
void launchMajorant(const float* tot, const unsigned int* maxBits, int n_mat, float* majorant, cudaStream_t s);
void launchEnergyToDose(const unsigned long long* tally, const unsigned int* voxels, double* dose, double* variance,
    unsigned long long* events, size_t n, double inv_e, double inv_e2, double factor, double vol, cudaStream_t s);
void launchFusedReduceToDose(const unsigned long long* tally, bool multicast, const unsigned long long* const* peers, int n_peers,
    int first_peer, const unsigned int* voxels, double* dose, double* variance, unsigned long long* events, size_t begin, size_t end, double inv_e,
    double inv_e2, double factor, double vol, cudaStream_t s);
void launchTallyToEnergy(const unsigned long long* tally, double* e, double* e2, unsigned long long* cnt, size_t n,
    double inv_e, double inv_e2, cudaStream_t s);
void launchPeerReduce(unsigned long long* dst, const unsigned long long* const* peers, int n_peers, size_t n_words, cudaStream_t s);
void launchPostprocess(const double* in, const unsigned int* voxels, double* out, size_t n, int maskAir, double scale, cudaStream_t s);
void launchU64ToDouble(const unsigned long long* in, const unsigned int* voxels, double* out, size_t n, int maskAir, cudaStream_t s);
void launchMax(const double* in, const unsigned int* voxels, size_t n, int maskAir, unsigned long long* outBits, cudaStream_t s);
void launchOrganDose(const double* dose, const double* variance, const unsigned int* voxels, const unsigned char* organ, size_t n,
    double vol, double* energy, double* mass, unsigned long long* count, double* varEnergy, cudaStream_t s);
void launchAttenuationProbe(const TablesDev& tab, int mat, const float* energy, int n, float* out4, cudaStream_t s);
void launchMajorantProbe(const float* majorant, const float* energy, int n, float* out, cudaStream_t s);
void launchSegment(const double* hu, size_t n, const double* sep, int n_sep, const double* matAtt, double waterAttDens,
    double airAttDens, unsigned char* material, double* density, cudaStream_t s);
}
