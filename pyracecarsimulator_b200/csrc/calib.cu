// calib.cu -- the denominator of the L2-gather roofline:
// throughput of independent random 4-byte gathers from a distance-field-sized, L2-resident
// buffer on this GPU.  Every march step is one such gather; nothing here is on the product path.
#include "common.h"

namespace {

constexpr int ILP = 8;

__global__ void __launch_bounds__(256)
gather_kernel(const float *__restrict__ buf, uint32_t n, int rounds, float *__restrict__ sink)
{
    uint32_t s[ILP];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s[i] = (tid * ILP + i) * 2654435761u + 12345u;
    float acc = 0.f;
    for (int r = 0; r < rounds; ++r) {
        float v[ILP];
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            s[i] = s[i] * 1664525u + 1013904223u;
            const uint32_t idx = (uint32_t)(((uint64_t)s[i] * n) >> 32);
            v[i] = __ldg(buf + idx);
        }
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc += v[i];
    }
    if (acc == 1234.5678f) sink[0] = acc;  // keep the loads alive
}

// Random FULL-SECTOR reads: every pair of lanes loads the two 16-byte halves of one random 32-byte
// sector (ld.global.cg: L2 only, nothing is served by L1), so a warp instruction asks L2 for 16
// distinct sectors and uses every byte of them.  Sectors per second of this kernel is the ceiling
// the L2 -> SM path offers to a random-access reader; the march's lts__t_sectors are compared with it.
__global__ void __launch_bounds__(256)
sector_kernel(const float4 *__restrict__ buf, uint32_t n_sectors, int rounds, float *__restrict__ sink)
{
    uint32_t s[ILP];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t pair = tid >> 1, half = tid & 1u;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s[i] = (pair * ILP + i) * 2654435761u + 12345u;
    float acc = 0.f;
    for (int r = 0; r < rounds; ++r) {
        float4 v[ILP];
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            s[i] = s[i] * 1664525u + 1013904223u;
            const uint32_t sec = (uint32_t)(((uint64_t)s[i] * n_sectors) >> 32);
            v[i] = __ldcg(buf + 2 * (size_t)sec + half);
        }
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    if (acc == 1234.5678f) sink[0] = acc;  // keep the loads alive
}

}  // namespace

extern "C" RL_API int32_t rl_l2_sector_bandwidth(int32_t device, int64_t buffer_bytes, int32_t rounds,
                                                 int32_t iters, float *gsectors_per_s)
{
    if (!gsectors_per_s || buffer_bytes < 4096 || rounds <= 0 || iters <= 0)
        return rl::fail(RL_ERR_BAD_ARG, "rl_l2_sector_bandwidth: bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
        return rl::fail(RL_ERR_NO_DEVICE, "rl_l2_sector_bandwidth: no such CUDA device");
    rl::DeviceGuard guard(device);
    const uint32_t n_sectors = (uint32_t)(buffer_bytes / 32);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    float4 *buf = nullptr;
    float *sink = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t e = cudaMalloc(&buf, (size_t)n_sectors * 32);
    if (e == cudaSuccess) e = cudaMalloc(&sink, 4);
    if (e == cudaSuccess) e = cudaMemset(buf, 0, (size_t)n_sectors * 32);
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    const int blocks = sms * 8;
    float ms = 0.f;
    if (e == cudaSuccess) {
        sector_kernel<<<blocks, 256>>>(buf, n_sectors, rounds, sink);  // warm-up: pulls the buffer into L2
        cudaEventRecord(e0, 0);
        for (int i = 0; i < iters; ++i) sector_kernel<<<blocks, 256>>>(buf, n_sectors, rounds, sink);
        cudaEventRecord(e1, 0);
        e = cudaEventSynchronize(e1);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (e == cudaSuccess) e = cudaGetLastError();
    }
    cudaFree(buf); cudaFree(sink);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (e != cudaSuccess) return rl::fail(RL_ERR_CUDA, std::string("rl_l2_sector_bandwidth: ") + cudaGetErrorString(e));
    const double sectors = (double)blocks * 128 * ILP * rounds * iters;   // one sector per lane pair and load
    *gsectors_per_s = (float)(sectors / (ms * 1e-3) / 1e9);
    return RL_OK;
}

extern "C" RL_API int32_t rl_gather_bandwidth(int32_t device, int64_t buffer_bytes, int32_t rounds,
                                       int32_t iters, float *gbytes_per_s)
{
    if (!gbytes_per_s || buffer_bytes < 4096 || rounds <= 0 || iters <= 0)
        return rl::fail(RL_ERR_BAD_ARG, "rl_gather_bandwidth: bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
        return rl::fail(RL_ERR_NO_DEVICE, "rl_gather_bandwidth: no such CUDA device");
    rl::DeviceGuard guard(device);
    const uint32_t n = (uint32_t)(buffer_bytes / 4);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    float *buf = nullptr, *sink = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t e = cudaMalloc(&buf, (size_t)n * 4);
    if (e == cudaSuccess) e = cudaMalloc(&sink, 4);
    if (e == cudaSuccess) e = cudaMemset(buf, 0, (size_t)n * 4);
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    const int blocks = sms * 8;  // 8 CTAs x 256 threads = 64 resident warps per SM
    float ms = 0.f;
    if (e == cudaSuccess) {
        gather_kernel<<<blocks, 256>>>(buf, n, rounds, sink);  // warm-up: pulls the buffer into L2
        cudaEventRecord(e0, 0);
        for (int i = 0; i < iters; ++i) gather_kernel<<<blocks, 256>>>(buf, n, rounds, sink);
        cudaEventRecord(e1, 0);
        e = cudaEventSynchronize(e1);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
        if (e == cudaSuccess) e = cudaGetLastError();
    }
    cudaFree(buf); cudaFree(sink);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (e != cudaSuccess) return rl::fail(RL_ERR_CUDA, std::string("rl_gather_bandwidth: ") + cudaGetErrorString(e));
    const double gathers = (double)blocks * 256 * ILP * rounds * iters;
    *gbytes_per_s = (float)(gathers * 4.0 / (ms * 1e-3) / 1e9);
    return RL_OK;
}

// Demote every persisting L2 line of the current context to normal (cudaCtxResetPersistingL2Cache), so
// that a benchmark's L2 flush also evicts the distance field the march launches pin with their
// access-policy window.
extern "C" RL_API int32_t rl_l2_reset_persisting(int32_t device)
{
    rl::DeviceGuard guard(device);
    if (!guard.ok) return rl::fail(RL_ERR_NO_DEVICE, "rl_l2_reset_persisting: bad device");
    RL_CUDA(cudaCtxResetPersistingL2Cache());
    return RL_OK;
}
