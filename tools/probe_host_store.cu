// probe_host_store.cu -- how fast can a KERNEL write 17.7 MB of ranges into page-locked host memory?
// The end-to-end scanMany step is bound by exactly this (0.376 ms, of which the march is 0.07 ms):
// the march kernel's zero-copy stores reach ~47 GB/s where the copy engine reaches ~55 GB/s.  This probe asks
// whether a different store instruction closes the gap: 4-, 16- and 32-byte stores per thread, and TMA bulk stores
// (cp.async.bulk.global.shared::cta) of 512 B .. 32 KB blocks staged in shared memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o probe_host_store probe_host_store.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void st4_kernel(const float *__restrict__ src, float *__restrict__ dst, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) __stcs(dst + i, src[i]);
}

__global__ void st16_kernel(const float4 *__restrict__ src, float4 *__restrict__ dst, int64_t n4)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) __stcs(dst + i, src[i]);
}

__global__ void st32_kernel(const float4 *__restrict__ src, float *__restrict__ dst, int64_t n8)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n8) {
        const float4 a = src[2 * i], b = src[2 * i + 1];
        asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 8 * i), "f"(a.x), "f"(a.y),
                     "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
                     : "memory");
    }
}

// One CTA per block of BYTES: global -> shared with 16-byte loads, then ONE bulk store shared -> host.
template <int BYTES>
__global__ void __launch_bounds__(256) tma_kernel(const float4 *__restrict__ src, char *__restrict__ dst, int64_t nblocks)
{
    extern __shared__ __align__(128) char smem[];
    const int64_t b = blockIdx.x;
    if (b >= nblocks) return;
    const float4 *s = src + b * (BYTES / 16);
    for (int i = threadIdx.x; i < BYTES / 16; i += blockDim.x) reinterpret_cast<float4 *>(smem)[i] = s[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + b * (int64_t)BYTES), "r"(sa), "r"(BYTES)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// Persistent form: G CTAs, each walks blocks b = blockIdx.x, +G, ... with two shared-memory buffers, so the
// bulk store of one block is in flight while the next is staged.
template <int BYTES>
__global__ void __launch_bounds__(256) tma_persistent_kernel(const float4 *__restrict__ src, char *__restrict__ dst, int64_t nblocks)
{
    extern __shared__ __align__(128) char smem[];
    int buf = 0;
    for (int64_t b = blockIdx.x; b < nblocks; b += gridDim.x, buf ^= 1) {
        char *sm = smem + buf * BYTES;
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store that last read this buffer
        __syncthreads();
        const float4 *s = src + b * (BYTES / 16);
        for (int i = threadIdx.x; i < BYTES / 16; i += blockDim.x) reinterpret_cast<float4 *>(sm)[i] = s[i];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sm);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + b * (int64_t)BYTES), "r"(sa), "r"(BYTES)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
static float time_it(F f, cudaStream_t s, int reps)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) f();
    CK(cudaStreamSynchronize(s));
    CK(cudaEventRecord(a, s));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(b, s));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    CK(cudaEventDestroy(a));
    CK(cudaEventDestroy(b));
    return ms / reps;
}

static void report(const char *name, float ms, size_t bytes, const float *h, const float *ref, size_t n)
{
    const bool ok = std::memcmp(h, ref, n * sizeof(float)) == 0;
    printf("{\"probe\": \"host_store\", \"variant\": \"%s\", \"ms\": %.4f, \"gb_per_s\": %.2f, \"identical\": %s}\n", name, ms,
           bytes / ms * 1e-6, ok ? "true" : "false");
    fflush(stdout);
}

template <int BYTES>
static void run_tma(const float *d_src, float *h_dst, float *d_alias, const float *ref, size_t n, cudaStream_t s, int reps)
{
    const size_t bytes = n * sizeof(float);
    const int64_t nblocks = bytes / BYTES;   // n is chosen as a multiple of 32 KB
    char name[64];
    CK(cudaFuncSetAttribute(tma_kernel<BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, BYTES));
    std::memset(h_dst, 0, bytes);
    float ms = time_it([&] { tma_kernel<BYTES><<<(unsigned)nblocks, 256, BYTES, s>>>(reinterpret_cast<const float4 *>(d_src), reinterpret_cast<char *>(d_alias), nblocks); }, s, reps);
    CK(cudaGetLastError());
    snprintf(name, sizeof name, "tma_bulk_%d", BYTES);
    report(name, ms, bytes, h_dst, ref, n);
    CK(cudaFuncSetAttribute(tma_persistent_kernel<BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * BYTES));
    for (int per_sm = 1; per_sm <= 4; per_sm *= 2) {
        std::memset(h_dst, 0, bytes);
        const unsigned grid = 148 * per_sm;
        ms = time_it([&] { tma_persistent_kernel<BYTES><<<grid, 256, 2 * BYTES, s>>>(reinterpret_cast<const float4 *>(d_src), reinterpret_cast<char *>(d_alias), nblocks); }, s, reps);
        CK(cudaGetLastError());
        snprintf(name, sizeof name, "tma_persistent_%d_x%d", BYTES, per_sm);
        report(name, ms, bytes, h_dst, ref, n);
    }
}

int main()
{
    const size_t n = (size_t)4096 * 1080;   // BASELINE config 2: 4 423 680 ranges = 17 694 720 B = 540 * 32 KB
    const size_t bytes = n * sizeof(float);
    const int reps = 20;
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    float *d_src, *h_dst, *d_alias, *ref;
    CK(cudaMalloc(&d_src, bytes));
    CK(cudaHostAlloc(&h_dst, bytes, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void **)&d_alias, h_dst, 0));
    ref = (float *)malloc(bytes);
    for (size_t i = 0; i < n; ++i) ref[i] = (float)(i % 1000003) * 0.25f;
    CK(cudaMemcpy(d_src, ref, bytes, cudaMemcpyHostToDevice));

    float ms = time_it([&] { CK(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, s)); }, s, reps);
    report("copy_engine", ms, bytes, h_dst, ref, n);

    std::memset(h_dst, 0, bytes);
    ms = time_it([&] { st4_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_src, d_alias, (int64_t)n); }, s, reps);
    report("st_4B_cta128", ms, bytes, h_dst, ref, n);

    std::memset(h_dst, 0, bytes);
    ms = time_it([&] { st16_kernel<<<(unsigned)((n / 4 + 127) / 128), 128, 0, s>>>(reinterpret_cast<const float4 *>(d_src), reinterpret_cast<float4 *>(d_alias), (int64_t)(n / 4)); }, s, reps);
    report("st_16B_cta128", ms, bytes, h_dst, ref, n);

    std::memset(h_dst, 0, bytes);
    ms = time_it([&] { st32_kernel<<<(unsigned)((n / 8 + 127) / 128), 128, 0, s>>>(reinterpret_cast<const float4 *>(d_src), d_alias, (int64_t)(n / 8)); }, s, reps);
    report("st_32B_cta128", ms, bytes, h_dst, ref, n);

    run_tma<512>(d_src, h_dst, d_alias, ref, n, s, reps);
    run_tma<2048>(d_src, h_dst, d_alias, ref, n, s, reps);
    run_tma<8192>(d_src, h_dst, d_alias, ref, n, s, reps);
    run_tma<32768>(d_src, h_dst, d_alias, ref, n, s, reps);
    return 0;
}
