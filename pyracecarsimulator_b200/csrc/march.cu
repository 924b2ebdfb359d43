// march.cu -- the ray-marching hot path and its C-ABI entry points (north_star (b)).
// Replaces range_libc's RayMarching / RayMarchingGPU calc_range_many (2-arg and the fork's
// 4-arg fan) and calc_range_repeat_angles; reference call sites scripts/scan_simulator.py:103-106,
// :130-133 and scripts/two_player/scan.py:69-70.
#include <cstring>
#include <new>

#include "glibc_trig.cuh"
#include "march.cuh"

namespace {

using rl::GridPose;
using rl::MarchParams;

constexpr int WARPS_PER_CTA = 8;
constexpr int CTA_THREADS = WARPS_PER_CTA * 32;

template <bool COUNT>
__device__ __forceinline__ void flush_steps(uint32_t steps, unsigned long long *counter)
{
    if (COUNT) {
        steps = __reduce_add_sync(0xffffffffu, steps);
        if ((threadIdx.x & 31) == 0 && steps) atomicAdd(counter, (unsigned long long)steps);
    }
}

// ---- one (x, y, theta) row per ray (upstream 2-arg calc_range_many) ----
template <bool COUNT>
__global__ void __launch_bounds__(CTA_THREADS)
march_many_kernel(MarchParams P, const float *__restrict__ ins, float *__restrict__ outs,
                  int64_t n, unsigned long long *counter)
{
    const int64_t i = (int64_t)blockIdx.x * CTA_THREADS + threadIdx.x;
    uint32_t steps = 0;
    if (i < n) {
        const GridPose g = rl::world_to_grid(P.w, ins[3 * i], ins[3 * i + 1], ins[3 * i + 2]);
        float s, c;
        rl::glibc_sincosf(g.theta, &s, &c);
        outs[i] = __fmul_rn(rl::march_ray<COUNT>(P, g.y, g.x, c, s, steps), P.w.scale);
    }
    flush_steps<COUNT>(steps, counter);
}

// ---- one warp per (pose, beam segment), lanes over beams ----
// FAN:   beam j heads theta + fmaf(j, fov/num_beams, -fov/2)   (fork's 4-arg calc_range_many)
// !FAN:  beam a heads theta + angles[a]                        (calc_range_repeat_angles)
template <bool FAN, bool COUNT>
__global__ void __launch_bounds__(CTA_THREADS)
march_pose_kernel(MarchParams P, const float *__restrict__ poses, int64_t pose_stride_floats,
                  const float *__restrict__ angles, float *__restrict__ outs, int64_t num_poses,
                  int num_beams, int segs_per_pose, int seg_len, float fov,
                  unsigned long long *counter)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
    const int64_t k = warp / segs_per_pose;
    uint32_t steps = 0;
    if (k < num_poses) {
        const int seg = (int)(warp - k * segs_per_pose);
        const float *p = poses + k * pose_stride_floats;
        const GridPose g = rl::world_to_grid(P.w, __ldg(p), __ldg(p + 1), __ldg(p + 2));
        const float thw = __ldg(p + 2);
        const float inc = fov / (float)num_beams;
        const float half = -0.5f * fov;
        const int j_end = min(num_beams, (seg + 1) * seg_len);
        float *o = outs + k * num_beams;
        for (int j = seg * seg_len + lane; j < j_end; j += 32) {
            float thg;
            if (FAN) thg = __fadd_rn(-__fadd_rn(thw, fmaf((float)j, inc, half)), P.w.rotation_const);
            else thg = __fsub_rn(g.theta, __ldg(angles + j));
            float s, c;
            rl::glibc_sincosf(thg, &s, &c);
            o[j] = __fmul_rn(rl::march_ray<COUNT>(P, g.y, g.x, c, s, steps), P.w.scale);
        }
    }
    flush_steps<COUNT>(steps, counter);
}

__global__ void trig_probe_kernel(const float *__restrict__ in, float *__restrict__ s,
                                  float *__restrict__ c, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rl::glibc_sincosf(in[i], s + i, c + i);
}

}  // namespace

struct rl_marcher {
    const rl_map *map = nullptr;
    MarchParams P{};
    uint32_t flags = 0;
    int sm_count = 148;
    // host-variant staging (guarded by mu)
    std::mutex mu;
    cudaStream_t stream = nullptr;
    float *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr, *d_angles = nullptr;
    size_t cap_in = 0, cap_out = 0, cap_angles = 0;  // floats
    // optional step counter
    bool count = false;
    unsigned long long *d_steps = nullptr;
};

namespace {

constexpr size_t HOST_CHUNK_RAYS = (size_t)16 << 20;  // 64 MiB of ranges per staged chunk

int32_t ensure(float **h, float **d, size_t *cap, size_t want)
{
    if (*cap >= want) return RL_OK;
    if (h) { cudaFreeHost(*h); *h = nullptr; }
    cudaFree(*d); *d = nullptr; *cap = 0;
    if (h) RL_CUDA(cudaMallocHost(h, want * sizeof(float)));
    RL_CUDA(cudaMalloc(d, want * sizeof(float)));
    *cap = want;
    return RL_OK;
}

bool is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

int32_t launch_many(rl_marcher *m, const float *d_ins, float *d_outs, int64_t n, cudaStream_t s)
{
    if (n == 0) return RL_OK;
    const int64_t blocks = (n + CTA_THREADS - 1) / CTA_THREADS;
    if (blocks > 0x7fffffffLL) return rl::fail(RL_ERR_BAD_ARG, "calc_range_many: too many rays for one call");
    if (m->count) march_many_kernel<true><<<(unsigned)blocks, CTA_THREADS, 0, s>>>(m->P, d_ins, d_outs, n, m->d_steps);
    else march_many_kernel<false><<<(unsigned)blocks, CTA_THREADS, 0, s>>>(m->P, d_ins, d_outs, n, nullptr);
    RL_CUDA(cudaGetLastError());
    return RL_OK;
}

// Split each pose's beams into segments so that small batches still fill the machine.
void plan_segments(const rl_marcher *m, int64_t num_poses, int num_beams, int *segs, int *seg_len)
{
    const int groups = (num_beams + 31) / 32;               // 32-beam groups per pose
    const int64_t want_warps = (int64_t)m->sm_count * 64 * 2;  // two full waves of resident warps
    int64_t s = (want_warps + num_poses - 1) / num_poses;
    if (s < 1) s = 1;
    if (s > groups) s = groups;
    int gl = (groups + (int)s - 1) / (int)s;                 // groups per segment
    *seg_len = gl * 32;
    *segs = (groups + gl - 1) / gl;
}

template <bool FAN>
int32_t launch_pose(rl_marcher *m, const float *d_poses, int64_t stride_rows, const float *d_angles,
                    float *d_outs, int64_t num_poses, int num_beams, float fov, cudaStream_t s)
{
    if (num_poses == 0 || num_beams == 0) return RL_OK;
    int segs, seg_len;
    plan_segments(m, num_poses, num_beams, &segs, &seg_len);
    const int64_t warps = num_poses * segs;
    const int64_t blocks = (warps + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    if (blocks > 0x7fffffffLL) return rl::fail(RL_ERR_BAD_ARG, "calc_range: too many poses for one call");
    if (m->count)
        march_pose_kernel<FAN, true><<<(unsigned)blocks, CTA_THREADS, 0, s>>>(
            m->P, d_poses, stride_rows * 3, d_angles, d_outs, num_poses, num_beams, segs, seg_len, fov, m->d_steps);
    else
        march_pose_kernel<FAN, false><<<(unsigned)blocks, CTA_THREADS, 0, s>>>(
            m->P, d_poses, stride_rows * 3, d_angles, d_outs, num_poses, num_beams, segs, seg_len, fov, nullptr);
    RL_CUDA(cudaGetLastError());
    return RL_OK;
}

// D2H of `n` floats into user memory: direct DMA when the user's buffer is pinned, else through
// the marcher's pinned staging buffer.
int32_t fetch(rl_marcher *m, float *outs, size_t n)
{
    if (is_pinned(outs)) {
        RL_CUDA(cudaMemcpyAsync(outs, m->d_out, n * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
        RL_CUDA(cudaStreamSynchronize(m->stream));
    } else {
        RL_CUDA(cudaMemcpyAsync(m->h_out, m->d_out, n * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
        RL_CUDA(cudaStreamSynchronize(m->stream));
        std::memcpy(outs, m->h_out, n * sizeof(float));
    }
    return RL_OK;
}

}  // namespace

extern "C" {

int32_t rl_probe_sincosf(const float *d_in, float *d_sin, float *d_cos, int64_t n, void *stream)
{
    if (n < 0 || (n > 0 && (!d_in || !d_sin || !d_cos))) return rl::fail(RL_ERR_BAD_ARG, "rl_probe_sincosf: bad argument");
    if (n == 0) return RL_OK;
    trig_probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_in, d_sin, d_cos, n);
    RL_CUDA(cudaGetLastError());
    return RL_OK;
}

int32_t rl_marcher_create(const rl_map *map, float max_range_px, uint32_t flags, rl_marcher **out)
{
    if (!map || !out) return rl::fail(RL_ERR_BAD_ARG, "rl_marcher_create: null pointer");
    if (!(max_range_px > 0.0f)) return rl::fail(RL_ERR_BAD_ARG, "rl_marcher_create: max_range_px must be > 0");
    rl::DeviceGuard guard(map->device);
    if (!guard.ok) return rl::fail(RL_ERR_CUDA, "rl_marcher_create: cudaSetDevice failed");
    rl_marcher *m = new (std::nothrow) rl_marcher();
    if (!m) return rl::fail(RL_ERR_OOM, "rl_marcher_create: host allocation failed");
    rl_map_retain(map);
    m->map = map;
    m->flags = flags;
    m->P.dist = map->d_dist;
    m->P.rows = map->rows;
    m->P.cols = map->cols;
    m->P.frows = (float)map->rows;
    m->P.fcols = (float)map->cols;
    m->P.max_range = max_range_px;
    m->P.w = map->world;
    cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, map->device);
    cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_steps, sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(m->d_steps, 0, sizeof(unsigned long long));
    if (e != cudaSuccess) {
        rl_marcher_destroy(m);
        return rl::fail(RL_ERR_CUDA, std::string("rl_marcher_create: ") + cudaGetErrorString(e));
    }
    *out = m;
    return RL_OK;
}

int32_t rl_marcher_destroy(rl_marcher *m)
{
    if (!m) return rl::fail(RL_ERR_BAD_ARG, "rl_marcher_destroy: null marcher");
    {
        rl::DeviceGuard guard(m->map->device);
        if (m->stream) { cudaStreamSynchronize(m->stream); cudaStreamDestroy(m->stream); }
        cudaFreeHost(m->h_in); cudaFreeHost(m->h_out);
        cudaFree(m->d_in); cudaFree(m->d_out); cudaFree(m->d_angles); cudaFree(m->d_steps);
    }
    rl_map_release(m->map);
    delete m;
    return RL_OK;
}

int32_t rl_marcher_count_steps(rl_marcher *m, int32_t enable)
{
    if (!m) return rl::fail(RL_ERR_BAD_ARG, "rl_marcher_count_steps: null marcher");
    m->count = enable != 0;
    return RL_OK;
}

int32_t rl_marcher_last_steps(rl_marcher *m, uint64_t *steps)
{
    if (!m || !steps) return rl::fail(RL_ERR_BAD_ARG, "rl_marcher_last_steps: null pointer");
    rl::DeviceGuard guard(m->map->device);
    RL_CUDA(cudaDeviceSynchronize());
    unsigned long long v = 0;
    RL_CUDA(cudaMemcpy(&v, m->d_steps, sizeof(v), cudaMemcpyDeviceToHost));
    RL_CUDA(cudaMemset(m->d_steps, 0, sizeof(v)));
    *steps = v;
    return RL_OK;
}

int32_t rl_calc_range_many(rl_marcher *m, const float *d_ins, float *d_outs, int64_t n, void *stream)
{
    if (!m || n < 0 || (n > 0 && (!d_ins || !d_outs))) return rl::fail(RL_ERR_BAD_ARG, "rl_calc_range_many: bad argument");
    rl::DeviceGuard guard(m->map->device);
    return launch_many(m, d_ins, d_outs, n, (cudaStream_t)stream);
}

int32_t rl_calc_range_fan(rl_marcher *m, const float *d_poses, int64_t pose_stride_rows, float *d_outs,
                          int64_t num_poses, int32_t num_rays, float fov, void *stream)
{
    if (!m || num_poses < 0 || num_rays <= 0 || pose_stride_rows < 1 ||
        (num_poses > 0 && (!d_poses || !d_outs)))
        return rl::fail(RL_ERR_BAD_ARG, "rl_calc_range_fan: bad argument");
    rl::DeviceGuard guard(m->map->device);
    return launch_pose<true>(m, d_poses, pose_stride_rows, nullptr, d_outs, num_poses, num_rays, fov,
                             (cudaStream_t)stream);
}

int32_t rl_calc_range_repeat_angles(rl_marcher *m, const float *d_poses, const float *d_angles,
                                    float *d_outs, int64_t num_poses, int32_t num_angles, void *stream)
{
    if (!m || num_poses < 0 || num_angles < 0 ||
        (num_poses > 0 && num_angles > 0 && (!d_poses || !d_angles || !d_outs)))
        return rl::fail(RL_ERR_BAD_ARG, "rl_calc_range_repeat_angles: bad argument");
    rl::DeviceGuard guard(m->map->device);
    return launch_pose<false>(m, d_poses, 1, d_angles, d_outs, num_poses, num_angles, 0.0f,
                              (cudaStream_t)stream);
}

int32_t rl_calc_range_many_host(rl_marcher *m, const float *ins, float *outs, int64_t n)
{
    if (!m || n < 0 || (n > 0 && (!ins || !outs))) return rl::fail(RL_ERR_BAD_ARG, "rl_calc_range_many_host: bad argument");
    rl::DeviceGuard guard(m->map->device);
    std::lock_guard<std::mutex> lock(m->mu);
    for (int64_t b = 0; b < n; b += (int64_t)HOST_CHUNK_RAYS) {
        const size_t c = (size_t)((n - b < (int64_t)HOST_CHUNK_RAYS) ? n - b : HOST_CHUNK_RAYS);
        int32_t rc = ensure(&m->h_in, &m->d_in, &m->cap_in, 3 * c);
        if (rc == RL_OK) rc = ensure(&m->h_out, &m->d_out, &m->cap_out, c);
        if (rc != RL_OK) return rc;
        const float *src = ins + 3 * b;
        if (!is_pinned(src)) { std::memcpy(m->h_in, src, 3 * c * sizeof(float)); src = m->h_in; }
        RL_CUDA(cudaMemcpyAsync(m->d_in, src, 3 * c * sizeof(float), cudaMemcpyHostToDevice, m->stream));
        rc = launch_many(m, m->d_in, m->d_out, (int64_t)c, m->stream);
        if (rc == RL_OK) rc = fetch(m, outs + b, c);
        if (rc != RL_OK) return rc;
    }
    return RL_OK;
}

int32_t rl_calc_range_fan_host(rl_marcher *m, const float *poses, int64_t pose_stride_rows, float *outs,
                               int64_t num_poses, int32_t num_rays, float fov)
{
    if (!m || num_poses < 0 || num_rays <= 0 || pose_stride_rows < 1 || (num_poses > 0 && (!poses || !outs)))
        return rl::fail(RL_ERR_BAD_ARG, "rl_calc_range_fan_host: bad argument");
    rl::DeviceGuard guard(m->map->device);
    std::lock_guard<std::mutex> lock(m->mu);
    int64_t chunk = (int64_t)(HOST_CHUNK_RAYS / (size_t)num_rays);
    if (chunk < 1) chunk = 1;
    for (int64_t b = 0; b < num_poses; b += chunk) {
        const int64_t c = (num_poses - b < chunk) ? num_poses - b : chunk;
        int32_t rc = ensure(&m->h_in, &m->d_in, &m->cap_in, 3 * (size_t)c);
        if (rc == RL_OK) rc = ensure(&m->h_out, &m->d_out, &m->cap_out, (size_t)c * num_rays);
        if (rc != RL_OK) return rc;
        // only row k*pose_stride_rows of each pose's block is meaningful: gather to a compact (c,3)
        for (int64_t k = 0; k < c; ++k)
            std::memcpy(m->h_in + 3 * k, poses + 3 * (b + k) * pose_stride_rows, 3 * sizeof(float));
        RL_CUDA(cudaMemcpyAsync(m->d_in, m->h_in, 3 * (size_t)c * sizeof(float), cudaMemcpyHostToDevice, m->stream));
        rc = launch_pose<true>(m, m->d_in, 1, nullptr, m->d_out, c, num_rays, fov, m->stream);
        if (rc == RL_OK) rc = fetch(m, outs + b * num_rays, (size_t)c * num_rays);
        if (rc != RL_OK) return rc;
    }
    return RL_OK;
}

int32_t rl_calc_range_repeat_angles_host(rl_marcher *m, const float *poses, const float *angles,
                                         float *outs, int64_t num_poses, int32_t num_angles)
{
    if (!m || num_poses < 0 || num_angles < 0 ||
        (num_poses > 0 && num_angles > 0 && (!poses || !angles || !outs)))
        return rl::fail(RL_ERR_BAD_ARG, "rl_calc_range_repeat_angles_host: bad argument");
    if (num_poses == 0 || num_angles == 0) return RL_OK;
    rl::DeviceGuard guard(m->map->device);
    std::lock_guard<std::mutex> lock(m->mu);
    int32_t rc = ensure(nullptr, &m->d_angles, &m->cap_angles, (size_t)num_angles);
    if (rc != RL_OK) return rc;
    RL_CUDA(cudaMemcpyAsync(m->d_angles, angles, (size_t)num_angles * sizeof(float), cudaMemcpyHostToDevice, m->stream));
    RL_CUDA(cudaStreamSynchronize(m->stream));  // `angles` may be pageable and reused by the caller
    int64_t chunk = (int64_t)(HOST_CHUNK_RAYS / (size_t)num_angles);
    if (chunk < 1) chunk = 1;
    for (int64_t b = 0; b < num_poses; b += chunk) {
        const int64_t c = (num_poses - b < chunk) ? num_poses - b : chunk;
        rc = ensure(&m->h_in, &m->d_in, &m->cap_in, 3 * (size_t)c);
        if (rc == RL_OK) rc = ensure(&m->h_out, &m->d_out, &m->cap_out, (size_t)c * num_angles);
        if (rc != RL_OK) return rc;
        const float *src = poses + 3 * b;
        if (!is_pinned(src)) { std::memcpy(m->h_in, src, 3 * (size_t)c * sizeof(float)); src = m->h_in; }
        RL_CUDA(cudaMemcpyAsync(m->d_in, src, 3 * (size_t)c * sizeof(float), cudaMemcpyHostToDevice, m->stream));
        rc = launch_pose<false>(m, m->d_in, 1, m->d_angles, m->d_out, c, num_angles, 0.0f, m->stream);
        if (rc == RL_OK) rc = fetch(m, outs + b * num_angles, (size_t)c * num_angles);
        if (rc != RL_OK) return rc;
    }
    return RL_OK;
}

}  // extern "C"
