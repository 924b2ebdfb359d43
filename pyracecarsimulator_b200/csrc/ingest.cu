// ingest.cu -- map ingest on the GPU: source cells -> occupancy -> exact integer squared
// Euclidean distance transform -> fp32 distance field.  Replaces the host-side work of
// range_libc's PyOMap + DistanceTransform constructors (reference call sites
// scripts/ros_interface.py:80-86, :210 and scripts/scan_simulator.py:72-76).
//
// Kernels (north_star (a): column pass + row pass):
//   edt_classify_kernel  one thread per (column, 64-row segment): classifies every source cell
//                    through a 256-entry bit LUT (map_server thresholds / binarisation / `> 10`
//                    cut are all pure functions of one byte, so the host folds them into the LUT
//                    with the reference's double arithmetic), applies map_server's y-flip, writes
//                    the occupancy byte, records the segment's first/last occupied row.
//   edt_cols_kernel  same mapping: distance g to the nearest occupied cell in the same column
//                    (down sweep seeded from the segments above, up sweep from those below), as
//                    saturating u16.
//   edt_rows_kernel  one CTA per map row with the row's g^2 staged in shared memory: each
//                    thread takes the lower envelope min_k (k^2 + g^2[q +- k]) by an outward
//                    scan that stops as soon as k^2 >= best, which is exact in integers and
//                    costs O(distance) per cell; writes d^2 (int32) and sqrt_rn((float)d^2).
// For max(rows, cols) <= 2896 the reference's float Felzenszwalb transform returns exactly
// this integer d^2, so the fp32 field is bit-identical to the reference's.
#include <cmath>
#include <new>

#include "common.h"

namespace {

constexpr uint32_t G_INF = 0xFFFFu;         // u16 column distance: no occupied cell in the column
constexpr int MAX_SIDE = 16384;             // keeps every real d^2 below RL_DIST2_INF

struct ByteLut { uint32_t w[8]; };          // bit p set <=> source byte p is an occupied cell

__device__ __forceinline__ uint32_t lut_bit(const ByteLut &lut, uint32_t p)
{
    return (lut.w[p >> 5] >> (p & 31u)) & 1u;
}

constexpr int SEG_ROWS = 64;   // rows per column segment: W * ceil(H/64) threads instead of W
constexpr int SEG_NONE_LAST = -1;
constexpr int SEG_NONE_FIRST = 0x3fffffff;

// Column pass, step 1: classify (byte LUT + map_server y-flip), write the occupancy byte, and record
// for every (segment, column) the first and last occupied row inside the segment.
__global__ void __launch_bounds__(128)
edt_classify_kernel(const uint8_t *__restrict__ src, int rows, int cols, int flip, ByteLut lut,
                    uint8_t *__restrict__ occ, int *__restrict__ seg_first, int *__restrict__ seg_last)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int seg = blockIdx.y;
    if (c >= cols) return;
    const int r0 = seg * SEG_ROWS, r1 = min(rows, r0 + SEG_ROWS);
    int first = SEG_NONE_FIRST, last = SEG_NONE_LAST;
#pragma unroll 8
    for (int r = r0; r < r1; ++r) {
        const int sr = flip ? rows - 1 - r : r;
        const uint32_t o = lut_bit(lut, src[(size_t)sr * cols + c]);
        occ[(size_t)r * cols + c] = (uint8_t)o;
        if (o) { first = min(first, r); last = r; }
    }
    seg_first[(size_t)seg * cols + c] = first;
    seg_last[(size_t)seg * cols + c] = last;
}

// Column pass, step 2: distance g (saturating u16) to the nearest occupied cell of the same column:
// down sweep seeded with the last occupied row above the segment, up sweep seeded with the first
// occupied row below it.
__global__ void __launch_bounds__(128)
edt_cols_kernel(const uint8_t *__restrict__ occ, int rows, int cols, int nseg,
                const int *__restrict__ seg_first, const int *__restrict__ seg_last,
                uint16_t *__restrict__ g)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int seg = blockIdx.y;
    if (c >= cols) return;
    const int r0 = seg * SEG_ROWS, r1 = min(rows, r0 + SEG_ROWS);
    int above = SEG_NONE_LAST, below = SEG_NONE_FIRST;
    for (int s2 = seg - 1; s2 >= 0; --s2) {
        const int l = seg_last[(size_t)s2 * cols + c];
        if (l != SEG_NONE_LAST) { above = l; break; }
    }
    for (int s2 = seg + 1; s2 < nseg; ++s2) {
        const int f = seg_first[(size_t)s2 * cols + c];
        if (f != SEG_NONE_FIRST) { below = f; break; }
    }
    uint32_t run = (above == SEG_NONE_LAST) ? G_INF : min((uint32_t)(r0 - 1 - above), G_INF);
#pragma unroll 8
    for (int r = r0; r < r1; ++r) {
        run = occ[(size_t)r * cols + c] ? 0u : min(run + 1u, G_INF);
        g[(size_t)r * cols + c] = (uint16_t)run;
    }
    run = (below == SEG_NONE_FIRST) ? G_INF : min((uint32_t)(below - r1), G_INF);
#pragma unroll 8
    for (int r = r1 - 1; r >= r0; --r) {
        const uint32_t gv = g[(size_t)r * cols + c];
        run = (gv == 0u) ? 0u : min(run + 1u, G_INF);
        if (run < gv) g[(size_t)r * cols + c] = (uint16_t)run;
    }
}

// Row pass.  g2 in shared memory holds g^2, or G2_FAR for columns without any occupied cell;
// G2_FAR + k^2 cannot wrap and stays above every real d^2 (<= 2 * 16384^2), so the scan needs no
// special case for it.  Out-of-row neighbours are clamped to the row ends: the clamped candidate
// k^2 + g2[end] is never below the true candidate (q - end)^2 + g2[end], so the minimum is
// unchanged.  Four offsets are examined per trip (independent shared-memory loads, one exit test).
constexpr uint32_t G2_FAR = 0x3fffffffu;

__global__ void __launch_bounds__(256)
edt_rows_kernel(const uint16_t *__restrict__ g, int rows, int cols,
                int32_t *__restrict__ dist2, float *__restrict__ dist)
{
    extern __shared__ uint32_t g2[];  // cols entries
    const int r = blockIdx.x;
    const uint16_t *grow = g + (size_t)r * cols;
    bool any_near = false;
    for (int q = threadIdx.x; q < cols; q += blockDim.x) {
        const uint32_t v = grow[q];
        any_near |= v < G_INF;
        g2[q] = (v >= G_INF) ? G2_FAR : v * v;
    }
    // A row whose every column is G_INF has no occupied cell within reach in any column (an empty
    // map, or obstacles further than 65534 rows away): skip the outward scans, which would
    // otherwise walk the whole row for every cell.
    if (!__syncthreads_or(any_near)) {
        for (int q = threadIdx.x; q < cols; q += blockDim.x) {
            const size_t o = (size_t)r * cols + q;
            dist2[o] = RL_DIST2_INF;
            dist[o] = sqrtf(1e20f);
        }
        return;
    }
    const int last = cols - 1;
    for (int q = threadIdx.x; q < cols; q += blockDim.x) {
        uint32_t best = g2[q];
        const int reach = max(q, last - q);
        for (int k = 1; k <= reach; k += 4) {
            if ((uint32_t)k * (uint32_t)k >= best) break;
            uint32_t c[8];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int kk = k + u;
                const uint32_t k2 = (uint32_t)kk * (uint32_t)kk;
                c[2 * u] = k2 + g2[max(q - kk, 0)];
                c[2 * u + 1] = k2 + g2[min(q + kk, last)];
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) best = min(best, c[u]);
        }
        const size_t o = (size_t)r * cols + q;
        if (best >= G2_FAR) {
            dist2[o] = RL_DIST2_INF;
            dist[o] = sqrtf(1e20f);  // what the reference's INF = 1e20 transform leaves on an empty map
        } else {
            dist2[o] = (int32_t)best;
            dist[o] = sqrtf((float)best);
        }
    }
}

rl::WorldFrame make_world(double resolution, double ox, double oy, double yaw)
{
    rl::WorldFrame w;
    const double angle = -1.0 * yaw;  // PyOMap: world_angle = -yaw(origin quaternion)
    w.scale = (float)resolution;
    w.angle = (float)angle;
    w.origin_x = (float)ox;
    w.origin_y = (float)oy;
    w.sin_angle = (float)std::sin(angle);
    w.cos_angle = (float)std::cos(angle);
    w.inv_scale = (float)(1.0 / (double)w.scale);
    w.rotation_const = (float)(-1.0 * (double)w.angle - 3.0 * M_PI / 2.0);
    return w;
}

int32_t build_map(const uint8_t *src, int width, int height, int flip, const ByteLut &lut,
                  double resolution, double ox, double oy, double yaw, int device, rl_map **out)
{
    if (!src || !out) return rl::fail(RL_ERR_BAD_ARG, "map ingest: null pointer");
    if (width <= 0 || height <= 0 || width > MAX_SIDE || height > MAX_SIDE)
        return rl::fail(RL_ERR_BAD_ARG, "map ingest: width/height must be in [1, 16384]");
    if (!(resolution > 0.0)) return rl::fail(RL_ERR_BAD_ARG, "map ingest: resolution must be > 0");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return rl::fail(RL_ERR_NO_DEVICE, "map ingest: no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= ndev) return rl::fail(RL_ERR_NO_DEVICE, "map ingest: bad device index");
    rl::DeviceGuard guard(device);
    if (!guard.ok) return rl::fail(RL_ERR_CUDA, "map ingest: cudaSetDevice failed");

    rl_map *m = new (std::nothrow) rl_map();
    if (!m) return rl::fail(RL_ERR_OOM, "map ingest: host allocation failed");
    m->device = device;
    m->rows = height;  // msg.info.height
    m->cols = width;   // msg.info.width
    m->world = make_world(resolution, ox, oy, yaw);
    const size_t n = (size_t)width * height;
    uint8_t *d_src = nullptr;
    uint16_t *d_g = nullptr;
    int *d_seg = nullptr;
    const int nseg = (height + SEG_ROWS - 1) / SEG_ROWS;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto cleanup = [&](bool all) {
        cudaFree(d_src);
        cudaFree(d_g);
        cudaFree(d_seg);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (all) { cudaFree(m->d_occ); cudaFree(m->d_dist2); cudaFree(m->d_dist); delete m; }
    };
#define RL_TRY(expr)                                                                           \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            cleanup(true);                                                                     \
            return rl::fail(_e == cudaErrorMemoryAllocation ? RL_ERR_OOM : RL_ERR_CUDA,        \
                            std::string(#expr) + ": " + cudaGetErrorString(_e));               \
        }                                                                                      \
    } while (0)
    RL_TRY(cudaMalloc(&d_src, n));
    RL_TRY(cudaMalloc(&d_g, n * sizeof(uint16_t)));
    RL_TRY(cudaMalloc(&d_seg, (size_t)2 * nseg * width * sizeof(int)));
    RL_TRY(cudaMalloc(&m->d_occ, n));
    RL_TRY(cudaMalloc(&m->d_dist2, n * sizeof(int32_t)));
    RL_TRY(cudaMalloc(&m->d_dist, n * sizeof(float)));
    RL_TRY(cudaMemcpy(d_src, src, n, cudaMemcpyHostToDevice));
    RL_TRY(cudaEventCreate(&e0));
    RL_TRY(cudaEventCreate(&e1));
    const size_t smem = (size_t)width * sizeof(uint32_t);
    RL_TRY(cudaFuncSetAttribute(edt_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RL_TRY(cudaEventRecord(e0, 0));
    {
        const dim3 grid((width + 127) / 128, nseg);
        int *seg_first = d_seg, *seg_last = d_seg + (size_t)nseg * width;
        edt_classify_kernel<<<grid, 128>>>(d_src, height, width, flip, lut, m->d_occ, seg_first, seg_last);
        edt_cols_kernel<<<grid, 128>>>(m->d_occ, height, width, nseg, seg_first, seg_last, d_g);
    }
    edt_rows_kernel<<<height, 256, smem>>>(d_g, height, width, m->d_dist2, m->d_dist);
    RL_TRY(cudaGetLastError());
    RL_TRY(cudaEventRecord(e1, 0));
    RL_TRY(cudaEventSynchronize(e1));
    RL_TRY(cudaEventElapsedTime(&m->ingest_ms, e0, e1));
#undef RL_TRY
    cleanup(false);
    *out = m;
    return RL_OK;
}

}  // namespace

void rl_map_retain(const rl_map *m) { const_cast<rl_map *>(m)->refs.fetch_add(1); }

void rl_map_release(const rl_map *cm)
{
    rl_map *m = const_cast<rl_map *>(cm);
    if (m->refs.fetch_sub(1) == 1) {
        rl::DeviceGuard guard(m->device);
        cudaFree(m->d_occ);
        cudaFree(m->d_dist2);
        cudaFree(m->d_dist);
        delete m;
    }
}

extern "C" {

int32_t rl_map_from_image(const uint8_t *pixels, int32_t width, int32_t height, int32_t negate,
                          double occupied_thresh, double free_thresh, int32_t mode, int32_t binarise,
                          double resolution, double origin_x, double origin_y, double origin_yaw,
                          int32_t device, rl_map **out)
{
    // map_server's rule per byte value (trinary / scale / raw), in double like map_server, then the
    // reference's binarisation and PyOMap's cut: all functions of one byte.
    if (mode < RL_MAP_TRINARY || mode > RL_MAP_RAW) return rl::fail(RL_ERR_BAD_ARG, "rl_map_from_image: unknown mode");
    ByteLut lut{};
    for (int p = 0; p < 256; ++p) {
        int v;
        if (mode == RL_MAP_RAW) v = (int8_t)(unsigned char)p;
        else {
            double shade = negate ? p / 255.0 : (255 - p) / 255.0;
            if (shade > occupied_thresh) v = 100;
            else if (shade < free_thresh) v = 0;
            else if (mode == RL_MAP_TRINARY) v = -1;
            else v = (int8_t)(unsigned char)(1 + 98 * ((shade - free_thresh) / (occupied_thresh - free_thresh)));
        }
        if (binarise) v = (v > 0) ? 255 : 0;
        if (v > 10) lut.w[p >> 5] |= 1u << (p & 31);
    }
    return build_map(pixels, width, height, /*flip=*/1, lut, resolution, origin_x, origin_y,
                     origin_yaw, device, out);
}

int32_t rl_map_from_occupancy(const int8_t *data, int32_t width, int32_t height, int32_t binarise,
                              double resolution, double origin_x, double origin_y,
                              double origin_yaw, int32_t device, rl_map **out)
{
    ByteLut lut{};
    for (int p = 0; p < 256; ++p) {
        int v = (int8_t)p;
        if (binarise) v = (v > 0) ? 255 : 0;
        if (v > 10) lut.w[p >> 5] |= 1u << (p & 31);
    }
    return build_map(reinterpret_cast<const uint8_t *>(data), width, height, /*flip=*/0, lut,
                     resolution, origin_x, origin_y, origin_yaw, device, out);
}

int32_t rl_map_from_cells(const uint8_t *occupied, int32_t width, int32_t height,
                          double resolution, double origin_x, double origin_y, double origin_yaw,
                          int32_t device, rl_map **out)
{
    ByteLut lut{};
    for (int p = 1; p < 256; ++p) lut.w[p >> 5] |= 1u << (p & 31);
    return build_map(occupied, width, height, /*flip=*/0, lut, resolution, origin_x, origin_y,
                     origin_yaw, device, out);
}

int32_t rl_map_shape(const rl_map *map, int32_t *width, int32_t *height, int32_t *device)
{
    if (!map) return rl::fail(RL_ERR_BAD_ARG, "rl_map_shape: null map");
    if (width) *width = map->cols;
    if (height) *height = map->rows;
    if (device) *device = map->device;
    return RL_OK;
}

static int32_t copy_out(const rl_map *map, const void *d_src, void *out, size_t elem)
{
    if (!map || !out) return rl::fail(RL_ERR_BAD_ARG, "rl_map_get_*: null pointer");
    rl::DeviceGuard guard(map->device);
    RL_CUDA(cudaMemcpy(out, d_src, (size_t)map->rows * map->cols * elem, cudaMemcpyDeviceToHost));
    return RL_OK;
}

int32_t rl_map_get_occupancy(const rl_map *map, uint8_t *out)
{
    return copy_out(map, map ? map->d_occ : nullptr, out, 1);
}

int32_t rl_map_get_dist2(const rl_map *map, int32_t *out)
{
    return copy_out(map, map ? map->d_dist2 : nullptr, out, sizeof(int32_t));
}

int32_t rl_map_get_dist(const rl_map *map, float *out)
{
    return copy_out(map, map ? map->d_dist : nullptr, out, sizeof(float));
}

int32_t rl_map_dist_device(const rl_map *map, const float **d_dist)
{
    if (!map || !d_dist) return rl::fail(RL_ERR_BAD_ARG, "rl_map_dist_device: null pointer");
    *d_dist = map->d_dist;
    return RL_OK;
}

int32_t rl_map_ingest_ms(const rl_map *map, float *ms)
{
    if (!map || !ms) return rl::fail(RL_ERR_BAD_ARG, "rl_map_ingest_ms: null pointer");
    *ms = map->ingest_ms;
    return RL_OK;
}

int32_t rl_map_destroy(rl_map *map)
{
    if (!map) return rl::fail(RL_ERR_BAD_ARG, "rl_map_destroy: null map");
    rl_map_release(map);
    return RL_OK;
}

}  // extern "C"
