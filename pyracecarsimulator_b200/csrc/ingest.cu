// ingest.cu -- map ingest on the GPU: source cells -> occupancy -> exact integer squared
// Euclidean distance transform -> fp32 distance field.  Replaces the host-side work of
// range_libc's PyOMap + DistanceTransform constructors (reference call sites
// scripts/ros_interface.py:80-86, :210 and scripts/scan_simulator.py:72-76).
//
// Kernels (north_star (a): column pass + row pass):
//   edt_classify_kernel  one thread per (column, 32-row segment): classifies every source cell
//                    through a 256-entry bit LUT (map_server thresholds / binarisation / `> 10`
//                    cut are all pure functions of one byte, so the host folds them into the LUT
//                    with the reference's double arithmetic), applies map_server's y-flip, writes
//                    the occupancy byte, records the segment's first/last occupied row.
//   edt_seg_scan_kernel  one thread per column: nearest occupied row above / below every segment.
//   edt_cols_kernel  one thread per (column, segment): distance g to the nearest occupied cell in the
//                    same column (down sweep seeded from above, up sweep from below), saturating u16.
//   edt_rows_kernel  one CTA per map row with the row's g^2 staged in shared memory: the lower
//                    envelope min_k (k^2 + g^2[q +- k]), by an outward scan that stops at k^2 >= best
//                    when the row's cost bound is small and by divide and conquer over the monotone
//                    argmin (O(W log W) per row whatever the map) otherwise; exact in integers either
//                    way; writes d^2 (int32) and sqrt_rn((float)d^2).  The scan stages the row unpadded (consecutive
//                    threads read consecutive words, the offsets of a trip are immediates of one address) and needs
//                    no clamping while both neighbours are inside the row: 88 -> 64 us on the 2049^2 stand-in.
// For max(rows, cols) <= 2896 the reference's float Felzenszwalb transform returns exactly
// this integer d^2, so the fp32 field is bit-identical to the reference's.
#include <cmath>
#include <cstdlib>
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

constexpr int SEG_ROWS = 32;   // rows per column segment: W * ceil(H/32) threads instead of W
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
#pragma unroll 16
    for (int r = r0; r < r1; ++r) {
        const int sr = flip ? rows - 1 - r : r;
        const uint32_t o = lut_bit(lut, src[(size_t)sr * cols + c]);
        occ[(size_t)r * cols + c] = (uint8_t)o;
        if (o) { first = min(first, r); last = r; }
    }
    seg_first[(size_t)seg * cols + c] = first;
    seg_last[(size_t)seg * cols + c] = last;
}

// The same for colour / alpha images (map_server averages the channels of a pixel before thresholding,
// ROS map_server image_loader.cpp): `channels` bytes per pixel, the first `avg` of them summed; every
// rule downstream is a function of that sum (0 .. 255*avg) and of "last byte == 0" (map_server's alpha
// test, which in scale mode turns an in-between cell into unknown), so two bit LUTs over the sum.
struct SumLut { uint32_t w[2][32]; };       // [last byte == 0][bit = sum]

__global__ void __launch_bounds__(128)
edt_classify_multi_kernel(const uint8_t *__restrict__ src, int rows, int cols, int channels, int avg, int flip,
                          SumLut lut, uint8_t *__restrict__ occ, int *__restrict__ seg_first,
                          int *__restrict__ seg_last)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int seg = blockIdx.y;
    if (c >= cols) return;
    const int r0 = seg * SEG_ROWS, r1 = min(rows, r0 + SEG_ROWS);
    int first = SEG_NONE_FIRST, last = SEG_NONE_LAST;
    for (int r = r0; r < r1; ++r) {
        const int sr = flip ? rows - 1 - r : r;
        const uint8_t *px = src + ((size_t)sr * cols + c) * channels;
        uint32_t sum = 0;
        for (int k = 0; k < avg; ++k) sum += px[k];
        const int a0 = channels > 1 && px[channels - 1] == 0;
        const uint32_t o = (lut.w[a0][sum >> 5] >> (sum & 31u)) & 1u;
        occ[(size_t)r * cols + c] = (uint8_t)o;
        if (o) { first = min(first, r); last = r; }
    }
    seg_first[(size_t)seg * cols + c] = first;
    seg_last[(size_t)seg * cols + c] = last;
}

// Column pass, step 2 (one thread per column, in place): seg_last[seg] becomes the last occupied row ABOVE
// the segment (in any earlier segment), seg_first[seg] the first occupied row BELOW it.  The loads do not
// depend on each other, so the 2 * nseg of them pipeline; searching the neighbouring segments from inside
// the sweep kernel instead cost O(nseg) dependent loads per (column, segment) on sparse maps.
__global__ void __launch_bounds__(128)
edt_seg_scan_kernel(int cols, int nseg, int *__restrict__ seg_first, int *__restrict__ seg_last)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    // batches of 16 loads that do not depend on each other, then the 16 dependent selects + stores
    constexpr int NB = 16;
    int carry = SEG_NONE_LAST;
    for (int s0 = 0; s0 < nseg; s0 += NB) {
        int v[NB];
#pragma unroll
        for (int u = 0; u < NB; ++u) v[u] = (s0 + u < nseg) ? seg_last[(size_t)(s0 + u) * cols + c] : SEG_NONE_LAST;
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            if (s0 + u < nseg) seg_last[(size_t)(s0 + u) * cols + c] = carry;
            if (v[u] != SEG_NONE_LAST) carry = v[u];
        }
    }
    carry = SEG_NONE_FIRST;
    for (int s0 = nseg - 1; s0 >= 0; s0 -= NB) {
        int v[NB];
#pragma unroll
        for (int u = 0; u < NB; ++u) v[u] = (s0 - u >= 0) ? seg_first[(size_t)(s0 - u) * cols + c] : SEG_NONE_FIRST;
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            if (s0 - u >= 0) seg_first[(size_t)(s0 - u) * cols + c] = carry;
            if (v[u] != SEG_NONE_FIRST) carry = v[u];
        }
    }
}

// Column pass, step 3: distance g (saturating u16) to the nearest occupied cell of the same column:
// down sweep seeded with the last occupied row above the segment, up sweep seeded with the first
// occupied row below it.
__global__ void __launch_bounds__(128)
edt_cols_kernel(const uint8_t *__restrict__ occ, int rows, int cols, int nseg,
                const int *__restrict__ seg_below, const int *__restrict__ seg_above,
                uint16_t *__restrict__ g)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int seg = blockIdx.y;
    if (c >= cols) return;
    const int r0 = seg * SEG_ROWS, r1 = min(rows, r0 + SEG_ROWS);
    const int above = seg_above[(size_t)seg * cols + c], below = seg_below[(size_t)seg * cols + c];
    uint32_t run = (above == SEG_NONE_LAST) ? G_INF : min((uint32_t)(r0 - 1 - above), G_INF);
#pragma unroll 16
    for (int r = r0; r < r1; ++r) {
        run = occ[(size_t)r * cols + c] ? 0u : min(run + 1u, G_INF);
        g[(size_t)r * cols + c] = (uint16_t)run;
    }
    run = (below == SEG_NONE_FIRST) ? G_INF : min((uint32_t)(below - r1), G_INF);
#pragma unroll 16
    for (int r = r1 - 1; r >= r0; --r) {
        const uint32_t gv = g[(size_t)r * cols + c];
        run = (gv == 0u) ? 0u : min(run + 1u, G_INF);
        if (run < gv) g[(size_t)r * cols + c] = (uint16_t)run;
    }
}

// Row pass: d2(q) = min_k (q - k)^2 + g2[k] over the row's g^2 staged in shared memory (G2_FAR for columns
// without any occupied cell; G2_FAR + k^2 cannot wrap and stays above every real d^2 <= 2 * 16384^2, so it
// needs no special case).  One CTA per map row, two exact integer algorithms, chosen PER ROW:
//
//   outward scan        each cell looks at k = 1, 2, ... on both sides and stops as soon as k^2 >= best.
//                       O(distance to the nearest obstacle) per cell: nothing on an indoor map, O(W) per cell
//                       on a sparse one (a lone obstacle on an 8192^2 map is ~10^11 probes).  The scan of cell q
//                       never goes beyond k = g[q] (best starts at g[q]^2), so sum_q min(g[q], W) bounds the
//                       row's cost before it is run.
//   divide and conquer  the matrix f(q, k) = (q-k)^2 + g2[k] is Monge (f(q1,k1) + f(q2,k2) <= f(q1,k2) +
//                       f(q2,k1) for q1 < q2, k1 < k2, because -2qk is), so some argmin is monotone in q and
//                       the argmin of a middle column splits the candidates of the columns left and right of
//                       it.  Columns are solved level by level in bisection order -- q+1 = odd multiples of
//                       h = T/2, T/4, ..., 1 -- each searching only [argmin(q - h), argmin(q + h)]; the ranges
//                       of one level sum to at most W + (queries), so a row costs O(W log W) probes whatever
//                       the map.  Ties may resolve to any argmin: the value is what is stored.
//
// A row takes the scan when its bound is at most SCAN_BUDGET probes per cell, the divide and conquer
// otherwise; RL_EDT_ROWS=scan|dc forces one (tests, measurements).  Both give the same exact integers.
// Shared-memory indices are padded by one word per 32 (sw()) so that the power-of-two strides of the
// bisection order do not pile onto one bank.
constexpr uint32_t G2_FAR = 0x3fffffffu;
constexpr int ROW_THREADS = 256;
constexpr int SCAN_BUDGET = 128;   // indoor-style maps (synth 2049^2: 64..128 per cell) stay on the scan
constexpr int ROWS_AUTO = 0, ROWS_SCAN = 1, ROWS_DC = 2;

__device__ __forceinline__ int sw(int i) { return i + (i >> 5); }

__device__ __forceinline__ void dc_bounds(const uint16_t *opt, int cols, int j, int h, int &q, int &lo, int &hi)
{
    const int qp = (2 * j + 1) * h;          // q + 1
    const int left = qp - h, right = qp + h; // solved at an earlier level (or outside the row)
    q = qp - 1;
    lo = left == 0 ? 0 : (int)opt[sw(left - 1)];
    hi = right > cols ? cols - 1 : (int)opt[sw(right - 1)];
}

__global__ void __launch_bounds__(ROW_THREADS)
edt_rows_kernel(const uint16_t *__restrict__ g, int rows, int cols, int log2T, int force, int budget,
                int32_t *__restrict__ dist2, float *__restrict__ dist, float *__restrict__ step)
{
    extern __shared__ uint32_t g2[];                                     // sw(cols) words of g^2 ...
    uint16_t *opt = reinterpret_cast<uint16_t *>(g2 + sw(cols) + 1);     // ... then sw(cols) argmins
    __shared__ uint32_t red_v[ROW_THREADS / 32];
    __shared__ int red_k[ROW_THREADS / 32];
    __shared__ unsigned long long cost;
    const int r = blockIdx.x, tid = threadIdx.x;
    const uint16_t *grow = g + (size_t)r * cols;
    if (tid == 0) cost = 0;
    __syncthreads();
    // first look at the row: is any column within reach of an obstacle, and what would the outward scan cost?
    bool any_near = false;
    uint32_t mine = 0;
    for (int q = tid; q < cols; q += ROW_THREADS) {
        const uint32_t v = grow[q];
        any_near |= v < G_INF;
        mine += min(v, (uint32_t)cols);
    }
    mine = __reduce_add_sync(0xffffffffu, mine);
    if ((tid & 31) == 0) atomicAdd(&cost, (unsigned long long)mine);
    // A row whose every column is G_INF has no occupied cell within reach in any column (an empty
    // map): nothing to minimise.
    if (!__syncthreads_or(any_near)) {
        for (int q = tid; q < cols; q += ROW_THREADS) {
            const size_t o = (size_t)r * cols + q;
            dist2[o] = RL_DIST2_INF;
            dist[o] = sqrtf(1e20f);   // what the reference's INF = 1e20 transform leaves on an empty map
            step[o] = rl::march_step_of(sqrtf(1e20f));
        }
        return;
    }
    const bool scan = force == ROWS_SCAN || (force == ROWS_AUTO && cost <= (unsigned long long)budget * cols);
    if (scan) {
        // The scan reads g2[q - k] and g2[q + k] for consecutive q: the row is staged UNPADDED (g2[q]), consecutive
        // threads read consecutive words and the four offsets of a trip are immediate offsets of one address.  (The
        // second read of the row comes from L1 / L2.)
        for (int q = tid; q < cols; q += ROW_THREADS) {
            const uint32_t v = grow[q];
            g2[q] = (v >= G_INF) ? G2_FAR : v * v;
        }
        __syncthreads();
        const int last = cols - 1;
        for (int q = tid; q < cols; q += ROW_THREADS) {
            uint32_t best = g2[q];
            const int both = min(q, last - q), reach = max(q, last - q);
            int k = 1;
            // both neighbours inside the row: no clamping.  Four offsets per trip (independent loads, one exit test).
            for (; k + 3 <= both; k += 4) {
                if ((uint32_t)k * (uint32_t)k >= best) break;
                const uint32_t *lo = g2 + (q - k), *hi = g2 + (q + k);
                uint32_t c[8];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t k2 = (uint32_t)(k + u) * (uint32_t)(k + u);
                    c[2 * u] = k2 + lo[-u];
                    c[2 * u + 1] = k2 + hi[u];
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) best = min(best, c[u]);
            }
            // the rest (one side has left the row): out-of-row neighbours are clamped to the row ends -- the clamped
            // candidate k^2 + g2[end] is never below the true candidate (q - end)^2 + g2[end], so the minimum is unchanged
            for (; k <= reach; k += 4) {
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
                dist[o] = sqrtf(1e20f);
                step[o] = rl::march_step_of(sqrtf(1e20f));
            } else {
                dist2[o] = (int32_t)best;
                dist[o] = sqrtf((float)best);
                step[o] = rl::march_step_of(sqrtf((float)best));
            }
        }
        return;
    }
    // divide and conquer: the row staged with one padding word per 32 (sw()), so that the power-of-two strides of the
    // bisection order do not pile onto one bank
    for (int q = tid; q < cols; q += ROW_THREADS) {
        const uint32_t v = grow[q];
        g2[sw(q)] = (v >= G_INF) ? G2_FAR : v * v;
    }
    __syncthreads();
    for (int h = 1 << (log2T - 1); h >= 1; h >>= 1) {
        const int nq = (cols / h + 1) >> 1;   // queries of this level: (2j+1) h <= cols
        if (nq > ROW_THREADS / 32) {          // a query per thread (several per thread at the bottom levels)
            for (int base = 0; base < nq; base += ROW_THREADS) {
                // consecutive queries go to different warps, so the few long ranges of a level spread over the CTA
                const int j = base + (tid & 31) * (ROW_THREADS / 32) + (tid >> 5);
                const bool active = j < nq;
                int q = 0, lo = 0, hi = -1;
                if (active) dc_bounds(opt, cols, j, h, q, lo, hi);
                uint32_t bv = 0xffffffffu;
                int bk = lo;
                // The ranges of a level sum to <= W + nq, but one query may own most of that (on a sparse row
                // the leftmost query of every level searches [0, argmin of its right neighbour]): ranges longer
                // than LONG are searched by the whole warp, one after the other, the short ones by their lane.
                constexpr int LONG = 24;
                const bool is_long = active && hi - lo >= LONG;
                if (active && !is_long) {
                    for (int k = lo; k <= hi; ++k) {
                        const int d = q - k;
                        const uint32_t v = (uint32_t)(d * d) + g2[sw(k)];
                        if (v < bv) { bv = v; bk = k; }
                    }
                }
                unsigned todo = __ballot_sync(0xffffffffu, is_long);
                const int lane = tid & 31;
                while (todo) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int qs = __shfl_sync(0xffffffffu, q, src), los = __shfl_sync(0xffffffffu, lo, src),
                              his = __shfl_sync(0xffffffffu, hi, src);
                    uint32_t cv = 0xffffffffu;
                    int ck = los;
                    for (int k = los + lane; k <= his; k += 32) {
                        const int d = qs - k;
                        const uint32_t v = (uint32_t)(d * d) + g2[sw(k)];
                        if (v < cv) { cv = v; ck = k; }
                    }
                    for (int o = 16; o > 0; o >>= 1) {
                        const uint32_t v2 = __shfl_xor_sync(0xffffffffu, cv, o);
                        const int k2 = __shfl_xor_sync(0xffffffffu, ck, o);
                        if (v2 < cv) { cv = v2; ck = k2; }
                    }
                    if (lane == src) { bv = cv; bk = ck; }
                }
                if (active) opt[sw(q)] = (uint16_t)bk;
            }
        } else if (nq > 0) {                  // top levels, at most 8 queries: a group of G >= 32 threads per query
            const int G = 1 << (31 - __clz(ROW_THREADS / nq));
            const int grp = tid / G, l = tid & (G - 1);
            const bool active = grp < nq;
            uint32_t bv = 0xffffffffu;
            int bk = 0, q = 0;
            if (active) {
                int lo, hi;
                dc_bounds(opt, cols, grp, h, q, lo, hi);
                bk = lo;
                for (int k = lo + l; k <= hi; k += G) {
                    const int d = q - k;
                    const uint32_t v = (uint32_t)(d * d) + g2[sw(k)];
                    if (v < bv) { bv = v; bk = k; }
                }
            }
            const int W = G < 32 ? G : 32;
            for (int o = W >> 1; o > 0; o >>= 1) {
                const uint32_t v2 = __shfl_xor_sync(0xffffffffu, bv, o);
                const int k2 = __shfl_xor_sync(0xffffffffu, bk, o);
                if (v2 < bv) { bv = v2; bk = k2; }
            }
            if (G <= 32) {
                if (active && l == 0) opt[sw(q)] = (uint16_t)bk;
            } else {
                if ((tid & 31) == 0) { red_v[tid >> 5] = bv; red_k[tid >> 5] = bk; }
                __syncthreads();
                if (active && l == 0) {
                    const int w0 = tid >> 5, nw = G >> 5;
                    for (int w = w0 + 1; w < w0 + nw; ++w)
                        if (red_v[w] < bv) { bv = red_v[w]; bk = red_k[w]; }
                    opt[sw(q)] = (uint16_t)bk;
                }
            }
        }
        __syncthreads();
    }
    for (int q = tid; q < cols; q += ROW_THREADS) {
        const int k = opt[sw(q)], d = q - k;
        const uint32_t best = (uint32_t)(d * d) + g2[sw(k)];
        const size_t o = (size_t)r * cols + q;
        if (best >= G2_FAR) {
            dist2[o] = RL_DIST2_INF;
            dist[o] = sqrtf(1e20f);
            step[o] = rl::march_step_of(sqrtf(1e20f));
        } else {
            dist2[o] = (int32_t)best;
            dist[o] = sqrtf((float)best);
            step[o] = rl::march_step_of(sqrtf((float)best));
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
                  double resolution, double ox, double oy, double yaw, int device, rl_map **out,
                  int channels = 1, int avg = 1, const SumLut *sum_lut = nullptr)
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
        if (all) { cudaFree(m->d_occ); cudaFree(m->d_dist2); cudaFree(m->d_dist); cudaFree(m->d_step); delete m; }
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
    RL_TRY(cudaMalloc(&d_src, n * channels));
    RL_TRY(cudaMalloc(&d_g, n * sizeof(uint16_t)));
    RL_TRY(cudaMalloc(&d_seg, (size_t)2 * nseg * width * sizeof(int)));
    RL_TRY(cudaMalloc(&m->d_occ, n));
    RL_TRY(cudaMalloc(&m->d_dist2, n * sizeof(int32_t)));
    RL_TRY(cudaMalloc(&m->d_dist, n * sizeof(float)));
    RL_TRY(cudaMalloc(&m->d_step, n * sizeof(float)));
    RL_TRY(cudaMemcpy(d_src, src, n * channels, cudaMemcpyHostToDevice));
    RL_TRY(cudaEventCreate(&e0));
    RL_TRY(cudaEventCreate(&e1));
    const size_t padded = (size_t)width + ((size_t)width >> 5) + 2;
    const size_t smem = padded * (sizeof(uint32_t) + sizeof(uint16_t));
    // RL_EDT_ROWS=scan|dc forces one row algorithm (tests and measurements); default: chosen per row
    const char *rows_env = std::getenv("RL_EDT_ROWS");
    const int force = !rows_env ? ROWS_AUTO : (rows_env[0] == 's' ? ROWS_SCAN : (rows_env[0] == 'd' ? ROWS_DC : ROWS_AUTO));
    int budget = SCAN_BUDGET;
    if (const char *e = std::getenv("RL_EDT_BUDGET")) { const long v = std::atol(e); if (v >= 0 && v <= 1 << 20) budget = (int)v; }
    RL_TRY(cudaFuncSetAttribute(edt_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RL_TRY(cudaEventRecord(e0, 0));
    {
        const dim3 grid((width + 127) / 128, nseg);
        int *seg_first = d_seg, *seg_last = d_seg + (size_t)nseg * width;
        if (sum_lut) edt_classify_multi_kernel<<<grid, 128>>>(d_src, height, width, channels, avg, flip, *sum_lut, m->d_occ, seg_first, seg_last);
        else edt_classify_kernel<<<grid, 128>>>(d_src, height, width, flip, lut, m->d_occ, seg_first, seg_last);
        edt_seg_scan_kernel<<<(width + 127) / 128, 128>>>(width, nseg, seg_first, seg_last);
        edt_cols_kernel<<<grid, 128>>>(m->d_occ, height, width, nseg, seg_first, seg_last, d_g);
    }
    {
        int log2T = 1;
        while ((1 << log2T) < width + 1) ++log2T;
        edt_rows_kernel<<<height, ROW_THREADS, smem>>>(d_g, height, width, log2T, force, budget, m->d_dist2, m->d_dist, m->d_step);
    }
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
        cudaFree(m->d_step);
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
        if (mode == RL_MAP_RAW) v = (int8_t)(unsigned char)(negate ? 255 - p : p);   // map_server negates before the raw copy
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

int32_t rl_map_from_image_channels(const uint8_t *pixels, int32_t width, int32_t height, int32_t channels,
                                   int32_t has_alpha, int32_t negate, double occupied_thresh, double free_thresh,
                                   int32_t mode, int32_t binarise, double resolution, double origin_x,
                                   double origin_y, double origin_yaw, int32_t device, rl_map **out)
{
    if (mode < RL_MAP_TRINARY || mode > RL_MAP_RAW) return rl::fail(RL_ERR_BAD_ARG, "rl_map_from_image_channels: unknown mode");
    if (channels < 1 || channels > 4) return rl::fail(RL_ERR_BAD_ARG, "rl_map_from_image_channels: 1 to 4 channels");
    // map_server: trinary mode averages every channel (alpha included), the other modes leave a real alpha out
    const int avg = (mode == RL_MAP_TRINARY || !has_alpha) ? channels : channels - 1;
    SumLut lut{};
    for (int a0 = 0; a0 < 2; ++a0) {
        for (int sum = 0; sum <= 255 * avg; ++sum) {
            double color_avg = sum / (double)avg;
            if (negate) color_avg = 255 - color_avg;
            int v;
            if (mode == RL_MAP_RAW) v = (int8_t)(unsigned char)color_avg;
            else {
                const double shade = (255 - color_avg) / 255.0;
                if (shade > occupied_thresh) v = 100;
                else if (shade < free_thresh) v = 0;
                else if (mode == RL_MAP_TRINARY || (channels > 1 && a0)) v = -1;
                else v = (int8_t)(unsigned char)(1 + 98 * ((shade - free_thresh) / (occupied_thresh - free_thresh)));
            }
            if (binarise) v = (v > 0) ? 255 : 0;
            if (v > 10) lut.w[a0][sum >> 5] |= 1u << (sum & 31);
        }
    }
    return build_map(pixels, width, height, /*flip=*/1, ByteLut{}, resolution, origin_x, origin_y, origin_yaw,
                     device, out, channels, avg, &lut);
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
