// tools/tune_march.cu -- development harness (not shipped, not on the product path): times
// candidate fan-march kernel structures on the same distance field and poses, checks every
// candidate bit-for-bit against the straightforward one, and prints a table.
//   python tools/tune_prep.py /tmp/tune && tools/tune_march /tmp/tune [reps]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../pyracecarsimulator_b200/csrc/glibc_trig.cuh"
#include "../pyracecarsimulator_b200/csrc/march.cuh"

using rl::GridPose;
using rl::MarchParams;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

struct Args {
    MarchParams P;
    const float *poses;
    float *outs;
    int num_poses, num_beams;
    float fov;
    unsigned int *counter;   // work queue
};

__device__ __forceinline__ float beam_heading(const Args &a, float thw, int j)
{
    const float inc = a.fov / (float)a.num_beams;
    return __fadd_rn(-__fadd_rn(thw, fmaf((float)j, inc, -0.5f * a.fov)), a.P.w.rotation_const);
}

// ---------------------------------------------------------------- V0: product structure
template <int WPC>
__global__ void __launch_bounds__(WPC * 32) k_base(Args a, int segs, int seg_len)
{
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * WPC + (threadIdx.x >> 5);
    const int k = warp / segs;
    if (k >= a.num_poses) return;
    const int seg = warp - k * segs;
    const float *p = a.poses + 3 * k;
    const GridPose g = rl::world_to_grid(a.P.w, __ldg(p), __ldg(p + 1), __ldg(p + 2));
    const float thw = __ldg(p + 2);
    const int j_end = min(a.num_beams, (seg + 1) * seg_len);
    float *o = a.outs + (size_t)k * a.num_beams;
    uint32_t st = 0;
    for (int j = seg * seg_len + lane; j < j_end; j += 32) {
        const rl::FirstSample f0 = rl::first_sample(a.P, g.y, g.x);
        float s, c;
        rl::glibc_sincosf(beam_heading(a, thw, j), &s, &c);
        o[j] = __fmul_rn(rl::march_ray<false>(a.P, g.y, g.x, c, s, st, f0), a.P.w.scale);
    }
}

// ---------------------------------------------------------------- V1: persistent warps, global queue of
// (pose, chunk) tasks, ILP independent rays per lane marched in lockstep
template <int ILP>
__device__ __forceinline__ void march_ilp(const MarchParams &P, float x0, float y0, const float (&dx)[ILP],
                                          const float (&dy)[ILP], bool (&live)[ILP], float (&res)[ILP])
{
    float t[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { t[i] = 0.f; res[i] = P.max_range; }
    bool any = false;
#pragma unroll
    for (int i = 0; i < ILP; ++i) any |= live[i];
    while (any) {
        float d[ILP];
        int px[ILP], py[ILP];
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            d[i] = 1.0f;
            if (live[i]) {
                const float fx = fmaf(dx[i], t[i], x0);
                const float fy = fmaf(dy[i], t[i], y0);
                if (!(fx > -1.0f && fx < P.frows && fy > -1.0f && fy < P.fcols)) { live[i] = false; }
                else {
                    px[i] = __float2int_rz(fx);
                    py[i] = __float2int_rz(fy);
                    d[i] = __ldg(P.dist + (px[i] * P.cols + py[i]));
                }
            }
        }
        any = false;
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (live[i]) {
                if (d[i] <= 0.0f) {
                    const float xd = __fsub_rn((float)px[i], x0);
                    const float yd = __fsub_rn((float)py[i], y0);
                    res[i] = sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
                    live[i] = false;
                } else {
                    t[i] = __fadd_rn(t[i], fmaxf(__fmul_rn(d[i], 0.999f), 1.0f));
                    if (!(t[i] < P.max_range)) live[i] = false;
                }
            }
            any |= live[i];
        }
    }
}

template <int ILP, int WPC>
__global__ void __launch_bounds__(WPC * 32) k_persist(Args a, int chunks_per_pose)
{
    const int lane = threadIdx.x & 31;
    const unsigned total = (unsigned)a.num_poses * chunks_per_pose;
    unsigned task = 0;
    if (lane == 0) task = atomicAdd(a.counter, 1u);
    task = __shfl_sync(0xffffffffu, task, 0);
    while (task < total) {
        unsigned next = 0;
        if (lane == 0) next = atomicAdd(a.counter, 1u);   // prefetch the next task id
        const int k = task / chunks_per_pose;
        const int chunk = task - k * chunks_per_pose;
        const float *p = a.poses + 3 * k;
        const GridPose g = rl::world_to_grid(a.P.w, __ldg(p), __ldg(p + 1), __ldg(p + 2));
        const float thw = __ldg(p + 2);
        float dx[ILP], dy[ILP], res[ILP];
        bool live[ILP];
        const int j0 = chunk * 32 * ILP + lane;
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            const int j = j0 + 32 * i;
            live[i] = j < a.num_beams;
            rl::glibc_sincosf(beam_heading(a, thw, j), &dy[i], &dx[i]);
        }
        march_ilp<ILP>(a.P, g.y, g.x, dx, dy, live, res);
        float *o = a.outs + (size_t)k * a.num_beams;
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            const int j = j0 + 32 * i;
            if (j < a.num_beams) o[j] = __fmul_rn(res[i], a.P.w.scale);
        }
        task = __shfl_sync(0xffffffffu, next, 0);
    }
}

// ---------------------------------------------------------------- V2: persistent warps, NB beams per lane with
// directions precomputed in registers, per-lane "flattened" loop (a lane starts its next beam as
// soon as its current one ends, without waiting for the other lanes)
template <int NB, int WPC>
__global__ void __launch_bounds__(WPC * 32) k_flat(Args a, int chunks_per_pose)
{
    const int lane = threadIdx.x & 31;
    const unsigned total = (unsigned)a.num_poses * chunks_per_pose;
    const MarchParams &P = a.P;
    unsigned task = 0;
    if (lane == 0) task = atomicAdd(a.counter, 1u);
    task = __shfl_sync(0xffffffffu, task, 0);
    while (task < total) {
        unsigned next = 0;
        if (lane == 0) next = atomicAdd(a.counter, 1u);
        const int k = task / chunks_per_pose;
        const int chunk = task - k * chunks_per_pose;
        const float *p = a.poses + 3 * k;
        const GridPose g = rl::world_to_grid(P.w, __ldg(p), __ldg(p + 1), __ldg(p + 2));
        const float thw = __ldg(p + 2);
        const float x0 = g.y, y0 = g.x;
        float dxs[NB], dys[NB];
        const int j0 = chunk * 32 * NB + lane;
#pragma unroll
        for (int i = 0; i < NB; ++i) rl::glibc_sincosf(beam_heading(a, thw, j0 + 32 * i), &dys[i], &dxs[i]);
        float *o = a.outs + (size_t)k * a.num_beams;
        int b = 0;
        int nb = 0;   // beams this lane owns in this chunk
#pragma unroll
        for (int i = 0; i < NB; ++i) nb += (j0 + 32 * i < a.num_beams) ? 1 : 0;
        float dx = dxs[0], dy = dys[0], t = 0.f;
        while (b < nb) {
            const float fx = fmaf(dx, t, x0);
            const float fy = fmaf(dy, t, y0);
            float r = P.max_range;
            bool done = false;
            if (!(fx > -1.0f && fx < P.frows && fy > -1.0f && fy < P.fcols)) done = true;
            else {
                const int px = __float2int_rz(fx), py = __float2int_rz(fy);
                const float d = __ldg(P.dist + (px * P.cols + py));
                if (d <= 0.0f) {
                    const float xd = __fsub_rn((float)px, x0), yd = __fsub_rn((float)py, y0);
                    r = sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
                    done = true;
                } else {
                    t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
                    done = !(t < P.max_range);
                }
            }
            if (done) {
                o[j0 + 32 * b] = __fmul_rn(r, P.w.scale);
                ++b;
                t = 0.f;
#pragma unroll
                for (int i = 1; i < NB; ++i)
                    if (b == i) { dx = dxs[i]; dy = dys[i]; }
            }
        }
        task = __shfl_sync(0xffffffffu, next, 0);
    }
}

// ---------------------------------------------------------------- V3: one CTA per pose (all its rays on one SM, so
// the near field of the pose is fetched into that SM's L1 once); warps pull 32*ILP-beam chunks
// from a shared-memory counter; CTAs pull poses from the global counter
template <int ILP, int WPC>
__global__ void __launch_bounds__(WPC * 32) k_cta_pose(Args a, int chunks_per_pose)
{
    __shared__ int s_pose, s_chunk;
    const int lane = threadIdx.x & 31;
    for (;;) {
        if (threadIdx.x == 0) { s_pose = (int)atomicAdd(a.counter, 1u); s_chunk = 0; }
        __syncthreads();
        const int k = s_pose;
        if (k >= a.num_poses) break;
        const float *p = a.poses + 3 * k;
        const GridPose g = rl::world_to_grid(a.P.w, __ldg(p), __ldg(p + 1), __ldg(p + 2));
        const float thw = __ldg(p + 2);
        float *o = a.outs + (size_t)k * a.num_beams;
        for (;;) {
            int chunk = 0;
            if (lane == 0) chunk = atomicAdd(&s_chunk, 1);
            chunk = __shfl_sync(0xffffffffu, chunk, 0);
            if (chunk >= chunks_per_pose) break;
            float dx[ILP], dy[ILP], res[ILP];
            bool live[ILP];
            const int j0 = chunk * 32 * ILP + lane;
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                const int j = j0 + 32 * i;
                live[i] = j < a.num_beams;
                rl::glibc_sincosf(beam_heading(a, thw, j), &dy[i], &dx[i]);
            }
            march_ilp<ILP>(a.P, g.y, g.x, dx, dy, live, res);
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                const int j = j0 + 32 * i;
                if (j < a.num_beams) o[j] = __fmul_rn(res[i], a.P.w.scale);
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- V4: flat thread-per-ray with alternative
// distance-field layouts (the march is bound by L2 sector requests: make sectors 2-D)
enum Layout { ROWMAJOR = 0, TILE_2x4 = 1, TILE_LINE = 2, TEX = 3, U16_4x4 = 4, TILE_4x2 = 5 };

struct LayoutArgs {
    const float *tiled;      // re-laid-out copy
    const unsigned short *d2u16;
    cudaTextureObject_t tex;
    int tiles_per_row;       // in tiles along cols
};

template <int L>
__device__ __forceinline__ float fetch(const MarchParams &P, const LayoutArgs &la, int px, int py)
{
    if (L == ROWMAJOR) return __ldg(P.dist + (px * P.cols + py));
    if (L == TILE_2x4) {   // 32 B sector = 2 rows x 4 cols
        const int idx = (((px >> 1) * la.tiles_per_row + (py >> 2)) << 3) | ((px & 1) << 2) | (py & 3);
        return __ldg(la.tiled + idx);
    }
    if (L == TILE_4x2) {   // 32 B sector = 4 rows x 2 cols
        const int idx = (((px >> 2) * la.tiles_per_row + (py >> 1)) << 3) | ((px & 3) << 1) | (py & 1);
        return __ldg(la.tiled + idx);
    }
    if (L == TILE_LINE) {  // 128 B line = 4 rows x 8 cols made of four 2x4 sectors
        const int line = (px >> 2) * la.tiles_per_row + (py >> 3);
        const int sector = ((px >> 1) & 1) * 2 + ((py >> 2) & 1);
        return __ldg(la.tiled + ((line << 5) | (sector << 3) | ((px & 1) << 2) | (py & 3)));
    }
    if (L == TEX) return tex2D<float>(la.tex, (float)py + 0.5f, (float)px + 0.5f);
    if (L == U16_4x4) {    // exact d^2 as u16 (valid when max d^2 < 65536), 32 B sector = 4x4 cells
        const int idx = (((px >> 2) * la.tiles_per_row + (py >> 2)) << 4) | ((px & 3) << 2) | (py & 3);
        return sqrtf((float)__ldg(la.d2u16 + idx));
    }
    return 0.f;
}

template <int L, int BS = 256>
__global__ void __launch_bounds__(BS) k_ray(Args a, LayoutArgs la, int cap = 1 << 30)
{
    const unsigned i = blockIdx.x * (unsigned)BS + threadIdx.x;
    const unsigned total = (unsigned)a.num_poses * a.num_beams;
    if (i >= total) return;
    const MarchParams &P = a.P;
    const unsigned k = i / (unsigned)a.num_beams;
    const int j = i - k * a.num_beams;
    const float *p = a.poses + 3 * k;
    const float thw = __ldg(p + 2);
    const GridPose g = rl::world_to_grid(P.w, __ldg(p), __ldg(p + 1), thw);
    float dx, dy;
    rl::glibc_sincosf(beam_heading(a, thw, j), &dy, &dx);
    const float x0 = g.y, y0 = g.x;
    float t = 0.f, r = P.max_range;
    int it = 0;
    while (t < P.max_range) {
        if (++it > cap) break;
        const float fx = fmaf(dx, t, x0), fy = fmaf(dy, t, y0);
        if (!(fx > -1.0f && fx < P.frows && fy > -1.0f && fy < P.fcols)) break;
        const int px = __float2int_rz(fx), py = __float2int_rz(fy);
        const float d = fetch<L>(P, la, px, py);
        if (d <= 0.0f) {
            const float xd = __fsub_rn((float)px, x0), yd = __fsub_rn((float)py, y0);
            r = sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
            break;
        }
        t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
    }
    a.outs[i] = __fmul_rn(r, P.w.scale);
}

// ---------------------------------------------------------------- V5: leaner thread-per-ray: magic-number
// divide, host-computed beam increment, integer bounds test, hit distance computed once after the
// loop has reconverged (not once per divergent exit group)
struct Lean { uint32_t magic; int shift; float inc; };

template <int BS, int RPL>
__global__ void __launch_bounds__(BS) k_lean(Args a, Lean q, int cap)
{
    const MarchParams &P = a.P;
    const unsigned total = (unsigned)a.num_poses * a.num_beams;
    // RPL rays per lane: ray r of this lane is i0 + r*32 (a warp covers 32*RPL consecutive rays)
    const unsigned warp_base = (blockIdx.x * (unsigned)BS + (threadIdx.x & ~31u)) * RPL + (threadIdx.x & 31u);
    float dx[RPL], dy[RPL], x0[RPL], y0[RPL];
    bool valid[RPL];
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const unsigned i = warp_base + 32u * r;
        valid[r] = i < total;
        const unsigned ii = valid[r] ? i : 0u;
        const unsigned k = __umulhi(ii, q.magic) >> q.shift;
        const int j = ii - k * a.num_beams;
        const float *p = a.poses + 3 * k;
        const float thw = __ldg(p + 2);
        const GridPose g = rl::world_to_grid(P.w, __ldg(p), __ldg(p + 1), thw);
        const float thg = __fadd_rn(-__fadd_rn(thw, fmaf((float)j, q.inc, -0.5f * a.fov)), P.w.rotation_const);
        rl::glibc_sincosf(thg, &dy[r], &dx[r]);
        x0[r] = g.y; y0[r] = g.x;
    }
    int hx[RPL], hy[RPL];   // hit cell, hx < 0: no hit (max range)
    int cur = 0;
    while (cur < RPL && !valid[cur]) ++cur;
    float cdx = dx[0], cdy = dy[0], cx0 = x0[0], cy0 = y0[0], t = 0.f;
#pragma unroll
    for (int r = 1; r < RPL; ++r) if (cur == r) { cdx = dx[r]; cdy = dy[r]; cx0 = x0[r]; cy0 = y0[r]; }
#pragma unroll
    for (int r = 0; r < RPL; ++r) hx[r] = -1;
    int it = 0;
    while (cur < RPL) {
        bool done = false;
        int px = -1, py = 0;
        if (!(cdx == cdx) || !(cx0 == cx0) || !(cy0 == cy0) || ++it > cap) done = true;   // NaN pose/heading: leaves the map
        else {
            px = __float2int_rz(fmaf(cdx, t, cx0));
            py = __float2int_rz(fmaf(cdy, t, cy0));
            if ((unsigned)px >= (unsigned)P.rows || (unsigned)py >= (unsigned)P.cols) { done = true; px = -1; }
            else {
                const float d = __ldg(P.dist + (px * P.cols + py));
                if (d <= 0.0f) done = true;
                else {
                    t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
                    if (!(t < P.max_range)) { done = true; px = -1; }
                }
            }
        }
        if (done) {
#pragma unroll
            for (int r = 0; r < RPL; ++r) if (cur == r) { hx[r] = px; hy[r] = py; }
            ++cur;
            t = 0.f; it = 0;
#pragma unroll
            for (int r = 1; r < RPL; ++r) if (cur == r) { cdx = dx[r]; cdy = dy[r]; cx0 = x0[r]; cy0 = y0[r]; if (!valid[r]) cur = RPL; }
        }
    }
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const unsigned i = warp_base + 32u * r;
        float res = P.max_range;
        if (hx[r] >= 0) {
            const float xd = __fsub_rn((float)hx[r], x0[r]), yd = __fsub_rn((float)hy[r], y0[r]);
            res = sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
        }
        if (valid[r]) a.outs[i] = __fmul_rn(res, P.w.scale);
    }
}

// ---------------------------------------------------------------- V6: k_ray with minimal changes, one at a time
// OPT bit0: magic divide + host inc; bit1: hit distance after the loop; bit2: integer bounds test
template <int OPT>
__global__ void __launch_bounds__(128) k_ray3(Args a, Lean q)
{
    const unsigned i = blockIdx.x * 128u + threadIdx.x;
    const unsigned total = (unsigned)a.num_poses * a.num_beams;
    if (i >= total) return;
    const MarchParams &P = a.P;
    unsigned k;
    if (OPT & 1) k = __umulhi(i, q.magic) >> q.shift; else k = i / (unsigned)a.num_beams;
    const int j = i - k * a.num_beams;
    const float *p = a.poses + 3 * k;
    const float thw = __ldg(p + 2);
    const GridPose g = rl::world_to_grid(P.w, __ldg(p), __ldg(p + 1), thw);
    const float inc = (OPT & 1) ? q.inc : a.fov / (float)a.num_beams;
    const float thg = __fadd_rn(-__fadd_rn(thw, fmaf((float)j, inc, -0.5f * a.fov)), P.w.rotation_const);
    float dx, dy;
    rl::glibc_sincosf(thg, &dy, &dx);
    const float x0 = g.y, y0 = g.x;
    float t = 0.f, r = P.max_range;
    int hx = -1, hy = 0;
    const bool bad = !(x0 == x0) || !(y0 == y0) || !(dx == dx);
    if (!bad) {
        while (t < P.max_range) {
            const float fx = fmaf(dx, t, x0), fy = fmaf(dy, t, y0);
            int px, py;
            if (OPT & 4) {
                px = __float2int_rz(fx); py = __float2int_rz(fy);
                if ((unsigned)px >= (unsigned)P.rows || (unsigned)py >= (unsigned)P.cols) break;
            } else {
                if (!(fx > -1.0f && fx < P.frows && fy > -1.0f && fy < P.fcols)) break;
                px = __float2int_rz(fx); py = __float2int_rz(fy);
            }
            const float d = __ldg(P.dist + (px * P.cols + py));
            if (d <= 0.0f) {
                if (OPT & 2) { hx = px; hy = py; }
                else {
                    const float xd = __fsub_rn((float)px, x0), yd = __fsub_rn((float)py, y0);
                    r = sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
                }
                break;
            }
            t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
        }
    }
    if ((OPT & 2) && hx >= 0) {
        const float xd = __fsub_rn((float)hx, x0), yd = __fsub_rn((float)hy, y0);
        r = sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
    }
    a.outs[i] = __fmul_rn(r, P.w.scale);
}

// ---------------------------------------------------------------- V7: product loop + look-ahead touch when the
// clearance is small (creeping rays pay one L2 round trip per 1-px step otherwise)
template <int AHEAD, int MODE>
__global__ void __launch_bounds__(128) k_touch(Args a, Lean q, float near)
{
    const unsigned i = blockIdx.x * 128u + threadIdx.x;
    const unsigned total = (unsigned)a.num_poses * a.num_beams;
    if (i >= total) return;
    const MarchParams &P = a.P;
    const unsigned k = __umulhi(i, q.magic) >> q.shift;
    const int j = i - k * a.num_beams;
    const float *p = a.poses + 3 * k;
    const float thw = __ldg(p + 2);
    const GridPose g = rl::world_to_grid(P.w, __ldg(p), __ldg(p + 1), thw);
    const float thg = __fadd_rn(-__fadd_rn(thw, fmaf((float)j, q.inc, -0.5f * a.fov)), P.w.rotation_const);
    float dx, dy;
    rl::glibc_sincosf(thg, &dy, &dx);
    const float x0 = g.y, y0 = g.x;
    float t = 0.f, r = P.max_range;
    if ((x0 == x0) && (y0 == y0) && (dx == dx)) {
        while (t < P.max_range) {
            const int px = __float2int_rz(fmaf(dx, t, x0)), py = __float2int_rz(fmaf(dy, t, y0));
            if ((unsigned)px >= (unsigned)P.rows || (unsigned)py >= (unsigned)P.cols) break;
            const float *cell = P.dist + (px * P.cols + py);
            const float d = __ldg(cell);
            if (d <= 0.0f) {
                const float xd = __fsub_rn((float)px, x0), yd = __fsub_rn((float)py, y0);
                r = sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
                break;
            }
            if (d < near) {   // touch the sector AHEAD px further along the ray
                const float ta = t + (float)AHEAD;
                const int ax = __float2int_rz(fmaf(dx, ta, x0)), ay = __float2int_rz(fmaf(dy, ta, y0));
                if ((unsigned)ax < (unsigned)P.rows && (unsigned)ay < (unsigned)P.cols) {
                    const float *pa = P.dist + (ax * P.cols + ay);
                    if (MODE == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(pa));
                    else { float junk; asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(junk) : "l"(pa)); }
                }
            }
            t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
        }
    }
    a.outs[i] = __fmul_rn(r, P.w.scale);
}

// ---------------------------------------------------------------- V8: two-phase march.  Phase 1 marches every
// ray for at most `cap` steps; a ray still alive then is appended (index, t) to a queue and phase 2
// marches the queue compacted, 32 long rays per warp, instead of leaving one live lane per warp.
struct TailQ { unsigned *count; uint2 *entries; unsigned capacity; };

__device__ __forceinline__ void ray_setup_fan(const Args &a, const Lean &q, unsigned i, float &x0, float &y0,
                                              float &dx, float &dy)
{
    const MarchParams &P = a.P;
    const unsigned k = __umulhi(i, q.magic) >> q.shift;
    const int j = i - k * a.num_beams;
    const float *p = a.poses + 3 * k;
    const float thw = __ldg(p + 2);
    const GridPose g = rl::world_to_grid(P.w, __ldg(p), __ldg(p + 1), thw);
    const float thg = __fadd_rn(-__fadd_rn(thw, fmaf((float)j, q.inc, -0.5f * a.fov)), P.w.rotation_const);
    rl::glibc_sincosf(thg, &dy, &dx);
    x0 = g.y; y0 = g.x;
}

// marches from t; returns true when finished (r valid), false when `budget` steps ran out (t updated)
__device__ __forceinline__ bool march_some(const MarchParams &P, float x0, float y0, float dx, float dy, float &t,
                                           int budget, float &r)
{
    r = P.max_range;
    if (!((x0 == x0) && (y0 == y0) && (dx == dx))) return true;
    while (t < P.max_range) {
        if (budget-- == 0) return false;
        const int px = __float2int_rz(fmaf(dx, t, x0)), py = __float2int_rz(fmaf(dy, t, y0));
        if ((unsigned)px >= (unsigned)P.rows || (unsigned)py >= (unsigned)P.cols) return true;
        const float d = __ldg(P.dist + (px * P.cols + py));
        if (d <= 0.0f) {
            const float xd = __fsub_rn((float)px, x0), yd = __fsub_rn((float)py, y0);
            r = sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
            return true;
        }
        t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
    }
    return true;
}

__global__ void __launch_bounds__(128) k_phase1(Args a, Lean q, TailQ Q, int cap)
{
    const unsigned i = blockIdx.x * 128u + threadIdx.x;
    const unsigned total = (unsigned)a.num_poses * a.num_beams;
    if (i >= total) return;
    float x0, y0, dx, dy, t = 0.f, r;
    ray_setup_fan(a, q, i, x0, y0, dx, dy);
    bool done = march_some(a.P, x0, y0, dx, dy, t, cap, r);
    if (!done) {
        // warp-aggregated append
        const unsigned mask = __activemask();
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(mask) - 1;
        unsigned base = 0;
        if (lane == leader) base = atomicAdd(Q.count, __popc(mask));
        base = __shfl_sync(mask, base, leader);
        const unsigned slot = base + __popc(mask & ((1u << lane) - 1));
        if (slot < Q.capacity) { Q.entries[slot] = make_uint2(i, __float_as_uint(t)); return; }
        done = march_some(a.P, x0, y0, dx, dy, t, 1 << 30, r);   // queue full: finish inline
    }
    a.outs[i] = __fmul_rn(r, a.P.w.scale);
}

__global__ void __launch_bounds__(128) k_phase2(Args a, Lean q, TailQ Q)
{
    const unsigned n = min(*Q.count, Q.capacity);
    for (unsigned e = blockIdx.x * 128u + threadIdx.x; e < n; e += gridDim.x * 128u) {
        const uint2 ent = Q.entries[e];
        float x0, y0, dx, dy, t = __uint_as_float(ent.y), r;
        ray_setup_fan(a, q, ent.x, x0, y0, dx, dy);
        march_some(a.P, x0, y0, dx, dy, t, 1 << 30, r);
        a.outs[ent.x] = __fmul_rn(r, a.P.w.scale);
    }
}

// ---------------------------------------------------------------- V9: two copies of the distance field, row-major
// and column-major; every ray reads the copy whose fast axis is its dominant direction, so that the
// small steps it takes near walls stay inside one 32-byte sector (L1 hits) whichever way it travels.
template <int MODE>   // 0: choose by |dx| vs |dy|; 1: always transposed (control)
__global__ void __launch_bounds__(128) k_dual(Args a, Lean q, const float *__restrict__ distT)
{
    const unsigned i = blockIdx.x * 128u + threadIdx.x;
    const unsigned total = (unsigned)a.num_poses * a.num_beams;
    if (i >= total) return;
    const MarchParams &P = a.P;
    float x0, y0, dx, dy;
    ray_setup_fan(a, q, i, x0, y0, dx, dy);
    // x runs over rows (first index).  Row-major: idx = px*cols + py (fast axis = y).
    const bool useT = MODE == 1 ? true : (fabsf(dx) > fabsf(dy));   // moving mostly along x: use the copy whose fast axis is x
    const float *base = useT ? distT : P.dist;
    const int sx = useT ? 1 : P.cols, sy = useT ? P.rows : 1;
    float t = 0.f, r = P.max_range;
    if ((x0 == x0) && (y0 == y0) && (dx == dx)) {
        while (t < P.max_range) {
            const int px = __float2int_rz(fmaf(dx, t, x0)), py = __float2int_rz(fmaf(dy, t, y0));
            if ((unsigned)px >= (unsigned)P.rows || (unsigned)py >= (unsigned)P.cols) break;
            const float d = __ldg(base + (px * sx + py * sy));
            if (d <= 0.0f) {
                const float xd = __fsub_rn((float)px, x0), yd = __fsub_rn((float)py, y0);
                r = sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
                break;
            }
            t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
        }
    }
    a.outs[i] = __fmul_rn(r, P.w.scale);
}

// ---------------------------------------------------------------- diagnostics: per-ray loop cycles and a timeline
struct Diag { unsigned long long *t_first, *t_last; unsigned *hist_end_us; unsigned long long *long_cycles; unsigned *long_steps; unsigned *n_long;
              unsigned long long *t0; };

__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

template <int DUAL>
__global__ void __launch_bounds__(128) k_diag(Args a, Lean q, Diag D, const float *__restrict__ distT)
{
    const unsigned i = blockIdx.x * 128u + threadIdx.x;
    const unsigned total = (unsigned)a.num_poses * a.num_beams;
    if (i >= total) return;
    const MarchParams &P = a.P;
    const unsigned long long tstart = gtime();
    float x0, y0, dx, dy;
    ray_setup_fan(a, q, i, x0, y0, dx, dy);
    float t = 0.f, r = P.max_range;
    unsigned steps = 0;
    const bool useT = DUAL && (fabsf(dx) > fabsf(dy));
    const float *base = useT ? distT : P.dist;
    const int sx = useT ? 1 : P.cols, sy = useT ? P.rows : 1;
    const long long c0 = clock64();
    while (t < P.max_range) {
        const int px = __float2int_rz(fmaf(dx, t, x0)), py = __float2int_rz(fmaf(dy, t, y0));
        if ((unsigned)px >= (unsigned)P.rows || (unsigned)py >= (unsigned)P.cols) break;
        const float d = __ldg(base + (px * sx + py * sy));
        ++steps;
        if (d <= 0.0f) {
            const float xd = __fsub_rn((float)px, x0), yd = __fsub_rn((float)py, y0);
            r = sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
            break;
        }
        t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
    }
    const long long c1 = clock64();
    a.outs[i] = __fmul_rn(r, P.w.scale);
    const unsigned long long tend = gtime();
    // end-time histogram in microseconds since the first thread started (filled in a second pass on the host)
    D.t_last[i >> 5] = tend;      // per warp (last writer wins, all lanes end together)
    D.t_first[i >> 5] = tstart;
    if (steps > 48) {
        const unsigned slot = atomicAdd(D.n_long, 1u);
        if (slot < 65536) { D.long_cycles[slot] = (unsigned long long)(c1 - c0); D.long_steps[slot] = steps; }
    }
}

// ---------------------------------------------------------------- V10: tail mode.  After TAIL_AFTER plain steps a
// ray switches to a loop that, besides each real sample, loads the cell AHEAD px further along the
// ray into a ring of 4 registers that are only consumed 4 steps later, so the touch never stalls the
// warp and the real sample finds its sector in L1.
template <int AHEAD, int TAIL_AFTER, bool TOUCH = true>
__global__ void __launch_bounds__(128) k_tail(Args a, Lean q)
{
    const unsigned i = blockIdx.x * 128u + threadIdx.x;
    const unsigned total = (unsigned)a.num_poses * a.num_beams;
    if (i >= total) return;
    const MarchParams &P = a.P;
    float x0, y0, dx, dy;
    ray_setup_fan(a, q, i, x0, y0, dx, dy);
    float t = 0.f, r = P.max_range;
    bool done = !((x0 == x0) && (y0 == y0) && (dx == dx));
    int it = 0;
    while (!done) {
        const int px = __float2int_rz(fmaf(dx, t, x0)), py = __float2int_rz(fmaf(dy, t, y0));
        if ((unsigned)px >= (unsigned)P.rows || (unsigned)py >= (unsigned)P.cols) { done = true; break; }
        const float d = __ldg(P.dist + (px * P.cols + py));
        if (d <= 0.0f) {
            const float xd = __fsub_rn((float)px, x0), yd = __fsub_rn((float)py, y0);
            r = sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
            done = true; break;
        }
        t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
        if (!(t < P.max_range)) { done = true; break; }
        if (++it == TAIL_AFTER) break;
    }
    if (!done) {
        float j0 = 0.f, j1 = 0.f, j2 = 0.f, j3 = 0.f, keep = 0.f;
        const float adx = dx * (float)AHEAD, ady = dy * (float)AHEAD;
#define TAIL_STEP(J)                                                                                  \
        {                                                                                             \
            const float fx = fmaf(dx, t, x0), fy = fmaf(dy, t, y0);                                   \
            const int px = __float2int_rz(fx), py = __float2int_rz(fy);                               \
            if ((unsigned)px >= (unsigned)P.rows || (unsigned)py >= (unsigned)P.cols) break;          \
            const float d = __ldg(P.dist + (px * P.cols + py));                                       \
            const int ax = __float2int_rz(fx + adx), ay = __float2int_rz(fy + ady);                   \
            keep += J;                                                                                \
            if (TOUCH && (unsigned)ax < (unsigned)P.rows && (unsigned)ay < (unsigned)P.cols)          \
                asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(J) : "l"(P.dist + (ax * P.cols + ay))); \
            if (d <= 0.0f) {                                                                          \
                const float xd = __fsub_rn((float)px, x0), yd = __fsub_rn((float)py, y0);             \
                r = sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));                                           \
                break;                                                                                \
            }                                                                                         \
            t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));                                      \
            if (!(t < P.max_range)) break;                                                            \
        }
        for (;;) { TAIL_STEP(j0) TAIL_STEP(j1) TAIL_STEP(j2) TAIL_STEP(j3) }
#undef TAIL_STEP
        if (keep + j0 + j1 + j2 + j3 < 0.0f) r = -1.0f;   // never true (distances are >= 0): keeps the touches alive
    }
    a.outs[i] = __fmul_rn(r, P.w.scale);
}

// ---------------------------------------------------------------- V11: persistent warps, static striding over
// 32-ray groups for the first `static_groups`, then a shared atomic counter for the rest
__global__ void __launch_bounds__(128) k_pstatic(Args a, Lean q, unsigned static_groups)
{
    const MarchParams &P = a.P;
    const unsigned total = (unsigned)a.num_poses * a.num_beams;
    const unsigned ngroups = (total + 31) / 32;
    const unsigned lane = threadIdx.x & 31;
    const unsigned nwarps = gridDim.x * 4;
    unsigned g = blockIdx.x * 4 + (threadIdx.x >> 5);
    bool dynamic = false;
    for (;;) {
        if (!dynamic) {
            if (g >= static_groups) {
                dynamic = true;
                continue;
            }
        } else {
            unsigned nx = 0;
            if (lane == 0) nx = atomicAdd(a.counter, 1u);
            g = static_groups + __shfl_sync(0xffffffffu, nx, 0);
            if (g >= ngroups) break;
        }
        const unsigned i = g * 32 + lane;
        if (i < total) {
            float x0, y0, dx, dy;
            ray_setup_fan(a, q, i, x0, y0, dx, dy);
            const rl::FirstSample f0 = rl::first_sample(P, x0, y0);
            uint32_t st = 0;
            a.outs[i] = __fmul_rn(rl::march_ray<false>(P, x0, y0, dx, dy, st, f0), P.w.scale);
        }
        if (!dynamic) g += nwarps;
    }
}

// ---------------------------------------------------------------- V17: the product loop with an interior variant
// (poses >= max_range + 1 px from every border skip the bounds tests: 13 instead of 16 instructions per step).
// Measured 86.1 vs 84.8 us: no gain, the loop is bound by load latency, not by issue slots -- not adopted.
// The loop after the first sample.  CHECK = false is the interior variant: the caller has proved that no
// sample with t < max_range can leave the map, so the two bounds compares and their branch (3 of the 16
// instructions of a step) are compiled out.  Same samples, same arithmetic, same result.
template <bool COUNT, bool CHECK>
__device__ __forceinline__ float march_loop(const MarchParams &P, float x0, float y0, float dx, float dy,
                                            float t, uint32_t &steps)
{
    // One exit branch per step: t is advanced before the hit test (harmless: a hit ends the ray) and
    // which of the two exits it was is decided once, after the loop.
    int px, py, it = 1;
    float d;
    bool tail = false;
    for (;;) {
        px = __float2int_rz(fmaf(dx, t, x0));
        py = __float2int_rz(fmaf(dy, t, y0));
        if (CHECK && ((unsigned)px >= (unsigned)P.rows || (unsigned)py >= (unsigned)P.cols)) return P.max_range;
        d = __ldg(P.dist + (px * P.cols + py));
        if (COUNT) ++steps;
        t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
        if (d <= 0.0f || !(t < P.max_range)) break;
        if (++it == rl::TAIL_AFTER) { tail = true; break; }
    }
    if (tail) {
        float j0 = 0.f, j1 = 0.f, j2 = 0.f, j3 = 0.f, keep = 0.f;
        const float adx = __fmul_rn(dx, (float)rl::TAIL_AHEAD), ady = __fmul_rn(dy, (float)rl::TAIL_AHEAD);
        bool inside = true;
#define RL_TAIL_STEP(J)                                                                            \
        {                                                                                          \
            const float fx = fmaf(dx, t, x0), fy = fmaf(dy, t, y0);                                \
            px = __float2int_rz(fx);                                                               \
            py = __float2int_rz(fy);                                                               \
            if (CHECK && ((unsigned)px >= (unsigned)P.rows || (unsigned)py >= (unsigned)P.cols)) { inside = false; break; } \
            d = __ldg(P.dist + (px * P.cols + py));                                                \
            if (COUNT) ++steps;                                                                    \
            const int ax = __float2int_rz(__fadd_rn(fx, adx)), ay = __float2int_rz(__fadd_rn(fy, ady)); \
            keep = __fadd_rn(keep, J);                                                             \
            if ((unsigned)ax < (unsigned)P.rows && (unsigned)ay < (unsigned)P.cols)                \
                asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(J) : "l"(P.dist + (ax * P.cols + ay))); \
            t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));                                   \
            if (d <= 0.0f || !(t < P.max_range)) break;                                            \
        }
        for (;;) { RL_TAIL_STEP(j0) RL_TAIL_STEP(j1) RL_TAIL_STEP(j2) RL_TAIL_STEP(j3) }
#undef RL_TAIL_STEP
        if (__fadd_rn(__fadd_rn(keep, j0), __fadd_rn(j1, __fadd_rn(j2, j3))) < 0.0f) return -1.0f;  // never true
        if (!inside) return P.max_range;
    }
    if (d <= 0.0f) {
        const float xd = __fsub_rn((float)px, x0);
        const float yd = __fsub_rn((float)py, y0);
        return sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
    }
    return P.max_range;
}

// Interior test: |dx|, |dy| <= 1 (they are sinf/cosf values), so a sample at parameter t < max_range lies
// within max_range (plus a rounding error far below the 1 px margin) of the pose along each axis.  A pose
// at least max_range + 1 px from every border of the map can therefore never produce an out-of-map sample
// and its rays take the loop without bounds tests.  The decision depends on the pose only, so the warps of
// a pose (almost all warps: one warp rarely straddles two poses) stay convergent.
__device__ __forceinline__ bool pose_is_interior(const MarchParams &P, float x0, float y0)
{
    const float margin = fminf(fminf(x0, y0), fminf(__fsub_rn(P.frows, x0), __fsub_rn(P.fcols, y0)));
    return __fsub_rn(margin, 1.0f) >= P.max_range;   // false for NaN
}

template <bool COUNT, bool INTERIOR_PATH>
__device__ __forceinline__ float march_ray_v17(const MarchParams &P, float x0, float y0, float dx,
                                           float dy, uint32_t &steps, const rl::FirstSample &f0)
{
    if (!f0.inside || !(dx == dx) || !(dy == dy)) return P.max_range;   // NaN pose/heading or pose outside the map
    if (COUNT) ++steps;
    if (f0.d <= 0.0f) {   // pose inside an occupied cell: distance to that cell's corner
        const float xd = __fsub_rn((float)f0.px, x0);
        const float yd = __fsub_rn((float)f0.py, y0);
        return sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
    }
    const float t = fmaxf(__fmul_rn(f0.d, 0.999f), 1.0f);   // 0 + step
    if (!(t < P.max_range)) return P.max_range;
    if (INTERIOR_PATH && pose_is_interior(P, x0, y0)) return march_loop<COUNT, false>(P, x0, y0, dx, dy, t, steps);
    return march_loop<COUNT, true>(P, x0, y0, dx, dy, t, steps);
}

template <bool INTERIOR>
__global__ void __launch_bounds__(128) k_product_like_t(Args a, Lean q)
{
    const unsigned i = blockIdx.x * 128u + threadIdx.x;
    const unsigned total = (unsigned)a.num_poses * a.num_beams;
    if (i >= total) return;
    float x0, y0, dx, dy;
    ray_setup_fan(a, q, i, x0, y0, dx, dy);
    const rl::FirstSample f0 = rl::first_sample(a.P, x0, y0);
    uint32_t st = 0;
    a.outs[i] = __fmul_rn(march_ray_v17<false, INTERIOR>(a.P, x0, y0, dx, dy, st, f0), a.P.w.scale);
}

__global__ void __launch_bounds__(128) k_product_like(Args a, Lean q)
{
    const unsigned i = blockIdx.x * 128u + threadIdx.x;
    const unsigned total = (unsigned)a.num_poses * a.num_beams;
    if (i >= total) return;
    float x0, y0, dx, dy;
    ray_setup_fan(a, q, i, x0, y0, dx, dy);
    const rl::FirstSample f0 = rl::first_sample(a.P, x0, y0);
    uint32_t st = 0;
    a.outs[i] = __fmul_rn(rl::march_ray<false>(a.P, x0, y0, dx, dy, st, f0), a.P.w.scale);
}

// ---------------------------------------------------------------- V12: one 1024-thread CTA per pose with a TILE x TILE
// window of the distance field around the pose staged in shared memory (SURVEY.md build plan step 8:
// "keep only if it wins in measurement").  Samples inside the window read shared memory, others global.
template <int TILE>
__global__ void __launch_bounds__(1024) k_tile(Args a, Lean q)
{
    __shared__ float tile[TILE * TILE];
    const MarchParams &P = a.P;
    const int k = blockIdx.x;
    const float *p = a.poses + 3 * k;
    const float thw = __ldg(p + 2);
    const GridPose g = rl::world_to_grid(P.w, __ldg(p), __ldg(p + 1), thw);
    const float x0 = g.y, y0 = g.x;
    const int cx = __float2int_rz(x0), cy = __float2int_rz(y0);
    const int tx0 = cx - TILE / 2, ty0 = cy - TILE / 2;
    for (int e = threadIdx.x; e < TILE * TILE; e += 1024) {
        const int r = tx0 + e / TILE, c = ty0 + e % TILE;
        tile[e] = ((unsigned)r < (unsigned)P.rows && (unsigned)c < (unsigned)P.cols) ? __ldg(P.dist + (r * P.cols + c)) : 1.0f;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < a.num_beams; j += 1024) {
        const float thg = __fadd_rn(-__fadd_rn(thw, fmaf((float)j, q.inc, -0.5f * a.fov)), P.w.rotation_const);
        float dx, dy;
        rl::glibc_sincosf(thg, &dy, &dx);
        float t = 0.f, r = P.max_range;
        if ((x0 == x0) && (y0 == y0) && (dx == dx)) {
            while (t < P.max_range) {
                const int px = __float2int_rz(fmaf(dx, t, x0)), py = __float2int_rz(fmaf(dy, t, y0));
                if ((unsigned)px >= (unsigned)P.rows || (unsigned)py >= (unsigned)P.cols) break;
                const unsigned lx = (unsigned)(px - tx0), ly = (unsigned)(py - ty0);
                float d;
                if (lx < (unsigned)TILE && ly < (unsigned)TILE) d = tile[lx * TILE + ly];
                else d = __ldg(P.dist + (px * P.cols + py));
                if (d <= 0.0f) {
                    const float xd = __fsub_rn((float)px, x0), yd = __fsub_rn((float)py, y0);
                    r = sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
                    break;
                }
                t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
            }
        }
        a.outs[(size_t)k * a.num_beams + j] = __fmul_rn(r, P.w.scale);
    }
}

// ---------------------------------------------------------------- V13: product loop with the hit / max-range exits
// merged into one branch per step (t is advanced speculatively; which exit it was is sorted out after the loop)
__global__ void __launch_bounds__(128, 16) k_merged(Args a, Lean q)
{
    const unsigned i = blockIdx.x * 128u + threadIdx.x;
    const unsigned total = (unsigned)a.num_poses * a.num_beams;
    if (i >= total) return;
    const MarchParams &P = a.P;
    float x0, y0, dx, dy;
    ray_setup_fan(a, q, i, x0, y0, dx, dy);
    float r = P.max_range;
    if ((x0 == x0) && (y0 == y0) && (dx == dx)) {
        float t = 0.f, d = 1.f;
        int px = 0, py = 0;
        bool inb = true;
        for (;;) {
            px = __float2int_rz(fmaf(dx, t, x0));
            py = __float2int_rz(fmaf(dy, t, y0));
            inb = (unsigned)px < (unsigned)P.rows && (unsigned)py < (unsigned)P.cols;
            if (!inb) break;
            d = __ldg(P.dist + (px * P.cols + py));
            t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
            if (d <= 0.0f || !(t < P.max_range)) break;
        }
        if (inb && d <= 0.0f) {
            const float xd = __fsub_rn((float)px, x0), yd = __fsub_rn((float)py, y0);
            r = sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
        }
    }
    a.outs[i] = __fmul_rn(r, P.w.scale);
}

// ---------------------------------------------------------------- V14: value-predicted tail.  Creeping rays see the
// same clearance d step after step (90 % of tail steps on axis-aligned walls), so after TAIL steps a ray
// loads, together with its real sample at t, the samples at t+s, t+2s, t+3s that it WOULD take if d
// repeated (s = step(d_prev)), then walks through them while the prediction holds: same t values bit
// for bit, but one memory round trip and one address chain per four steps instead of per step.
template <int TAIL, int DEPTH>
__global__ void __launch_bounds__(128) k_spec(Args a, Lean q)
{
    const unsigned i = blockIdx.x * 128u + threadIdx.x;
    const unsigned total = (unsigned)a.num_poses * a.num_beams;
    if (i >= total) return;
    const MarchParams &P = a.P;
    float x0, y0, dx, dy;
    ray_setup_fan(a, q, i, x0, y0, dx, dy);
    float r = P.max_range;
    if ((x0 == x0) && (y0 == y0) && (dx == dx)) {
        float t = 0.f, d = 1.f;
        int px = 0, py = 0, it = 0;
        bool inb = true, tail = false;
        for (;;) {
            px = __float2int_rz(fmaf(dx, t, x0));
            py = __float2int_rz(fmaf(dy, t, y0));
            inb = (unsigned)px < (unsigned)P.rows && (unsigned)py < (unsigned)P.cols;
            if (!inb) break;
            d = __ldg(P.dist + (px * P.cols + py));
            t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
            if (d <= 0.0f || !(t < P.max_range)) break;
            if (++it == TAIL) { tail = true; break; }
        }
        if (tail) {
            // here: d = last clearance (> 0), t = parameter of the next sample (< max_range)
            bool finished = false;
            while (!finished) {
                const float s = fmaxf(__fmul_rn(d, 0.999f), 1.0f);
                float tt[DEPTH], vv[DEPTH];
                int cx[DEPTH], cy[DEPTH];
                bool ok[DEPTH];
                tt[0] = t;
#pragma unroll
                for (int k = 1; k < DEPTH; ++k) tt[k] = __fadd_rn(tt[k - 1], s);
#pragma unroll
                for (int k = 0; k < DEPTH; ++k) {
                    cx[k] = __float2int_rz(fmaf(dx, tt[k], x0));
                    cy[k] = __float2int_rz(fmaf(dy, tt[k], y0));
                    ok[k] = (unsigned)cx[k] < (unsigned)P.rows && (unsigned)cy[k] < (unsigned)P.cols;
                    vv[k] = ok[k] ? __ldg(P.dist + (cx[k] * P.cols + cy[k])) : 0.0f;
                }
                const float dprev = d;
#pragma unroll
                for (int k = 0; k < DEPTH; ++k) {
                    // sample k is real iff every earlier sample of this round repeated dprev
                    px = cx[k]; py = cy[k];
                    if (!ok[k]) { inb = false; finished = true; break; }
                    d = vv[k];
                    t = __fadd_rn(tt[k], fmaxf(__fmul_rn(d, 0.999f), 1.0f));
                    if (d <= 0.0f || !(t < P.max_range)) { finished = true; break; }
                    if (d != dprev) break;   // prediction ends here: next round starts from the true t
                }
            }
        }
        if (inb && d <= 0.0f) {
            const float xd = __fsub_rn((float)px, x0), yd = __fsub_rn((float)py, y0);
            r = sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
        }
    }
    a.outs[i] = __fmul_rn(r, P.w.scale);
}

// ---------------------------------------------------------------- V15: value-predicted tail, registers only
template <int TAIL>
__global__ void __launch_bounds__(128) k_spec2(Args a, Lean q)
{
    const unsigned i = blockIdx.x * 128u + threadIdx.x;
    const unsigned total = (unsigned)a.num_poses * a.num_beams;
    if (i >= total) return;
    const MarchParams &P = a.P;
    float x0, y0, dx, dy;
    ray_setup_fan(a, q, i, x0, y0, dx, dy);
    float r = P.max_range;
    if ((x0 == x0) && (y0 == y0) && (dx == dx)) {
        float t = 0.f, d = 1.f, th = 0.f;   // th: parameter of the sample that hit
        int it = 0;
        int state = 0;                      // 0 running, 1 hit at th, 2 max range / left the map
        for (;;) {
            const int px = __float2int_rz(fmaf(dx, t, x0)), py = __float2int_rz(fmaf(dy, t, y0));
            if ((unsigned)px >= (unsigned)P.rows || (unsigned)py >= (unsigned)P.cols) { state = 2; break; }
            d = __ldg(P.dist + (px * P.cols + py));
            th = t;
            t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
            if (d <= 0.0f) { state = 1; break; }
            if (!(t < P.max_range)) { state = 2; break; }
            if (++it == TAIL) break;
        }
        while (state == 0) {
            const float s = fmaxf(__fmul_rn(d, 0.999f), 1.0f);
            const float ta = t, tb = __fadd_rn(ta, s), tc = __fadd_rn(tb, s), td = __fadd_rn(tc, s);
            float va = 0.f, vb = 0.f, vc = 0.f, vd = 0.f;
            unsigned ok = 0;
#define SPEC_LOAD(T, V, BIT)                                                                         \
            {                                                                                        \
                const int cx = __float2int_rz(fmaf(dx, T, x0)), cy = __float2int_rz(fmaf(dy, T, y0)); \
                if ((unsigned)cx < (unsigned)P.rows && (unsigned)cy < (unsigned)P.cols) {            \
                    V = __ldg(P.dist + (cx * P.cols + cy));                                          \
                    ok |= BIT;                                                                       \
                }                                                                                    \
            }
            SPEC_LOAD(ta, va, 1u) SPEC_LOAD(tb, vb, 2u) SPEC_LOAD(tc, vc, 4u) SPEC_LOAD(td, vd, 8u)
#undef SPEC_LOAD
            const float dprev = d;
#define SPEC_USE(T, V, BIT)                                                                          \
            if (!(ok & BIT)) { state = 2; break; }                                                   \
            d = V; th = T;                                                                           \
            t = __fadd_rn(T, fmaxf(__fmul_rn(d, 0.999f), 1.0f));                                     \
            if (d <= 0.0f) { state = 1; break; }                                                     \
            if (!(t < P.max_range)) { state = 2; break; }                                            \
            if (d != dprev) continue;
            SPEC_USE(ta, va, 1u) SPEC_USE(tb, vb, 2u) SPEC_USE(tc, vc, 4u) SPEC_USE(td, vd, 8u)
#undef SPEC_USE
        }
        if (state == 1) {
            const int px = __float2int_rz(fmaf(dx, th, x0)), py = __float2int_rz(fmaf(dy, th, y0));
            const float xd = __fsub_rn((float)px, x0), yd = __fsub_rn((float)py, y0);
            r = sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
        }
    }
    a.outs[i] = __fmul_rn(r, P.w.scale);
}

// ---------------------------------------------------------------- V16 (round 1, second session): three ideas on top of
// the product loop, separately switchable.
//  SAFE : a conservative t_safe below which no sample can leave the map (|dx|,|dy| <= 1, so a ray is still
//         >= 1 px inside every border while t < min(x0, y0, rows-x0, cols-y0) - 1): steps below it skip the two
//         bounds compares + branch (3 of 16 instructions per step).
//  COOP : after COOP steps in lockstep the lanes whose rays are still marching are served by the whole
//         warp: 32/m lanes per surviving ray sample t, t+s, t+2s, ... (s = the ray's previous step), a ballot
//         finds how far the "same step again" prediction held, and the ray advances that many steps in one
//         memory round trip.  Same t sequence bit for bit (each lane adds s sequentially).
//  PERSIST: persistent CTAs; a warp takes CH consecutive 32-ray groups per global atomic (CH = 1 for the
//         last part of the batch), so all 64 warp slots of an SM stay busy and L1 locality is kept.
struct RayState { float x0, y0, dx, dy, t, d; int px, py; int st; };   // st 0 running, 1 hit at (px,py), 2 max range

template <bool SAFE, int LIMIT>
__device__ __forceinline__ void march_limited(const MarchParams &P, RayState &r)
{
    // first sample (t = 0)
    r.st = 2; r.t = 0.f; r.d = 1.f; r.px = 0; r.py = 0;
    if (!((r.x0 == r.x0) && (r.y0 == r.y0) && (r.dx == r.dx) && (r.dy == r.dy))) return;
    int px = __float2int_rz(r.x0), py = __float2int_rz(r.y0);
    if ((unsigned)px >= (unsigned)P.rows || (unsigned)py >= (unsigned)P.cols) return;
    float d = __ldg(P.dist + (px * P.cols + py));
    if (d <= 0.0f) { r.st = 1; r.px = px; r.py = py; return; }
    float t = fmaxf(__fmul_rn(d, 0.999f), 1.0f);
    if (!(t < P.max_range)) return;
    int it = 1;
    if (SAFE) {
        const float t_safe = __fadd_rn(fminf(fminf(r.x0, r.y0), fminf(__fsub_rn(P.frows, r.x0), __fsub_rn(P.fcols, r.y0))), -1.0f);
        const float t_lim = fminf(t_safe, P.max_range);
        if (t < t_lim) {
            for (;;) {
                px = __float2int_rz(fmaf(r.dx, t, r.x0));
                py = __float2int_rz(fmaf(r.dy, t, r.y0));
                d = __ldg(P.dist + (px * P.cols + py));
                t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
                if (d <= 0.0f) { r.st = 1; r.px = px; r.py = py; return; }
                if (!(t < t_lim)) break;
                if (++it == LIMIT) { r.st = 0; r.t = t; r.d = d; return; }
            }
            if (!(t < P.max_range)) return;
            if (++it == LIMIT) { r.st = 0; r.t = t; r.d = d; return; }
        }
    }
    for (;;) {
        px = __float2int_rz(fmaf(r.dx, t, r.x0));
        py = __float2int_rz(fmaf(r.dy, t, r.y0));
        if ((unsigned)px >= (unsigned)P.rows || (unsigned)py >= (unsigned)P.cols) return;
        d = __ldg(P.dist + (px * P.cols + py));
        t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
        if (d <= 0.0f) { r.st = 1; r.px = px; r.py = py; return; }
        if (!(t < P.max_range)) return;
        if (++it == LIMIT) { r.st = 0; r.t = t; r.d = d; return; }
    }
}

// plain continuation (no cooperation): checked loop with the product's look-ahead touches
__device__ __forceinline__ void march_rest_plain(const MarchParams &P, RayState &r)
{
    float t = r.t, d;
    int px, py;
    float j0 = 0.f, j1 = 0.f, j2 = 0.f, j3 = 0.f, keep = 0.f;
    const float adx = __fmul_rn(r.dx, 12.0f), ady = __fmul_rn(r.dy, 12.0f);
#define V16_STEP(J)                                                                                \
    {                                                                                              \
        const float fx = fmaf(r.dx, t, r.x0), fy = fmaf(r.dy, t, r.y0);                            \
        px = __float2int_rz(fx); py = __float2int_rz(fy);                                          \
        if ((unsigned)px >= (unsigned)P.rows || (unsigned)py >= (unsigned)P.cols) { r.st = 2; break; } \
        d = __ldg(P.dist + (px * P.cols + py));                                                    \
        const int ax = __float2int_rz(__fadd_rn(fx, adx)), ay = __float2int_rz(__fadd_rn(fy, ady)); \
        keep = __fadd_rn(keep, J);                                                                 \
        if ((unsigned)ax < (unsigned)P.rows && (unsigned)ay < (unsigned)P.cols)                    \
            asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(J) : "l"(P.dist + (ax * P.cols + ay))); \
        t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));                                       \
        if (d <= 0.0f) { r.st = 1; r.px = px; r.py = py; break; }                                  \
        if (!(t < P.max_range)) { r.st = 2; break; }                                               \
    }
    for (;;) { V16_STEP(j0) V16_STEP(j1) V16_STEP(j2) V16_STEP(j3) }
#undef V16_STEP
    if (__fadd_rn(__fadd_rn(keep, j0), __fadd_rn(j1, __fadd_rn(j2, j3))) < 0.0f) r.st = 3;   // never true
}

// warp-cooperative continuation.  Must be called by all 32 lanes; `unfinished` = this lane's ray is running
// (r.t = parameter of its next sample, < max_range; r.d = clearance of its previous sample, > 0).
__device__ __forceinline__ void march_rest_coop(const MarchParams &P, RayState &r, bool unfinished)
{
    const unsigned lane = threadIdx.x & 31u;
    unsigned m = __ballot_sync(0xffffffffu, unfinished);
    while (m) {
        const int ma = __popc(m);
        const int lg = ma <= 1 ? 0 : 32 - __clz(ma - 1);   // ceil(log2(ma)), 0..5
        const int G = 32 >> lg;                            // lanes per served ray
        const int g = lane >> (5 - lg);                    // group of this lane
        const int k = lane & (G - 1);                      // position in the group
        const bool has = g < ma;
        const int owner = has ? (int)__fns(m, 0, g + 1) : 0;
        const float ox0 = __shfl_sync(0xffffffffu, r.x0, owner), oy0 = __shfl_sync(0xffffffffu, r.y0, owner);
        const float odx = __shfl_sync(0xffffffffu, r.dx, owner), ody = __shfl_sync(0xffffffffu, r.dy, owner);
        const float ot = __shfl_sync(0xffffffffu, r.t, owner), od = __shfl_sync(0xffffffffu, r.d, owner);
        const float s = fmaxf(__fmul_rn(od, 0.999f), 1.0f);
        float tk = ot;
        for (int j = 1; j < G; ++j) tk = (j <= k) ? __fadd_rn(tk, s) : tk;
        // sample k of the group
        const int px = __float2int_rz(fmaf(odx, tk, ox0)), py = __float2int_rz(fmaf(ody, tk, oy0));
        const bool live = tk < P.max_range;    // lane 0: true by invariant
        const bool inb = (unsigned)px < (unsigned)P.rows && (unsigned)py < (unsigned)P.cols;
        float d = 1.0f;
        if (has && live && inb) d = __ldg(P.dist + (px * P.cols + py));
        const float sk = fmaxf(__fmul_rn(d, 0.999f), 1.0f);
        const bool good = live && inb && d > 0.0f && sk == s;           // the next lane's position is right
        const unsigned bad = __ballot_sync(0xffffffffu, !good);
        const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << (g * G);
        const unsigned gb = bad & gmask;
        const int f = gb ? (__ffs(gb) - 1) : (g * G + G - 1);           // lane that decides the ray's new state
        // state as seen by the deciding lane
        int st = 0; float nt = __fadd_rn(tk, sk), nd = d;
        if (!live) st = 2;
        else if (!inb) st = 2;
        else if (d <= 0.0f) st = 1;
        else if (!(nt < P.max_range)) st = 2;
        // ship it to the owner: the owner lane reads from lane f of ITS group
        // (owner's group index = rank of the owner among the set bits of m)
        const int myrank = __popc(m & ((1u << lane) - 1u));
        const unsigned gmask_o = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << (myrank * G);
        const unsigned gb_o = bad & gmask_o;
        const int src = unfinished ? (gb_o ? (__ffs(gb_o) - 1) : (myrank * G + G - 1)) : (int)lane;
        const int st_o = __shfl_sync(0xffffffffu, st, src);
        const float nt_o = __shfl_sync(0xffffffffu, nt, src), nd_o = __shfl_sync(0xffffffffu, nd, src);
        const int px_o = __shfl_sync(0xffffffffu, px, src), py_o = __shfl_sync(0xffffffffu, py, src);
        (void)f;
        if (unfinished) {
            r.st = st_o; r.t = nt_o; r.d = nd_o; r.px = px_o; r.py = py_o;
            unfinished = st_o == 0;
        }
        m = __ballot_sync(0xffffffffu, unfinished);
    }
}

__device__ __forceinline__ float ray_result(const MarchParams &P, const RayState &r)
{
    if (r.st == 1) {
        const float xd = __fsub_rn((float)r.px, r.x0), yd = __fsub_rn((float)r.py, r.y0);
        return sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
    }
    return r.st == 3 ? -1.0f : P.max_range;
}

template <bool SAFE, int COOP>   // COOP = 0: plain continuation after 32 steps (product behaviour)
__global__ void __launch_bounds__(128) k_v16(Args a, Lean q)
{
    const unsigned i = blockIdx.x * 128u + threadIdx.x;
    const unsigned total = (unsigned)a.num_poses * a.num_beams;
    const bool valid = i < total;
    RayState r;
    bool unfinished = false;
    if (valid) {
        ray_setup_fan(a, q, i, r.x0, r.y0, r.dx, r.dy);
        march_limited<SAFE, COOP ? COOP : 32>(a.P, r);
        unfinished = r.st == 0;
    }
    if (COOP) march_rest_coop(a.P, r, unfinished);
    else if (unfinished) march_rest_plain(a.P, r);
    if (valid) a.outs[i] = __fmul_rn(ray_result(a.P, r), a.P.w.scale);
}

template <bool SAFE, int COOP, int CH>
__global__ void __launch_bounds__(128, 16) k_v16p(Args a, Lean q, unsigned ngroups, unsigned chunked_groups)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned total = (unsigned)a.num_poses * a.num_beams;
    const unsigned nchunks = chunked_groups / CH;    // chunked_groups is a multiple of CH
    unsigned v = 0;
    if (lane == 0) v = atomicAdd(a.counter, 1u);
    v = __shfl_sync(0xffffffffu, v, 0);
    for (;;) {
        unsigned g0, g1;
        if (v < nchunks) { g0 = v * CH; g1 = g0 + CH; }
        else { g0 = chunked_groups + (v - nchunks); g1 = g0 + 1; }
        if (g0 >= ngroups) break;
        // fetch the next ticket now; it is needed only after this chunk
        unsigned vn = 0;
        if (lane == 0) vn = atomicAdd(a.counter, 1u);
        for (unsigned g = g0; g < g1; ++g) {
            const unsigned i = g * 32u + lane;
            const bool valid = i < total;
            RayState r;
            bool unfinished = false;
            if (valid) {
                ray_setup_fan(a, q, i, r.x0, r.y0, r.dx, r.dy);
                march_limited<SAFE, COOP ? COOP : 32>(a.P, r);
                unfinished = r.st == 0;
            }
            if (COOP) march_rest_coop(a.P, r, unfinished);
            else if (unfinished) march_rest_plain(a.P, r);
            if (valid) a.outs[i] = __fmul_rn(ray_result(a.P, r), a.P.w.scale);
            __syncwarp();
        }
        v = __shfl_sync(0xffffffffu, vn, 0);
    }
}

// ---------------------------------------------------------------- V18: memory-level parallelism.  The march is a chain
// of dependent loads and a warp has one of them in flight; with the register file full at 32 registers x 2048
// threads the only way to more loads in flight per SM is R rays per thread that share the pose (x0, y0), marched
// in lockstep with their R loads issued back to back.  A warp takes 32*R consecutive beams of ONE pose
// (chunks_per_pose = ceil(beams / 32R)); rays still marching after 32 lockstep steps finish one after another
// through the product's look-ahead loop.
template <int R>
__global__ void __launch_bounds__(128) k_mlp(Args a, Lean q, unsigned chunks_per_pose, unsigned cmagic, unsigned cshift)
{
    const MarchParams &P = a.P;
    const unsigned w = (blockIdx.x * 128u + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    const unsigned k = chunks_per_pose == 1 ? w : (__umulhi(w, cmagic) >> cshift);
    if (k >= (unsigned)a.num_poses) return;
    const unsigned c = w - k * chunks_per_pose;
    const float *p = a.poses + 3 * k;
    const float thw = __ldg(p + 2);
    const GridPose g = rl::world_to_grid(P.w, __ldg(p), __ldg(p + 1), thw);
    const float x0 = g.y, y0 = g.x;
    const rl::FirstSample f0 = rl::first_sample(P, x0, y0);
    const int jb = (int)(c * 32u * R + lane);
    float dx[R], dy[R], t[R], res[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int j = jb + 32 * r;
        const float thg = __fadd_rn(-__fadd_rn(thw, fmaf((float)j, q.inc, -0.5f * a.fov)), P.w.rotation_const);
        rl::glibc_sincosf(thg, &dy[r], &dx[r]);
        res[r] = P.max_range;
    }
    const float INF = __int_as_float(0x7f800000);
    float t_first = INF;
    if (f0.inside) {
        if (f0.d <= 0.0f) {
            const float xd = __fsub_rn((float)f0.px, x0), yd = __fsub_rn((float)f0.py, y0);
            const float r0 = sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
#pragma unroll
            for (int r = 0; r < R; ++r) if ((dx[r] == dx[r]) && (dy[r] == dy[r])) res[r] = r0;
        } else {
            t_first = fmaxf(__fmul_rn(f0.d, 0.999f), 1.0f);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const bool ok = (dx[r] == dx[r]) && (dy[r] == dy[r]) && (jb + 32 * r < a.num_beams);
        t[r] = ok ? t_first : INF;     // t >= max_range (or INF) = finished
    }
    int it = 1;
    for (;;) {
        bool live[R];
        bool any = false;
#pragma unroll
        for (int r = 0; r < R; ++r) { live[r] = t[r] < P.max_range; any |= live[r]; }
        if (!any) break;
        if (it == 32) break;
        ++it;
        int px[R], py[R];
        float d[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            px[r] = __float2int_rz(fmaf(dx[r], t[r], x0));
            py[r] = __float2int_rz(fmaf(dy[r], t[r], y0));
            const bool inb = (unsigned)px[r] < (unsigned)P.rows && (unsigned)py[r] < (unsigned)P.cols;
            d[r] = 1.0f;
            if (live[r] && !inb) { t[r] = INF; live[r] = false; }
            if (live[r]) d[r] = __ldg(P.dist + (px[r] * P.cols + py[r]));
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (live[r]) {
                if (d[r] <= 0.0f) {
                    const float xd = __fsub_rn((float)px[r], x0), yd = __fsub_rn((float)py[r], y0);
                    res[r] = sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
                    t[r] = INF;
                } else {
                    t[r] = __fadd_rn(t[r], fmaxf(__fmul_rn(d[r], 0.999f), 1.0f));
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (t[r] < P.max_range) {      // still marching after 32 steps: the product's tail loop
            RayState s;
            s.x0 = x0; s.y0 = y0; s.dx = dx[r]; s.dy = dy[r]; s.t = t[r]; s.d = 1.f; s.px = 0; s.py = 0; s.st = 0;
            march_rest_plain(P, s);
            res[r] = ray_result(P, s);
        }
    }
    float *o = a.outs + (size_t)k * a.num_beams;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int j = jb + 32 * r;
        if (j < a.num_beams) o[j] = __fmul_rn(res[r], P.w.scale);
    }
}

// ---------------------------------------------------------------- diag2: the PRODUCT march (tail mode included) with
// per-warp start/end stamps and per-ray cycle counts, to see what the end of the kernel consists of now.
struct Diag2 { unsigned long long *t_first, *t_last; unsigned *max_steps; unsigned *ray_steps; unsigned *ray_cycles; };

__global__ void __launch_bounds__(128) k_diag2(Args a, Lean q, Diag2 D)
{
    const unsigned i = blockIdx.x * 128u + threadIdx.x;
    const unsigned total = (unsigned)a.num_poses * a.num_beams;
    if (i >= total) return;
    const unsigned long long tstart = gtime();
    float x0, y0, dx, dy;
    ray_setup_fan(a, q, i, x0, y0, dx, dy);
    const rl::FirstSample f0 = rl::first_sample(a.P, x0, y0);
    uint32_t st = 0;
    const long long c0 = clock64();
    a.outs[i] = __fmul_rn(rl::march_ray<true>(a.P, x0, y0, dx, dy, st, f0), a.P.w.scale);
    const long long c1 = clock64();
    D.ray_steps[i] = st;
    D.ray_cycles[i] = (unsigned)(c1 - c0);
    const unsigned long long tend = gtime();
    D.t_first[i >> 5] = tstart;
    atomicMax(D.t_last + (i >> 5), tend);
    atomicMax(D.max_steps + (i >> 5), st);
}

// ---------------------------------------------------------------- V19: the product march with the tail touch aimed
// at the PREDICTED sample K steps ahead (t + K * s, s = the step just taken) instead of a fixed 12 px ahead:
// long rays creep along walls with the same clearance for 10-20 steps in a row, so the touched cell is the
// very cell the ray will sample K steps later (a fixed distance ahead falls between samples and, for rays
// that cross rows, into a different sector).  Touches never influence the result.
template <int AFTER, int K>
__device__ __forceinline__ float march_ray_v19(const MarchParams &P, float x0, float y0, float dx, float dy,
                                               const rl::FirstSample &f0)
{
    if (!f0.inside || !(dx == dx) || !(dy == dy)) return P.max_range;
    if (f0.d <= 0.0f) {
        const float xd = __fsub_rn((float)f0.px, x0), yd = __fsub_rn((float)f0.py, y0);
        return sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
    }
    float t = fmaxf(__fmul_rn(f0.d, 0.999f), 1.0f);
    if (!(t < P.max_range)) return P.max_range;
    int px, py, it = 1;
    float d;
    bool tail = false;
    for (;;) {
        px = __float2int_rz(fmaf(dx, t, x0));
        py = __float2int_rz(fmaf(dy, t, y0));
        if ((unsigned)px >= (unsigned)P.rows || (unsigned)py >= (unsigned)P.cols) return P.max_range;
        d = __ldg(P.dist + (px * P.cols + py));
        t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
        if (d <= 0.0f || !(t < P.max_range)) break;
        if (++it == AFTER) { tail = true; break; }
    }
    if (tail) {
        float j0 = 0.f, j1 = 0.f, j2 = 0.f, j3 = 0.f, keep = 0.f;
        float s = fmaxf(__fmul_rn(d, 0.999f), 1.0f);
        bool inside = true;
#define V19_STEP(J)                                                                                \
        {                                                                                          \
            px = __float2int_rz(fmaf(dx, t, x0));                                                  \
            py = __float2int_rz(fmaf(dy, t, y0));                                                  \
            if ((unsigned)px >= (unsigned)P.rows || (unsigned)py >= (unsigned)P.cols) { inside = false; break; } \
            d = __ldg(P.dist + (px * P.cols + py));                                                \
            const float tt = fmaf((float)K, s, t);                                                 \
            const int ax = __float2int_rz(fmaf(dx, tt, x0)), ay = __float2int_rz(fmaf(dy, tt, y0)); \
            keep = __fadd_rn(keep, J);                                                             \
            if ((unsigned)ax < (unsigned)P.rows && (unsigned)ay < (unsigned)P.cols)                \
                asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(J) : "l"(P.dist + (ax * P.cols + ay))); \
            s = fmaxf(__fmul_rn(d, 0.999f), 1.0f);                                                 \
            t = __fadd_rn(t, s);                                                                   \
            if (d <= 0.0f || !(t < P.max_range)) break;                                            \
        }
        for (;;) { V19_STEP(j0) V19_STEP(j1) V19_STEP(j2) V19_STEP(j3) }
#undef V19_STEP
        if (__fadd_rn(__fadd_rn(keep, j0), __fadd_rn(j1, __fadd_rn(j2, j3))) < 0.0f) return -1.0f;  // never true
        if (!inside) return P.max_range;
    }
    if (d <= 0.0f) {
        const float xd = __fsub_rn((float)px, x0), yd = __fsub_rn((float)py, y0);
        return sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
    }
    return P.max_range;
}

template <int AFTER, int K>
__global__ void __launch_bounds__(128) k_v19(Args a, Lean q)
{
    const unsigned i = blockIdx.x * 128u + threadIdx.x;
    const unsigned total = (unsigned)a.num_poses * a.num_beams;
    if (i >= total) return;
    float x0, y0, dx, dy;
    ray_setup_fan(a, q, i, x0, y0, dx, dy);
    const rl::FirstSample f0 = rl::first_sample(a.P, x0, y0);
    a.outs[i] = __fmul_rn(march_ray_v19<AFTER, K>(a.P, x0, y0, dx, dy, f0), a.P.w.scale);
}

// diag3: light-weight stamps only (lane 0 of each warp), product march untouched otherwise
template <int MODE>   // 0: product march; 1: v19<32,6>
__global__ void __launch_bounds__(128) k_diag3(Args a, Lean q, unsigned long long *t_first, unsigned long long *t_last)
{
    const unsigned i = blockIdx.x * 128u + threadIdx.x;
    const unsigned total = (unsigned)a.num_poses * a.num_beams;
    if (i >= total) return;
    const unsigned long long tstart = gtime();
    float x0, y0, dx, dy;
    ray_setup_fan(a, q, i, x0, y0, dx, dy);
    const rl::FirstSample f0 = rl::first_sample(a.P, x0, y0);
    uint32_t st = 0;
    float r;
    if (MODE == 0) r = rl::march_ray<false>(a.P, x0, y0, dx, dy, st, f0);
    else r = march_ray_v19<32, 6>(a.P, x0, y0, dx, dy, f0);
    a.outs[i] = __fmul_rn(r, a.P.w.scale);
    __syncwarp(__activemask());
    if ((threadIdx.x & 31) == 0) { t_first[i >> 5] = tstart; t_last[i >> 5] = gtime(); }
}

// ---------------------------------------------------------------- chain: what one march step costs when nothing else
// runs.  One warp, `lanes` active rays over a field of constant clearance `dval` (no obstacle, so every ray
// runs to max_range in steps of max(0.999*dval, 1) px); heading `th0` (+ lane * 0.004 rad).  MODE 0: the plain loop
// (no touches), 1: the product march (look-ahead touches after 32 steps), 2: v19<32,6>.
template <int MODE>
__global__ void k_chain(MarchParams P, float x0, float y0, float th0, int lanes, unsigned *out_steps, long long *out_cycles, float *sink)
{
    const int lane = threadIdx.x;
    if (lane >= lanes) return;
    float dx, dy;
    rl::glibc_sincosf(th0 + 0.004f * lane, &dy, &dx);
    const rl::FirstSample f0 = rl::first_sample(P, x0, y0);
    uint32_t st = 0;
    float r;
    const long long c0 = clock64();
    if (MODE == 0) {
        float t = 0.f; r = P.max_range;
        while (t < P.max_range) {
            const int px = __float2int_rz(fmaf(dx, t, x0)), py = __float2int_rz(fmaf(dy, t, y0));
            if ((unsigned)px >= (unsigned)P.rows || (unsigned)py >= (unsigned)P.cols) break;
            const float d = __ldg(P.dist + (px * P.cols + py));
            ++st;
            if (d <= 0.0f) { r = 0.f; break; }
            t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
        }
    } else if (MODE == 1) {
        r = rl::march_ray<true>(P, x0, y0, dx, dy, st, f0);
    } else {
        r = march_ray_v19<32, 6>(P, x0, y0, dx, dy, f0);
        st = 0;
    }
    const long long c1 = clock64();
    sink[lane] = r;
    out_steps[lane] = st;
    out_cycles[lane] = c1 - c0;
}

// ---------------------------------------------------------------- step: latency of one march step with every load an
// L1 hit (64x64 field of constant clearance, the same ray marched REP times in one launch; the last repetition
// is timed).  VAR 0: product step (cvt.rzi, integer bounds test, separate out-of-map and exit branches).
// VAR 1: truncation by a round-down add of 2^23 on max(x, 0) + float bounds test (no cvt in the chain).
// VAR 2: VAR 0 with the out-of-map test folded into the one exit branch (predicated load).
// VAR 3: VAR 1 + VAR 2.
template <int VAR>
__global__ void k_step(MarchParams P, float x0, float y0, float th0, int reps, long long *out_cycles, unsigned *out_steps, float *sink)
{
    float dx, dy;
    rl::glibc_sincosf(th0, &dy, &dx);
    float r = 0.f;
    long long c0 = 0;
    unsigned st = 0;
    const int K = -(0x4B000000) * (P.cols + 1);
    for (int rep = 0; rep < reps; ++rep) {
        if (rep == reps - 1) { c0 = clock64(); st = 0; }
        float t = 0.f;
        r = P.max_range;
        if (VAR == 0) {
            for (;;) {
                const int px = __float2int_rz(fmaf(dx, t, x0)), py = __float2int_rz(fmaf(dy, t, y0));
                if ((unsigned)px >= (unsigned)P.rows || (unsigned)py >= (unsigned)P.cols) break;
                const float d = __ldg(P.dist + (px * P.cols + py));
                ++st;
                t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
                if (d <= 0.0f || !(t < P.max_range)) { r = d; break; }
            }
        } else if (VAR == 1) {
            for (;;) {
                const float fx = fmaf(dx, t, x0), fy = fmaf(dy, t, y0);
                if (!(fx > -1.0f && fx < P.frows && fy > -1.0f && fy < P.fcols)) break;
                const int bx = __float_as_int(__fadd_rd(fmaxf(fx, 0.0f), 8388608.0f));
                const int by = __float_as_int(__fadd_rd(fmaxf(fy, 0.0f), 8388608.0f));
                const float d = __ldg(P.dist + (bx * P.cols + by + K));
                ++st;
                t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
                if (d <= 0.0f || !(t < P.max_range)) { r = d; break; }
            }
        } else if (VAR == 2) {
            for (;;) {
                const int px = __float2int_rz(fmaf(dx, t, x0)), py = __float2int_rz(fmaf(dy, t, y0));
                const bool inb = (unsigned)px < (unsigned)P.rows && (unsigned)py < (unsigned)P.cols;
                float d = 0.0f;
                if (inb) d = __ldg(P.dist + (px * P.cols + py));
                ++st;
                t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
                if (d <= 0.0f || !(t < P.max_range)) { r = inb ? d : P.max_range; break; }
            }
        } else {
            for (;;) {
                const float fx = fmaf(dx, t, x0), fy = fmaf(dy, t, y0);
                const bool inb = fx > -1.0f && fx < P.frows && fy > -1.0f && fy < P.fcols;
                const int bx = __float_as_int(__fadd_rd(fmaxf(fx, 0.0f), 8388608.0f));
                const int by = __float_as_int(__fadd_rd(fmaxf(fy, 0.0f), 8388608.0f));
                float d = 0.0f;
                if (inb) d = __ldg(P.dist + (bx * P.cols + by + K));
                ++st;
                t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
                if (d <= 0.0f || !(t < P.max_range)) { r = inb ? d : P.max_range; break; }
            }
        }
    }
    const long long c1 = clock64();
    sink[0] = r;
    out_cycles[0] = c1 - c0;
    out_steps[0] = st;
}

// ---------------------------------------------------------------- V20: shorter dependent chain per step.  The load is
// predicated on the bounds test instead of guarded by a branch (an out-of-map sample reads d = 0 and leaves
// through the one exit branch), and the loop is rotated: the next sample's cell is computed BEFORE the exit
// test of the current one, so the address arithmetic no longer waits for the branch to resolve.  Same samples,
// same arithmetic.  TOUCH: the product's look-ahead touches after 32 steps.
__device__ __forceinline__ int cvt_rz_pinned(float v)   // volatile: stays where it is written (before the exit branch)
{
    int r;
    asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float ldg_if(const float *p, bool pred)   // predicated load, 0 when pred is false
{
    float d;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\tmov.f32 %0, 0f00000000;\n\t@q ld.global.nc.f32 %0, [%1];\n\t}"
                 : "=f"(d) : "l"(p), "r"((int)pred));
    return d;
}
template <bool TOUCH>
__device__ __forceinline__ float march_ray_v20(const MarchParams &P, float x0, float y0, float dx, float dy,
                                               const rl::FirstSample &f0)
{
    if (!f0.inside || !(dx == dx) || !(dy == dy)) return P.max_range;
    if (f0.d <= 0.0f) {
        const float xd = __fsub_rn((float)f0.px, x0), yd = __fsub_rn((float)f0.py, y0);
        return sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
    }
    float t = fmaxf(__fmul_rn(f0.d, 0.999f), 1.0f);
    if (!(t < P.max_range)) return P.max_range;
    float th, d;                 // th: parameter of the sample whose clearance d is
    int it = 1;
    bool tail = false;
    int px = __float2int_rz(fmaf(dx, t, x0)), py = __float2int_rz(fmaf(dy, t, y0));
    bool inb = (unsigned)px < (unsigned)P.rows && (unsigned)py < (unsigned)P.cols;
    for (;;) {
        d = ldg_if(P.dist + (px * P.cols + py), inb);
        th = t;
        t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
        px = cvt_rz_pinned(fmaf(dx, t, x0));
        py = cvt_rz_pinned(fmaf(dy, t, y0));
        const bool was_in = inb;
        inb = (unsigned)px < (unsigned)P.rows && (unsigned)py < (unsigned)P.cols;
        if (d <= 0.0f || !(t < P.max_range)) { inb = was_in; break; }
        if (++it == 32) { tail = true; break; }
    }
    if (tail) {
        float j0 = 0.f, j1 = 0.f, j2 = 0.f, j3 = 0.f, keep = 0.f;
        const float adx = __fmul_rn(dx, 12.0f), ady = __fmul_rn(dy, 12.0f);
#define V20_STEP(J)                                                                                \
        {                                                                                          \
            d = ldg_if(P.dist + (px * P.cols + py), inb);                                          \
            if (TOUCH) {                                                                           \
                const int ax = __float2int_rz(__fadd_rn(fmaf(dx, t, x0), adx)), ay = __float2int_rz(__fadd_rn(fmaf(dy, t, y0), ady)); \
                keep = __fadd_rn(keep, J);                                                         \
                if ((unsigned)ax < (unsigned)P.rows && (unsigned)ay < (unsigned)P.cols)            \
                    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(J) : "l"(P.dist + (ax * P.cols + ay))); \
            }                                                                                      \
            th = t;                                                                                \
            t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));                                   \
            px = cvt_rz_pinned(fmaf(dx, t, x0));                                                   \
            py = cvt_rz_pinned(fmaf(dy, t, y0));                                                   \
            const bool was_in = inb;                                                               \
            inb = (unsigned)px < (unsigned)P.rows && (unsigned)py < (unsigned)P.cols;              \
            if (d <= 0.0f || !(t < P.max_range)) { inb = was_in; break; }                          \
        }
        for (;;) { V20_STEP(j0) V20_STEP(j1) V20_STEP(j2) V20_STEP(j3) }
#undef V20_STEP
        if (__fadd_rn(__fadd_rn(keep, j0), __fadd_rn(j1, __fadd_rn(j2, j3))) < 0.0f) return -1.0f;  // never true
    }
    // inb: whether the sample at th was inside the map; d its clearance (0 when outside)
    if (inb && d <= 0.0f) {
        const float xd = __fsub_rn((float)__float2int_rz(fmaf(dx, th, x0)), x0);
        const float yd = __fsub_rn((float)__float2int_rz(fmaf(dy, th, y0)), y0);
        return sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
    }
    return P.max_range;
}

template <bool TOUCH>
__global__ void __launch_bounds__(128) k_v20(Args a, Lean q)
{
    const unsigned i = blockIdx.x * 128u + threadIdx.x;
    const unsigned total = (unsigned)a.num_poses * a.num_beams;
    if (i >= total) return;
    float x0, y0, dx, dy;
    ray_setup_fan(a, q, i, x0, y0, dx, dy);
    const rl::FirstSample f0 = rl::first_sample(a.P, x0, y0);
    a.outs[i] = __fmul_rn(march_ray_v20<TOUCH>(a.P, x0, y0, dx, dy, f0), a.P.w.scale);
}

// ---------------------------------------------------------------- V21: the product kernel with the leaner trig
__global__ void __launch_bounds__(128) k_v21(Args a, Lean q)
{
    const unsigned i = blockIdx.x * 128u + threadIdx.x;
    const unsigned total = (unsigned)a.num_poses * a.num_beams;
    if (i >= total) return;
    const MarchParams &P = a.P;
    const unsigned k = __umulhi(i, q.magic) >> q.shift;
    const int j = i - k * a.num_beams;
    const float *p = a.poses + 3 * k;
    const float thw = __ldg(p + 2);
    const GridPose g = rl::world_to_grid(P.w, __ldg(p), __ldg(p + 1), thw);
    const float thg = __fadd_rn(-__fadd_rn(thw, fmaf((float)j, q.inc, -0.5f * a.fov)), P.w.rotation_const);
    const rl::FirstSample f0 = rl::first_sample(a.P, g.y, g.x);
    float dx, dy;
    rl::glibc_sincosf(thg, &dy, &dx);
    uint32_t st = 0;
    a.outs[i] = __fmul_rn(rl::march_ray<false>(a.P, g.y, g.x, dx, dy, st, f0), a.P.w.scale);
}

// ---------------------------------------------------------------- V22 (third session): unit-step cooperative tail.
// After HEAD lockstep steps (the product's loop) the rays of a warp that are still marching are served by the
// WHOLE warp, finished lanes included: G = 32, 16, 8 or 4 lanes per ray (1, 2, 3-4, 5+ rays; at most 8 rays per
// round, the lowest lanes first).  Lane k of a group samples at t + k -- the parameter the ray reaches after k
// steps of exactly 1 px, which is what a creeping ray takes (61 % of the steps beyond the 32nd on the bench map,
// in runs of 12): for t >= 16 the sequence t, t+1, (t+1)+1, ... equals fl(t + k) bit for bit (adds inside a
// binade are exact; the single rounding at a binade crossing commutes with adding integers), so no chain of
// dependent adds is needed.  A ballot finds the first sample of the group that does not continue with a unit
// step; that lane decides the ray's new state.  One memory round trip per run of unit steps instead of one per step.
template <int HEAD>
__device__ __forceinline__ bool march_head(const MarchParams &P, float x0, float y0, float dx, float dy,
                                           const rl::FirstSample &f0, float &t_out, float &res)
{
    res = P.max_range;
    if (!f0.inside || !(dx == dx) || !(dy == dy)) return false;
    if (f0.d <= 0.0f) {
        const float xd = __fsub_rn((float)f0.px, x0), yd = __fsub_rn((float)f0.py, y0);
        res = sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
        return false;
    }
    float t = fmaxf(__fmul_rn(f0.d, 0.999f), 1.0f);
    if (!(t < P.max_range)) return false;
    int px, py, it = 1;
    float d;
    for (;;) {
        px = __float2int_rz(fmaf(dx, t, x0));
        py = __float2int_rz(fmaf(dy, t, y0));
        if ((unsigned)px >= (unsigned)P.rows || (unsigned)py >= (unsigned)P.cols) return false;
        d = __ldg(P.dist + (px * P.cols + py));
        t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
        if (d <= 0.0f || !(t < P.max_range)) break;
        if (++it == HEAD) { t_out = t; return true; }
    }
    if (d <= 0.0f) {
        const float xd = __fsub_rn((float)px, x0), yd = __fsub_rn((float)py, y0);
        res = sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
    }
    return false;
}

// MAXG: largest number of rays served per round (1, 2, 4 or 8)
// PLAIN: while more than PLAIN rays of the warp are marching, every one of them takes an ordinary step of its own
// (lockstep, as in the product loop); the cooperative rounds start once PLAIN or fewer are left (0: always cooperate).
template <int MAXG, int PLAIN = 0>
__device__ __forceinline__ void coop_unit_tail(const MarchParams &P, float x0, float y0, float dx, float dy, float &t,
                                               float &res, bool running)
{
    const unsigned lane = threadIdx.x & 31u;
    unsigned m = __ballot_sync(0xffffffffu, running);
    while (m) {
        const int ma = __popc(m);
        if (PLAIN > 0 && ma > PLAIN) {
            if (running) {
                const int px = __float2int_rz(fmaf(dx, t, x0)), py = __float2int_rz(fmaf(dy, t, y0));
                if ((unsigned)px >= (unsigned)P.rows || (unsigned)py >= (unsigned)P.cols) { res = P.max_range; running = false; }
                else {
                    const float d = __ldg(P.dist + (px * P.cols + py));
                    const float nt = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
                    if (d <= 0.0f) {
                        const float xd = __fsub_rn((float)px, x0), yd = __fsub_rn((float)py, y0);
                        res = sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
                        running = false;
                    } else if (!(nt < P.max_range)) { res = P.max_range; running = false; }
                    else t = nt;
                }
            }
            m = __ballot_sync(0xffffffffu, running);
            continue;
        }
        int lg = ma <= 1 ? 0 : (ma == 2 ? 1 : (ma <= 4 ? 2 : 3));   // log2(groups)
        if ((1 << lg) > MAXG) lg = MAXG == 1 ? 0 : (MAXG == 2 ? 1 : (MAXG == 4 ? 2 : 3));
        const int ng = 1 << lg;                     // groups this round
        const int G = 32 >> lg;                     // lanes per group
        const int g = lane >> (5 - lg);             // my group (0 when lg == 0)
        const int k = lane & (G - 1);
        // owner of group g: the (g+1)-th lowest running lane
        unsigned mm = m;
        int owner = 0;
        bool has = false;
#pragma unroll
        for (int gg = 0; gg < MAXG; ++gg) {
            if (gg < ng && mm) {
                const int o = __ffs(mm) - 1;
                if (gg == g) { owner = o; has = true; }
                mm &= mm - 1;
            }
        }
        const float ox0 = __shfl_sync(0xffffffffu, x0, owner), oy0 = __shfl_sync(0xffffffffu, y0, owner);
        const float odx = __shfl_sync(0xffffffffu, dx, owner), ody = __shfl_sync(0xffffffffu, dy, owner);
        const float ot = __shfl_sync(0xffffffffu, t, owner);
        const float tk = __fadd_rn(ot, (float)k);               // == k sequential unit steps (ot >= 16)
        const bool live = tk < P.max_range;                    // k = 0: true by invariant
        const int px = __float2int_rz(fmaf(odx, tk, ox0)), py = __float2int_rz(fmaf(ody, tk, oy0));
        const bool inb = (unsigned)px < (unsigned)P.rows && (unsigned)py < (unsigned)P.cols;
        float d = 1.0f;
        if (has && live && inb) d = __ldg(P.dist + (px * P.cols + py));
        const float sk = fmaxf(__fmul_rn(d, 0.999f), 1.0f);
        const float nt = __fadd_rn(tk, sk);
        // what this sample decides if it is the first of its group that does not continue with a unit step
        int st = 0;                                            // 0: continue from nt, 1: finished with `r`
        float r = P.max_range;
        if (!inb) st = 1;
        else if (d <= 0.0f) {
            st = 1;
            const float xd = __fsub_rn((float)px, ox0), yd = __fsub_rn((float)py, oy0);
            r = sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
        } else if (!(nt < P.max_range)) st = 1;
        const bool unit_go = live && st == 0 && sk == 1.0f;      // the next lane's parameter is this ray's next one
        const unsigned bad = __ballot_sync(0xffffffffu, !unit_go);
        // the owner reads the verdict of the deciding lane of ITS group
        const int myrank = __popc(m & ((1u << lane) - 1u));
        const bool served = running && myrank < ng;
        const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << ((myrank & (ng - 1)) * G);
        const unsigned gb = bad & gmask;
        int src = served ? (gb ? (__ffs(gb) - 1) : ((myrank & (ng - 1)) * G + G - 1)) : (int)lane;
        const int st_o = __shfl_sync(0xffffffffu, st, src);
        const float nt_o = __shfl_sync(0xffffffffu, nt, src), r_o = __shfl_sync(0xffffffffu, r, src);
        const bool live_o = __shfl_sync(0xffffffffu, (int)live, src) != 0;
        if (served) {
            // a deciding lane that is not live cannot happen: the lane before it would have failed nt < max_range
            if (st_o == 0 && live_o) t = nt_o;
            else { res = r_o; running = false; }
        }
        m = __ballot_sync(0xffffffffu, running);
    }
}

template <int HEAD, int MAXG, int PLAIN = 0>
__global__ void __launch_bounds__(128, 16) k_v22(Args a, Lean q)
{
    const unsigned i = blockIdx.x * 128u + threadIdx.x;
    const unsigned total = (unsigned)a.num_poses * a.num_beams;
    const bool valid = i < total;
    float x0 = 0.f, y0 = 0.f, dx = 0.f, dy = 0.f, t = 0.f, res = 0.f;
    bool running = false;
    if (valid) {
        ray_setup_fan(a, q, i, x0, y0, dx, dy);
        const rl::FirstSample f0 = rl::first_sample(a.P, x0, y0);
        running = march_head<HEAD>(a.P, x0, y0, dx, dy, f0, t, res);
    }
    if (__any_sync(0xffffffffu, running)) coop_unit_tail<MAXG, PLAIN>(a.P, x0, y0, dx, dy, t, res, running);
    if (valid) a.outs[i] = __fmul_rn(res, a.P.w.scale);
}

// ---------------------------------------------------------------- harness
static std::vector<char> slurp(const std::string &path)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { printf("cannot open %s\n", path.c_str()); exit(1); }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> b(n);
    if (fread(b.data(), 1, n, f) != (size_t)n) exit(1);
    fclose(f);
    return b;
}

struct Runner {
    Args a;
    float *d_ref = nullptr;
    size_t n_rays;
    char *flush = nullptr;
    int reps;
    cudaEvent_t e0, e1;
    std::vector<float> h_ref, h_out;

    template <typename F> void run(const char *name, F launch, bool is_ref = false)
    {
        CK(cudaMemset(a.outs, 0xff, n_rays * 4));
        float best = 1e30f, sum = 0.f;
        for (int r = 0; r < reps + 2; ++r) {
            CK(cudaMemsetAsync(flush, r, 256u << 20));
            CK(cudaMemsetAsync(a.counter, 0, 4));
            CK(cudaEventRecord(e0));
            launch();
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r >= 2) { best = ms < best ? ms : best; sum += ms; }
        }
        CK(cudaMemcpy(h_out.data(), a.outs, n_rays * 4, cudaMemcpyDeviceToHost));
        size_t bad = 0;
        if (is_ref) h_ref = h_out;
        else for (size_t i = 0; i < n_rays; ++i) bad += memcmp(&h_out[i], &h_ref[i], 4) != 0;
        printf("%-44s mean %8.2f us  best %8.2f us  %7.2f Grays/s  mismatches %zu\n", name, sum / reps * 1e3,
               best * 1e3, n_rays / (sum / reps * 1e-3) / 1e9, bad);
        fflush(stdout);
    }
};

int main(int argc, char **argv)
{
    std::string dir = argc > 1 ? argv[1] : "/tmp/tune";
    int reps = argc > 2 ? atoi(argv[2]) : 20;
    auto meta = slurp(dir + "/meta.bin");   // int32 rows, cols, num_poses, num_beams; float max_range, fov; WorldFrame
    auto dist = slurp(dir + "/dist.bin");
    auto poses = slurp(dir + "/poses.bin");
    const int32_t *mi = (const int32_t *)meta.data();
    const float *mf = (const float *)(meta.data() + 16);
    Runner R;
    Args &a = R.a;
    a.P.rows = mi[0]; a.P.cols = mi[1]; a.num_poses = mi[2]; a.num_beams = mi[3];
    a.P.frows = (float)a.P.rows; a.P.fcols = (float)a.P.cols;
    a.P.max_range = mf[0]; a.fov = mf[1];
    memcpy(&a.P.w, mf + 2, sizeof(rl::WorldFrame));
    R.n_rays = (size_t)a.num_poses * a.num_beams;
    R.reps = reps;
    float *d_dist, *d_poses;
    CK(cudaMalloc(&d_dist, dist.size())); CK(cudaMemcpy(d_dist, dist.data(), dist.size(), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_poses, poses.size())); CK(cudaMemcpy(d_poses, poses.data(), poses.size(), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&a.outs, R.n_rays * 4));
    CK(cudaMalloc(&a.counter, 4));
    CK(cudaMalloc(&R.flush, 256u << 20));
    CK(cudaEventCreate(&R.e0)); CK(cudaEventCreate(&R.e1));
    a.P.dist = d_dist; a.poses = d_poses;
    R.h_out.resize(R.n_rays);
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    printf("map %dx%d, %d poses x %d beams, %d SMs, reps %d\n", a.P.rows, a.P.cols, a.num_poses, a.num_beams, sms, reps);

    const bool quick = argc > 3;
    const int groups = (a.num_beams + 31) / 32;
    auto base = [&](auto wpc_tag, int segs_target) {
        constexpr int WPC = decltype(wpc_tag)::value;
        int s = segs_target < 1 ? 1 : (segs_target > groups ? groups : segs_target);
        int gl = (groups + s - 1) / s, seg_len = gl * 32, segs = (groups + gl - 1) / gl;
        long warps = (long)a.num_poses * segs;
        k_base<WPC><<<(unsigned)((warps + WPC - 1) / WPC), WPC * 32>>>(a, segs, seg_len);
    };
    R.run("base wpc8 segs5 (product)", [&] { base(std::integral_constant<int, 8>{}, 5); }, true);
    if (!quick) {
    R.run("base wpc8 segs1 (warp per pose)", [&] { base(std::integral_constant<int, 8>{}, 1); });
    R.run("base wpc8 segs9", [&] { base(std::integral_constant<int, 8>{}, 9); });
    R.run("base wpc8 segs34 (32 beams/warp)", [&] { base(std::integral_constant<int, 8>{}, 34); });
    R.run("base wpc4 segs5", [&] { base(std::integral_constant<int, 4>{}, 5); });
    R.run("base wpc2 segs5", [&] { base(std::integral_constant<int, 2>{}, 5); });
    R.run("base wpc2 segs17", [&] { base(std::integral_constant<int, 2>{}, 17); });
    R.run("base wpc1 segs5", [&] { base(std::integral_constant<int, 1>{}, 5); });

#define PERSIST(ILP, WPC, BPS) R.run("persist ilp" #ILP " wpc" #WPC " blocks/SM " #BPS, [&] { \
        k_persist<ILP, WPC><<<sms * BPS, WPC * 32>>>(a, (a.num_beams + 32 * ILP - 1) / (32 * ILP)); })
    PERSIST(1, 8, 8); PERSIST(1, 4, 16); PERSIST(1, 8, 4);
    PERSIST(2, 8, 8); PERSIST(2, 8, 4); PERSIST(2, 4, 16); PERSIST(2, 8, 6);
    PERSIST(4, 8, 4); PERSIST(4, 8, 2); PERSIST(4, 8, 8); PERSIST(3, 8, 6);

#define FLAT(NB, WPC, BPS) R.run("flat nb" #NB " wpc" #WPC " blocks/SM " #BPS, [&] { \
        k_flat<NB, WPC><<<sms * BPS, WPC * 32>>>(a, (a.num_beams + 32 * NB - 1) / (32 * NB)); })
    FLAT(2, 8, 8); FLAT(4, 8, 8); FLAT(4, 8, 6); FLAT(6, 8, 6); FLAT(8, 8, 4); FLAT(8, 8, 8); FLAT(4, 4, 16);

#define CTAPOSE(ILP, WPC, BPS) R.run("cta-per-pose ilp" #ILP " wpc" #WPC " blocks/SM " #BPS, [&] { \
        k_cta_pose<ILP, WPC><<<sms * BPS, WPC * 32>>>(a, (a.num_beams + 32 * ILP - 1) / (32 * ILP)); })
    CTAPOSE(1, 8, 8); CTAPOSE(1, 4, 16); CTAPOSE(2, 4, 16); CTAPOSE(2, 8, 8); CTAPOSE(2, 8, 4); CTAPOSE(1, 16, 4);
    CTAPOSE(2, 16, 4); CTAPOSE(1, 2, 32); CTAPOSE(2, 2, 32);
    }

    // ---- lean variants ----
    {
        Lean q;
        // magic for unsigned division by num_beams valid for all 32-bit numerators: ceil(2^(32+s)/d)
        // exact for numerators < 2^31: k = umulhi(i, ceil(2^(31+s)/d)) >> (s-1), s = ceil(log2 d) >= 1
        int sft = 1; while ((1u << sft) < (unsigned)a.num_beams) ++sft;
        unsigned long long m = ((1ull << (31 + sft)) + a.num_beams - 1) / a.num_beams;
        bool ok = m <= 0xffffffffull && R.n_rays < (1ull << 31);
        sft -= 1;
        if (!ok) { printf("magic overflow, need 33-bit path\n"); }
        q.magic = (uint32_t)m; q.shift = sft; q.inc = a.fov / (float)a.num_beams;
        auto nb = [&](int bs, int rpl) { return (unsigned)((R.n_rays + (size_t)bs * rpl - 1) / ((size_t)bs * rpl)); };
        if (ok) {
        unsigned b3 = (unsigned)((R.n_rays + 127) / 128);
        if (argc > 3 && !strcmp(argv[3], "v17")) {
            R.run("product-like, checked loop only (reference)", [&] { k_product_like<<<b3, 128>>>(a, q); }, true);
            R.run("interior poses skip the bounds tests", [&] { k_product_like_t<true><<<b3, 128>>>(a, q); });
            R.run("checked loop only (templated twin)", [&] { k_product_like_t<false><<<b3, 128>>>(a, q); });
            CK(cudaFuncSetAttribute(k_product_like_t<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1));
            CK(cudaFuncSetAttribute(k_product_like_t<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1));
            R.run("interior + carveout max L1", [&] { k_product_like_t<true><<<b3, 128>>>(a, q); });
            R.run("checked + carveout max L1", [&] { k_product_like_t<false><<<b3, 128>>>(a, q); });
            R.run("product-like again", [&] { k_product_like<<<b3, 128>>>(a, q); });
            return 0;
        }
        if (argc > 3 && !strcmp(argv[3], "v18")) {
            R.run("product-like (reference of this table)", [&] { k_product_like<<<b3, 128>>>(a, q); }, true);
            auto mlp = [&](auto rtag) {
                constexpr int RR = decltype(rtag)::value;
                const unsigned cpp = (unsigned)((a.num_beams + 32 * RR - 1) / (32 * RR));
                int sf = 1; while ((1u << sf) < cpp) ++sf;
                const unsigned long long mg = ((1ull << (31 + sf)) + cpp - 1) / cpp;
                const unsigned long long warps = (unsigned long long)a.num_poses * cpp;
                k_mlp<RR><<<(unsigned)((warps + 3) / 4), 128>>>(a, q, cpp, (unsigned)mg, (unsigned)(sf - 1));
            };
            R.run("mlp R=1 (pose-aligned warps, control)", [&] { mlp(std::integral_constant<int, 1>{}); });
            R.run("mlp R=2", [&] { mlp(std::integral_constant<int, 2>{}); });
            R.run("mlp R=3", [&] { mlp(std::integral_constant<int, 3>{}); });
            R.run("mlp R=4", [&] { mlp(std::integral_constant<int, 4>{}); });
            R.run("product-like again", [&] { k_product_like<<<b3, 128>>>(a, q); });
            return 0;
        }
        if (argc > 3 && !strcmp(argv[3], "diag2")) {
            Diag2 D;
            const size_t nwarps = (R.n_rays + 31) / 32;
            CK(cudaMalloc(&D.t_first, nwarps * 8)); CK(cudaMalloc(&D.t_last, nwarps * 8)); CK(cudaMalloc(&D.max_steps, nwarps * 4));
            CK(cudaMalloc(&D.ray_steps, R.n_rays * 4)); CK(cudaMalloc(&D.ray_cycles, R.n_rays * 4));
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaMemset(D.t_last, 0, nwarps * 8)); CK(cudaMemset(D.max_steps, 0, nwarps * 4));
                CK(cudaMemsetAsync(R.flush, rep, 256u << 20));
                k_diag2<<<b3, 128>>>(a, q, D);
                CK(cudaDeviceSynchronize());
            }
            std::vector<unsigned long long> tf(nwarps), tl(nwarps);
            std::vector<unsigned> ms(nwarps), rs(R.n_rays), rc(R.n_rays);
            CK(cudaMemcpy(tf.data(), D.t_first, nwarps * 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(tl.data(), D.t_last, nwarps * 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(ms.data(), D.max_steps, nwarps * 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(rs.data(), D.ray_steps, R.n_rays * 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(rc.data(), D.ray_cycles, R.n_rays * 4, cudaMemcpyDeviceToHost));
            unsigned long long t0 = ~0ull, tmax = 0;
            for (auto v : tf) t0 = v < t0 ? v : t0;
            for (auto v : tl) tmax = v > tmax ? v : tmax;
            printf("DIAG2 kernel span %.1f us\n", (tmax - t0) / 1e3);
            for (unsigned long long tick = 0; tick <= (tmax - t0); tick += 4000) {
                size_t running = 0, started = 0;
                for (size_t w = 0; w < nwarps; ++w) { if (tf[w] - t0 <= tick) { ++started; if (tl[w] - t0 > tick) ++running; } }
                printf("DIAG2 t=%5.1f us  warps started %7zu  running %6zu\n", tick / 1e3, started, running);
            }
            // the 25 warps that finish last
            std::vector<size_t> idx(nwarps);
            for (size_t w = 0; w < nwarps; ++w) idx[w] = w;
            std::partial_sort(idx.begin(), idx.begin() + 25, idx.end(), [&](size_t x, size_t y) { return tl[x] > tl[y]; });
            for (int n = 0; n < 25; ++n) {
                const size_t w = idx[n];
                printf("DIAG2 last #%2d: warp %7zu start %6.1f us end %6.1f us  lived %5.1f us  max steps %u\n", n, w,
                       (tf[w] - t0) / 1e3, (tl[w] - t0) / 1e3, (tl[w] - tf[w]) / 1e3, ms[w]);
            }
            for (int mode = 0; mode < 2; ++mode) {
                for (int rep = 0; rep < 3; ++rep) {
                    CK(cudaMemsetAsync(R.flush, rep, 256u << 20));
                    if (mode == 0) k_diag3<0><<<b3, 128>>>(a, q, D.t_first, D.t_last); else k_diag3<1><<<b3, 128>>>(a, q, D.t_first, D.t_last);
                    CK(cudaDeviceSynchronize());
                }
                CK(cudaMemcpy(tf.data(), D.t_first, nwarps * 8, cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(tl.data(), D.t_last, nwarps * 8, cudaMemcpyDeviceToHost));
                t0 = ~0ull; tmax = 0;
                for (auto v : tf) t0 = v < t0 ? v : t0;
                for (auto v : tl) tmax = v > tmax ? v : tmax;
                printf("DIAG3 mode %d (0 = product march, 1 = predicted-sample touch) kernel span %.1f us\n", mode, (tmax - t0) / 1e3);
                for (unsigned long long tick = 0; tick <= (tmax - t0); tick += 2000) {
                    size_t running = 0, started = 0;
                    for (size_t w = 0; w < nwarps; ++w) { if (tf[w] - t0 <= tick) { ++started; if (tl[w] - t0 > tick) ++running; } }
                    if (tick % 8000 == 0 || started == nwarps) printf("DIAG3 t=%5.1f us  warps started %7zu  running %6zu\n", tick / 1e3, started, running);
                }
                for (size_t w = 0; w < nwarps; ++w) idx[w] = w;
                std::partial_sort(idx.begin(), idx.begin() + 25, idx.end(), [&](size_t x, size_t y) { return tl[x] > tl[y]; });
                for (int n = 0; n < 25; ++n) {
                    const size_t w = idx[n];
                    printf("DIAG3 last #%2d: warp %7zu start %6.1f us end %6.1f us  lived %5.1f us  max steps %u  (%.0f cycles/step)\n", n, w,
                           (tf[w] - t0) / 1e3, (tl[w] - t0) / 1e3, (tl[w] - tf[w]) / 1e3, ms[w], (tl[w] - tf[w]) * 1.965 / ms[w]);
                }
                for (int b = 0; b < 9; ++b) {
                    const unsigned edges2[10] = {1, 4, 8, 16, 32, 48, 64, 96, 128, 100000};
                    double life = 0; size_t n = 0;
                    for (size_t w = 0; w < nwarps; ++w) if (ms[w] >= edges2[b] && ms[w] < edges2[b + 1]) { life += (double)(tl[w] - tf[w]); ++n; }
                    if (n) printf("DIAG3 warps with max steps %u..%u: %zu warps, mean lifetime %.2f us\n", edges2[b], edges2[b + 1] - 1, n, life / n / 1e3);
                }
            }
            // cycles per step by ray length
            const unsigned edges[10] = {1, 4, 8, 16, 32, 48, 64, 96, 128, 100000};
            for (int b = 0; b < 9; ++b) {
                double cyc = 0, stp = 0; size_t n = 0;
                for (size_t i = 0; i < R.n_rays; ++i) if (rs[i] >= edges[b] && rs[i] < edges[b + 1]) { cyc += rc[i]; stp += rs[i]; ++n; }
                if (n) printf("DIAG2 rays with %u..%u steps: %zu rays, %.0f cycles/ray, %.0f cycles/step\n", edges[b], edges[b + 1] - 1, n, cyc / n, cyc / stp);
            }
            // warp lifetime by max steps
            for (int b = 0; b < 9; ++b) {
                double life = 0; size_t n = 0;
                for (size_t w = 0; w < nwarps; ++w) if (ms[w] >= edges[b] && ms[w] < edges[b + 1]) { life += (double)(tl[w] - tf[w]); ++n; }
                if (n) printf("DIAG2 warps with max steps %u..%u: %zu warps, mean lifetime %.2f us\n", edges[b], edges[b + 1] - 1, n, life / n / 1e3);
            }
            return 0;
        }
        if (argc > 3 && !strcmp(argv[3], "v19")) {
            R.run("product-like (reference of this table)", [&] { k_product_like<<<b3, 128>>>(a, q); }, true);
#define V19(AFTER, K) R.run("touch the predicted sample: after" #AFTER " K" #K, [&] { k_v19<AFTER, K><<<b3, 128>>>(a, q); });
            V19(32, 2) V19(32, 3) V19(32, 4) V19(32, 5) V19(32, 6) V19(32, 8) V19(32, 12)
            V19(24, 3) V19(24, 4) V19(24, 6) V19(16, 3) V19(16, 4) V19(16, 6) V19(12, 4) V19(8, 4) V19(48, 4)
            R.run("product-like again", [&] { k_product_like<<<b3, 128>>>(a, q); });
            return 0;
        }
        if (argc > 3 && !strcmp(argv[3], "chain")) {
            const int n = 2049;
            float *cf; unsigned *osteps; long long *ocyc; float *sink;
            CK(cudaMalloc(&cf, (size_t)n * n * 4)); CK(cudaMalloc(&osteps, 128)); CK(cudaMalloc(&ocyc, 256)); CK(cudaMalloc(&sink, 128));
            MarchParams CP = a.P; CP.dist = cf; CP.rows = n; CP.cols = n; CP.frows = (float)n; CP.fcols = (float)n; CP.max_range = 300.f;
            for (float dval : {1.0f, 2.0f, 3.0f, 5.0f}) {
                std::vector<float> h((size_t)n * n, dval);
                CK(cudaMemcpy(cf, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
                const unsigned nsteps = 1 + (unsigned)(300.0f / fmaxf(dval * 0.999f, 1.0f));
                for (int dir = 0; dir < 3; ++dir) {
                    const float th0 = dir == 0 ? 0.02f : dir == 1 ? 1.55f : 0.8f;   // dx ~ 1: across rows; dy ~ 1: along a row; diagonal
                    for (int lanes : {1, 4, 32}) {
                        for (int mode = 0; mode < 3; ++mode) {
                            long long cyc[32]; unsigned stp[32];
                            for (int rep = 0; rep < 2; ++rep) {   // second repetition: L2 warm (L1 is invalidated per launch)
                                if (mode == 0) k_chain<0><<<1, 32>>>(CP, 1000.5f, 1000.5f, th0, lanes, osteps, ocyc, sink);
                                else if (mode == 1) k_chain<1><<<1, 32>>>(CP, 1000.5f, 1000.5f, th0, lanes, osteps, ocyc, sink);
                                else k_chain<2><<<1, 32>>>(CP, 1000.5f, 1000.5f, th0, lanes, osteps, ocyc, sink);
                                CK(cudaDeviceSynchronize());
                            }
                            CK(cudaMemcpy(cyc, ocyc, 256, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(stp, osteps, 128, cudaMemcpyDeviceToHost));
                            printf("CHAIN d=%.0f dir=%s lanes=%2d mode=%d (%s): %5.1f cycles/step (%u steps)\n", dval,
                                   dir == 0 ? "across-rows" : dir == 1 ? "along-row " : "diagonal   ", lanes, mode,
                                   mode == 0 ? "plain loop   " : mode == 1 ? "product march" : "predicted K6 ", (double)cyc[0] / nsteps, nsteps);
                        }
                    }
                }
            }
            return 0;
        }
        if (argc > 3 && !strcmp(argv[3], "step")) {
            const int n = 64;
            float *cf; unsigned *osteps; long long *ocyc; float *sink;
            CK(cudaMalloc(&cf, (size_t)n * n * 4)); CK(cudaMalloc(&osteps, 128)); CK(cudaMalloc(&ocyc, 256)); CK(cudaMalloc(&sink, 128));
            MarchParams CP = a.P; CP.dist = cf; CP.rows = n; CP.cols = n; CP.frows = (float)n; CP.fcols = (float)n; CP.max_range = 50.f;
            std::vector<float> h((size_t)n * n, 1.0f);
            CK(cudaMemcpy(cf, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
            for (int dir = 0; dir < 2; ++dir) {
                const float th0 = dir == 0 ? 0.02f : 0.8f;
                for (int var = 0; var < 4; ++var) {
                    long long cyc; unsigned stp;
                    for (int rep = 0; rep < 2; ++rep) {
                        if (var == 0) k_step<0><<<1, 1>>>(CP, 5.5f, 5.5f, th0, 8, ocyc, osteps, sink);
                        else if (var == 1) k_step<1><<<1, 1>>>(CP, 5.5f, 5.5f, th0, 8, ocyc, osteps, sink);
                        else if (var == 2) k_step<2><<<1, 1>>>(CP, 5.5f, 5.5f, th0, 8, ocyc, osteps, sink);
                        else k_step<3><<<1, 1>>>(CP, 5.5f, 5.5f, th0, 8, ocyc, osteps, sink);
                        CK(cudaDeviceSynchronize());
                    }
                    CK(cudaMemcpy(&cyc, ocyc, 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&stp, osteps, 4, cudaMemcpyDeviceToHost));
                    printf("STEP dir=%d var=%d: %6.1f cycles/step (%u steps, all L1 hits)\n", dir, var, (double)cyc / stp, stp);
                }
            }
            return 0;
        }
        if (argc > 3 && !strcmp(argv[3], "v20")) {
            R.run("product-like (reference of this table)", [&] { k_product_like<<<b3, 128>>>(a, q); }, true);
            R.run("v20 rotated loop, predicated load, touches", [&] { k_v20<true><<<b3, 128>>>(a, q); });
            R.run("v20 rotated loop, predicated load, no touch", [&] { k_v20<false><<<b3, 128>>>(a, q); });
            R.run("product-like again", [&] { k_product_like<<<b3, 128>>>(a, q); });
            R.run("v20 touches again", [&] { k_v20<true><<<b3, 128>>>(a, q); });
            return 0;
        }
        if (argc > 3 && !strcmp(argv[3], "v21")) {
            R.run("product-like (reference of this table)", [&] { k_product_like<<<b3, 128>>>(a, q); }, true);
            R.run("v21 lean trig", [&] { k_v21<<<b3, 128>>>(a, q); });
            R.run("product-like again", [&] { k_product_like<<<b3, 128>>>(a, q); });
            R.run("v21 lean trig again", [&] { k_v21<<<b3, 128>>>(a, q); });
            return 0;
        }
        if (argc > 3 && !strcmp(argv[3], "v22")) {
            R.run("product-like (reference of this table)", [&] { k_product_like<<<b3, 128>>>(a, q); }, true);
#define V22(HEAD, MAXG) R.run("v22 unit-step coop tail: head" #HEAD " maxgroups" #MAXG, [&] { k_v22<HEAD, MAXG><<<b3, 128>>>(a, q); });
            V22(32, 8) V22(32, 4) V22(32, 2) V22(32, 1) V22(24, 8) V22(24, 4) V22(48, 8) V22(48, 4) V22(20, 8) V22(64, 8)
#define V22P(HEAD, MAXG, PLAIN) R.run("v22 head" #HEAD " maxgroups" #MAXG ", plain steps while > " #PLAIN " rays", [&] { k_v22<HEAD, MAXG, PLAIN><<<b3, 128>>>(a, q); });
            V22P(32, 4, 4) V22P(32, 2, 2) V22P(32, 1, 1) V22P(20, 4, 4) V22P(20, 2, 2) V22P(20, 1, 1) V22P(24, 8, 8)
            R.run("product-like again", [&] { k_product_like<<<b3, 128>>>(a, q); });
            return 0;
        }
        if (argc > 3 && !strcmp(argv[3], "v16")) {
            const unsigned ngroups = (unsigned)((R.n_rays + 31) / 32);
            R.run("product-like (reference of this table)", [&] { k_product_like<<<b3, 128>>>(a, q); }, true);
            R.run("v16 plain (restructured, no new idea)", [&] { k_v16<false, 0><<<b3, 128>>>(a, q); });
            R.run("v16 safe", [&] { k_v16<true, 0><<<b3, 128>>>(a, q); });
            R.run("v16 coop after32", [&] { k_v16<false, 32><<<b3, 128>>>(a, q); });
            R.run("v16 coop after24", [&] { k_v16<false, 24><<<b3, 128>>>(a, q); });
            R.run("v16 coop after16", [&] { k_v16<false, 16><<<b3, 128>>>(a, q); });
            R.run("v16 coop after12", [&] { k_v16<false, 12><<<b3, 128>>>(a, q); });
            R.run("v16 safe + coop after24", [&] { k_v16<true, 24><<<b3, 128>>>(a, q); });
            R.run("v16 safe + coop after16", [&] { k_v16<true, 16><<<b3, 128>>>(a, q); });
#define V16P(SAFE, COOP, CH, PCT) { char nm[96]; snprintf(nm, sizeof nm, "v16 persistent safe%d coop%d chunk%d for %d%%", SAFE, COOP, CH, PCT); \
            const unsigned cg = (unsigned)((unsigned long long)ngroups * PCT / 100) / CH * CH; \
            R.run(nm, [&] { k_v16p<SAFE, COOP, CH><<<sms * 16, 128>>>(a, q, ngroups, cg); }); }
            V16P(false, 0, 4, 90) V16P(false, 0, 4, 97) V16P(false, 0, 2, 95) V16P(false, 0, 8, 90) V16P(false, 0, 1, 0)
            V16P(true, 0, 4, 90) V16P(true, 24, 4, 90) V16P(true, 24, 4, 97) V16P(true, 16, 2, 95)
            R.run("product-like again", [&] { k_product_like<<<b3, 128>>>(a, q); });
            return 0;
        }
        R.run("ray3 opt0", [&] { k_ray3<0><<<b3, 128>>>(a, q); });
        R.run("ray3 opt1 magic", [&] { k_ray3<1><<<b3, 128>>>(a, q); });
        R.run("ray3 opt2 hit-after", [&] { k_ray3<2><<<b3, 128>>>(a, q); });
        R.run("ray3 opt4 int-bounds", [&] { k_ray3<4><<<b3, 128>>>(a, q); });
        R.run("ray3 opt3", [&] { k_ray3<3><<<b3, 128>>>(a, q); });
        R.run("ray3 opt7", [&] { k_ray3<7><<<b3, 128>>>(a, q); });
        R.run("ray3 opt5", [&] { k_ray3<5><<<b3, 128>>>(a, q); });
        if (argc > 5) {
            const int rows_ = a.P.rows, cols_ = a.P.cols;
            const float *hd_ = (const float *)dist.data();
            std::vector<float> tr_((size_t)rows_ * cols_);
            for (int px = 0; px < rows_; ++px)
                for (int py = 0; py < cols_; ++py) tr_[(size_t)py * rows_ + px] = hd_[(size_t)px * cols_ + py];
            float *dT_;
            CK(cudaMalloc(&dT_, tr_.size() * 4));
            CK(cudaMemcpy(dT_, tr_.data(), tr_.size() * 4, cudaMemcpyHostToDevice));
            for (int dual = 0; dual < 2; ++dual) {
            printf("DIAG ===== dual layout %d\n", dual);
            Diag D;
            const size_t nwarps = (R.n_rays + 31) / 32;
            CK(cudaMalloc(&D.t_first, nwarps * 8)); CK(cudaMalloc(&D.t_last, nwarps * 8));
            CK(cudaMalloc(&D.long_cycles, 65536 * 8)); CK(cudaMalloc(&D.long_steps, 65536 * 4));
            CK(cudaMalloc(&D.n_long, 4)); CK(cudaMalloc(&D.t0, 8));
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaMemset(D.n_long, 0, 4)); CK(cudaMemset(D.t0, 0xff, 8));
                CK(cudaMemsetAsync(R.flush, rep, 256u << 20));
                if (dual) k_diag<1><<<b3, 128>>>(a, q, D, dT_); else k_diag<0><<<b3, 128>>>(a, q, D, dT_);
                CK(cudaDeviceSynchronize());
            }
            std::vector<unsigned long long> tf(nwarps), tl(nwarps), lc(65536);
            std::vector<unsigned> ls(65536);
            unsigned nl; unsigned long long t0;
            CK(cudaMemcpy(tf.data(), D.t_first, nwarps * 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(tl.data(), D.t_last, nwarps * 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(lc.data(), D.long_cycles, 65536 * 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(ls.data(), D.long_steps, 65536 * 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(&nl, D.n_long, 4, cudaMemcpyDeviceToHost));
            t0 = ~0ull; for (auto v : tf) t0 = v < t0 ? v : t0;
            // timeline: warps running at each 4 us tick
            unsigned long long tmax = 0; for (auto v : tl) tmax = v > tmax ? v : tmax;
            printf("DIAG kernel span %.1f us, long rays (>48 steps): %u\n", (tmax - t0) / 1e3, nl);
            for (unsigned long long tick = 0; tick <= (tmax - t0); tick += 4000) {
                size_t running = 0, started = 0;
                for (size_t w = 0; w < nwarps; ++w) { if (tf[w] - t0 <= tick) { ++started; if (tl[w] - t0 > tick) ++running; } }
                printf("DIAG t=%5.1f us  warps started %7zu  running %6zu\n", tick / 1e3, started, running);
            }
            // per-step cycles of long rays, bucketed by steps
            double cs[6] = {0}; unsigned cn[6] = {0}; const unsigned edges[7] = {48, 64, 96, 128, 192, 256, 100000};
            unsigned long long worst = 0; unsigned worst_steps = 0;
            for (unsigned k = 0; k < (nl < 65536 ? nl : 65536); ++k) {
                for (int b = 0; b < 6; ++b) if (ls[k] > edges[b] && ls[k] <= edges[b + 1]) { cs[b] += (double)lc[k] / ls[k]; ++cn[b]; }
                if (lc[k] > worst) { worst = lc[k]; worst_steps = ls[k]; }
            }
            for (int b = 0; b < 6; ++b) if (cn[b]) printf("DIAG rays with %u..%u steps: %u rays, %.0f cycles/step\n", edges[b], edges[b + 1], cn[b], cs[b] / cn[b]);
            printf("DIAG slowest ray: %u steps, %llu cycles (%.1f us at 1.965 GHz)\n", worst_steps, worst, worst / 1965.0);
            }
        }
        {
            const int rows = a.P.rows, cols = a.P.cols;
            const float *hd = (const float *)dist.data();
            std::vector<float> tr((size_t)rows * cols);
            for (int px = 0; px < rows; ++px)
                for (int py = 0; py < cols; ++py) tr[(size_t)py * rows + px] = hd[(size_t)px * cols + py];
            float *dT;
            CK(cudaMalloc(&dT, tr.size() * 4));
            CK(cudaMemcpy(dT, tr.data(), tr.size() * 4, cudaMemcpyHostToDevice));
            R.run("dual layout (row/col-major by ray direction)", [&] { k_dual<0><<<b3, 128>>>(a, q, dT); });
            R.run("always transposed (control)", [&] { k_dual<1><<<b3, 128>>>(a, q, dT); });
        }
        if (argc > 4) {
            TailQ Q;
            Q.capacity = (unsigned)(R.n_rays / 4);
            CK(cudaMalloc(&Q.count, 4));
            CK(cudaMalloc(&Q.entries, (size_t)Q.capacity * 8));
            for (int cap : {6, 8, 10, 12, 16, 20, 24, 32, 48}) {
                char nm[64]; snprintf(nm, sizeof nm, "two-phase cap %d", cap);
                for (int p2blocks : {sms * 4, sms * 16}) {
                    char nm2[96]; snprintf(nm2, sizeof nm2, "%s, phase2 grid %d", nm, p2blocks);
                    R.run(nm2, [&] {
                        cudaMemsetAsync(Q.count, 0, 4);
                        k_phase1<<<b3, 128>>>(a, q, Q, cap);
                        k_phase2<<<p2blocks, 128>>>(a, q, Q);
                    });
                }
            }
            unsigned cnt; CK(cudaMemcpy(&cnt, Q.count, 4, cudaMemcpyDeviceToHost));
            printf("queue entries at last cap: %u\n", cnt);
        }
        R.run("smem tile 32x32 per pose (1024-thread CTA)", [&] { k_tile<32><<<a.num_poses, 1024>>>(a, q); });
        R.run("smem tile 64x64 per pose (1024-thread CTA)", [&] { k_tile<64><<<a.num_poses, 1024>>>(a, q); });
        R.run("smem tile 96x96 per pose (1024-thread CTA)", [&] { k_tile<96><<<a.num_poses, 1024>>>(a, q); });
        R.run("value-predicted tail (registers) after32", [&] { k_spec2<32><<<b3, 128>>>(a, q); });
        R.run("value-predicted tail (registers) after16", [&] { k_spec2<16><<<b3, 128>>>(a, q); });
        R.run("value-predicted tail (registers) after8", [&] { k_spec2<8><<<b3, 128>>>(a, q); });
        R.run("value-predicted tail (registers) after24", [&] { k_spec2<24><<<b3, 128>>>(a, q); });
        R.run("value-predicted tail after32 depth4", [&] { k_spec<32, 4><<<b3, 128>>>(a, q); });
        R.run("value-predicted tail after32 depth3", [&] { k_spec<32, 3><<<b3, 128>>>(a, q); });
        R.run("value-predicted tail after32 depth6", [&] { k_spec<32, 6><<<b3, 128>>>(a, q); });
        R.run("value-predicted tail after16 depth4", [&] { k_spec<16, 4><<<b3, 128>>>(a, q); });
        R.run("value-predicted tail after24 depth4", [&] { k_spec<24, 4><<<b3, 128>>>(a, q); });
        R.run("value-predicted tail after48 depth4", [&] { k_spec<48, 4><<<b3, 128>>>(a, q); });
        R.run("merged exits (no tail mode)", [&] { k_merged<<<b3, 128>>>(a, q); });
        R.run("product-like (tail mode, first sample)", [&] { k_product_like<<<b3, 128>>>(a, q); });
        {
            const unsigned ngroups = (unsigned)((R.n_rays + 31) / 32);
            for (int pct : {100, 90, 80, 60, 0}) {
                for (int bps : {16, 12}) {
                    char nm[96]; snprintf(nm, sizeof nm, "persistent static %d%% then atomic, %d CTAs/SM", pct, bps);
                    R.run(nm, [&] { k_pstatic<<<sms * bps, 128>>>(a, q, (unsigned)((unsigned long long)ngroups * pct / 100)); });
                }
            }
        }
        R.run("tail mode NO TOUCH after32 (control)", [&] { k_tail<12, 32, false><<<b3, 128>>>(a, q); });
        R.run("tail mode NO TOUCH after1 (control)", [&] { k_tail<12, 1, false><<<b3, 128>>>(a, q); });
        R.run("tail mode ahead12 after1", [&] { k_tail<12, 1><<<b3, 128>>>(a, q); });
        R.run("tail mode ahead16 after32", [&] { k_tail<16, 32><<<b3, 128>>>(a, q); });
        R.run("tail mode ahead24 after32", [&] { k_tail<24, 32><<<b3, 128>>>(a, q); });
        R.run("tail mode ahead16 after20", [&] { k_tail<16, 20><<<b3, 128>>>(a, q); });
        R.run("tail mode ahead12 after12", [&] { k_tail<12, 12><<<b3, 128>>>(a, q); });
        R.run("tail mode ahead12 after8", [&] { k_tail<12, 8><<<b3, 128>>>(a, q); });
        R.run("tail mode ahead32 after32", [&] { k_tail<32, 32><<<b3, 128>>>(a, q); });
        R.run("tail mode ahead4 after32", [&] { k_tail<4, 32><<<b3, 128>>>(a, q); });
        R.run("tail mode ahead8 after32", [&] { k_tail<8, 32><<<b3, 128>>>(a, q); });
        R.run("tail mode ahead6 after16", [&] { k_tail<6, 16><<<b3, 128>>>(a, q); });
        R.run("tail mode ahead8 after24", [&] { k_tail<8, 24><<<b3, 128>>>(a, q); });
        R.run("tail mode ahead12 after32", [&] { k_tail<12, 32><<<b3, 128>>>(a, q); });
        R.run("tail mode ahead8 after48", [&] { k_tail<8, 48><<<b3, 128>>>(a, q); });
        R.run("touch prefetch.L1 ahead4 near3", [&] { k_touch<4, 0><<<b3, 128>>>(a, q, 3.0f); });
        R.run("touch ld ahead4 near3", [&] { k_touch<4, 1><<<b3, 128>>>(a, q, 3.0f); });
        R.run("touch ld ahead3 near2", [&] { k_touch<3, 1><<<b3, 128>>>(a, q, 2.0f); });
        R.run("touch ld ahead6 near3", [&] { k_touch<6, 1><<<b3, 128>>>(a, q, 3.0f); });
        R.run("touch ld ahead4 near1.5", [&] { k_touch<4, 1><<<b3, 128>>>(a, q, 1.5f); });
        R.run("touch ld ahead8 near4", [&] { k_touch<8, 1><<<b3, 128>>>(a, q, 4.0f); });
        R.run("touch prefetch.L1 ahead8 near4", [&] { k_touch<8, 0><<<b3, 128>>>(a, q, 4.0f); });
        R.run("lean bs128 rpl1", [&] { k_lean<128, 1><<<nb(128, 1), 128>>>(a, q, 1 << 30); });
        R.run("lean bs256 rpl1", [&] { k_lean<256, 1><<<nb(256, 1), 256>>>(a, q, 1 << 30); });
        R.run("lean bs128 rpl2", [&] { k_lean<128, 2><<<nb(128, 2), 128>>>(a, q, 1 << 30); });
        R.run("lean bs64 rpl2", [&] { k_lean<64, 2><<<nb(64, 2), 64>>>(a, q, 1 << 30); });
        R.run("lean bs128 rpl3", [&] { k_lean<128, 3><<<nb(128, 3), 128>>>(a, q, 1 << 30); });
        R.run("lean bs128 rpl4", [&] { k_lean<128, 4><<<nb(128, 4), 128>>>(a, q, 1 << 30); });
        R.run("lean bs64 rpl4", [&] { k_lean<64, 4><<<nb(64, 4), 64>>>(a, q, 1 << 30); });
        R.run("lean bs128 rpl1 cap32 (timing only)", [&] { k_lean<128, 1><<<nb(128, 1), 128>>>(a, q, 32); });
        R.run("lean bs128 rpl2 cap32 (timing only)", [&] { k_lean<128, 2><<<nb(128, 2), 128>>>(a, q, 32); });
        }
    }
    // ---- alternative layouts ----
    {
        const int rows = a.P.rows, cols = a.P.cols;
        const float *hd = (const float *)dist.data();
        auto build = [&](int th, int tw, bool line, int *tpr) {
            int TH = line ? 4 : th, TW = line ? 8 : tw;
            int trows = (rows + TH - 1) / TH, tcols = (cols + TW - 1) / TW;
            *tpr = tcols;
            std::vector<float> t((size_t)trows * tcols * TH * TW, 0.f);
            for (int px = 0; px < rows; ++px)
                for (int py = 0; py < cols; ++py) {
                    size_t idx;
                    if (line) {
                        size_t ln = (size_t)(px >> 2) * tcols + (py >> 3);
                        int sector = ((px >> 1) & 1) * 2 + ((py >> 2) & 1);
                        idx = (ln << 5) | (sector << 3) | ((px & 1) << 2) | (py & 3);
                    } else {
                        int lh = th == 2 ? 1 : 2, lw = tw == 4 ? 2 : 1;
                        idx = ((((size_t)(px >> lh)) * tcols + (py >> lw)) << 3) | ((px & (th - 1)) << lw) | (py & (tw - 1));
                    }
                    t[idx] = hd[(size_t)px * cols + py];
                }
            float *d;
            CK(cudaMalloc(&d, t.size() * 4));
            CK(cudaMemcpy(d, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
            return d;
        };
        unsigned blocks = (unsigned)((R.n_rays + 255) / 256);
        LayoutArgs la{};
        R.run("ray rowmajor (new product)", [&] { k_ray<ROWMAJOR><<<blocks, 256>>>(a, la); });
        R.run("ray rowmajor bs64", [&] { k_ray<ROWMAJOR, 64><<<(unsigned)((R.n_rays + 63) / 64), 64>>>(a, la); });
        R.run("ray rowmajor bs128", [&] { k_ray<ROWMAJOR, 128><<<(unsigned)((R.n_rays + 127) / 128), 128>>>(a, la); });
        R.run("ray rowmajor bs512", [&] { k_ray<ROWMAJOR, 512><<<(unsigned)((R.n_rays + 511) / 512), 512>>>(a, la); });
        R.run("ray rowmajor bs32", [&] { k_ray<ROWMAJOR, 32><<<(unsigned)((R.n_rays + 31) / 32), 32>>>(a, la); });
        R.run("ray rowmajor cap 8 (timing only)", [&] { k_ray<ROWMAJOR><<<blocks, 256>>>(a, la, 8); });
        R.run("ray rowmajor cap 16 (timing only)", [&] { k_ray<ROWMAJOR><<<blocks, 256>>>(a, la, 16); });
        R.run("ray rowmajor cap 32 (timing only)", [&] { k_ray<ROWMAJOR><<<blocks, 256>>>(a, la, 32); });
        R.run("ray rowmajor cap 64 (timing only)", [&] { k_ray<ROWMAJOR><<<blocks, 256>>>(a, la, 64); });
        R.run("ray rowmajor cap 128 (timing only)", [&] { k_ray<ROWMAJOR><<<blocks, 256>>>(a, la, 128); });
        R.run("ray rowmajor bs64 cap 32 (timing only)", [&] { k_ray<ROWMAJOR, 64><<<(unsigned)((R.n_rays + 63) / 64), 64>>>(a, la, 32); });
        la.tiled = build(2, 4, false, &la.tiles_per_row);
        R.run("ray tiled 2x4 sectors", [&] { k_ray<TILE_2x4><<<blocks, 256>>>(a, la); });
        la.tiled = build(4, 2, false, &la.tiles_per_row);
        R.run("ray tiled 4x2 sectors", [&] { k_ray<TILE_4x2><<<blocks, 256>>>(a, la); });
        la.tiled = build(0, 0, true, &la.tiles_per_row);
        R.run("ray tiled line 4x8 of 2x4 sectors", [&] { k_ray<TILE_LINE><<<blocks, 256>>>(a, la); });
        // texture (block-linear cudaArray, point sampling)
        {
            cudaArray_t arr;
            cudaChannelFormatDesc cd = cudaCreateChannelDesc<float>();
            CK(cudaMallocArray(&arr, &cd, cols, rows));
            CK(cudaMemcpy2DToArray(arr, 0, 0, hd, (size_t)cols * 4, (size_t)cols * 4, rows, cudaMemcpyHostToDevice));
            cudaResourceDesc rd{}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
            cudaTextureDesc td{}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
            td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
            CK(cudaCreateTextureObject(&la.tex, &rd, &td, nullptr));
            R.run("ray texture (cudaArray, point)", [&] { k_ray<TEX><<<blocks, 256>>>(a, la); });
        }
        // u16 d^2, 4x4 per sector
        {
            int trows = (rows + 3) / 4, tcols = (cols + 3) / 4;
            la.tiles_per_row = tcols;
            std::vector<unsigned short> t((size_t)trows * tcols * 16, 0);
            float mx = 0;
            for (int px = 0; px < rows; ++px)
                for (int py = 0; py < cols; ++py) {
                    float d = hd[(size_t)px * cols + py];
                    mx = d > mx ? d : mx;
                    long d2 = lroundf(d * d);
                    size_t idx = ((((size_t)(px >> 2)) * tcols + (py >> 2)) << 4) | ((px & 3) << 2) | (py & 3);
                    t[idx] = (unsigned short)(d2 > 65535 ? 65535 : d2);
                }
            unsigned short *d;
            CK(cudaMalloc(&d, t.size() * 2));
            CK(cudaMemcpy(d, t.data(), t.size() * 2, cudaMemcpyHostToDevice));
            la.d2u16 = d;
            printf("max dist %.2f px\n", mx);
            R.run("ray u16 d^2 tiled 4x4 + sqrt", [&] { k_ray<U16_4x4><<<blocks, 256>>>(a, la); });
        }
    }
    return 0;
}
