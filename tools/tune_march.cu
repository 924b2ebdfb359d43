// tools/tune_march.cu -- development harness (not shipped, not on the product path): times
// candidate fan-march kernel structures on the same distance field and poses, checks every
// candidate bit-for-bit against the straightforward one, and prints a table.
//   python tools/tune_prep.py /tmp/tune && tools/tune_march /tmp/tune [reps]
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
