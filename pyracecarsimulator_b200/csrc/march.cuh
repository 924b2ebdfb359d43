// march.cuh -- device-side restatement of RangeMethod::numpy_calc_range + RayMarching::calc_range
// (external range_libc), shared by every march kernel.  The fp32 operation order
// is the one oracle/rangelib_oracle.c fixes: this file is compiled with -fmad=false, so only
// the fmaf() written here are fused.
#pragma once
#include "common.h"

namespace rl {

struct GridPose { float x, y, theta; };  // x along msg columns, y along msg rows (before the swap)

__device__ __forceinline__ GridPose world_to_grid(const WorldFrame &w, float xw, float yw, float thw)
{
    const float x = __fmul_rn(__fsub_rn(xw, w.origin_x), w.inv_scale);
    const float y = __fmul_rn(__fsub_rn(yw, w.origin_y), w.inv_scale);
    GridPose g;
    g.x = fmaf(w.cos_angle, x, -__fmul_rn(w.sin_angle, y));
    g.y = fmaf(w.sin_angle, x, __fmul_rn(w.cos_angle, y));
    g.theta = __fadd_rn(-thw, w.rotation_const);
    return g;
}

// Sphere-trace one ray in grid coordinates.  (x0, dx) run along the FIRST grid index (msg rows),
// (y0, dy) along the second (msg columns) -- the caller has already applied upstream's
// calc_range(y, x, theta) argument swap.  Returns pixels.  `steps` counts distance-field loads.
//
// Bounds: upstream tests the TRUNCATED integers (px < 0 || px >= width ...), so (-1, 0) is cell 0
// and in bounds; cvt.rzi saturates, so +-inf and huge values leave the map like x86's cvttss2si
// INT_MIN does.  Only NaN converts differently (0 here, INT_MIN on x86), hence the one check
// before the loop: a NaN pose or heading leaves the map at once.
//
// Tail mode: 0.4 % of rays need more than TAIL_AFTER steps (they creep along walls at the 1 px minimum step) and,
// being one dependent load per step, they decide when the kernel ends (profiles/r01_timeline.md).  After TAIL_AFTER
// plain steps a ray therefore also touches the cell TAIL_AHEAD px further along itself at every step, so that the
// real sample finds its sector in L1 a dozen steps later.  The touch never influences a result.
#ifndef RL_TAIL_AFTER
#define RL_TAIL_AFTER 32
#endif
#ifndef RL_TAIL_AHEAD
#define RL_TAIL_AHEAD 12
#endif
constexpr int TAIL_AFTER = RL_TAIL_AFTER;
constexpr int TAIL_AHEAD = RL_TAIL_AHEAD;

// The first sample (t = 0) is the pose's own cell whatever the heading, so its load is issued
// before the heading's sin/cos are evaluated and its latency hides behind that arithmetic.
struct FirstSample {
    int px, py;
    float s;     // march field of the cell (valid when inside): +inf = occupied, else max(0.999 d, 1)
    bool inside;
};

__device__ __forceinline__ FirstSample first_sample(const MarchParams &P, float x0, float y0)
{
    FirstSample f;
    f.px = __float2int_rz(x0);   // fmaf(dx, 0, x0) == x0 for every finite dx
    f.py = __float2int_rz(y0);
    f.inside = (x0 == x0) && (y0 == y0) && (unsigned)f.px < (unsigned)P.rows && (unsigned)f.py < (unsigned)P.cols;
    f.s = f.inside ? __ldg(P.dist + (f.px * P.stride + f.py)) : 0.0f;
    return f;
}

// The look-ahead touch of the tail mode is a 4-byte cp.async (LDGSTS through L1) into a shared-memory word nobody
// reads: it pulls the sector into L1 like a load does, but has no destination register, so nothing ever waits for
// it.  (The first form touched with ordinary loads into a ring of four registers, read four steps later.  ptxas
// counts all those loads on ONE scoreboard, and waiting on a scoreboard waits for every load counted on it -- so
// each step really waited for the touch issued one step earlier, a full L2 or DRAM latency: tools/probe_chain.cu,
// 299 cycles per step across rows against 112 when every sample hits L1; tools/sass_ctrl.py shows the barriers.)
// compute-sanitizer's racecheck reports write-after-write hazards between a thread's successive touches into its
// own sink word (profiles/r02_sanitizer_racecheck_march.log): the word is write-only, whichever copy lands last.
__device__ __forceinline__ uint32_t touch_sink_address()
{
    __shared__ float touch_sink[128];   // a word per thread of the 128-thread march CTAs
    return (uint32_t)__cvta_generic_to_shared(&touch_sink[threadIdx.x & 127]);
}

__device__ __forceinline__ void touch(uint32_t sink, const float *cell)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sink), "l"(cell));
}

__device__ __forceinline__ void touches_done()   // nothing of this thread is in flight when its CTA retires
{
    asm volatile("cp.async.wait_all;" ::: "memory");
}

// float -> int, toward zero (cvt.rzi saturates), as a volatile asm: the compiler may not sink it below the exit
// branch of the step it is written in, which is the point of march_padded's ordering.
__device__ __forceinline__ int trunc_here(float v)
{
    int r;
    asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}

// The loop over the NaN-padded field (see below), entered with the parameter t of the second sample.  No sample can
// fall outside the padded field, so the address of the NEXT sample is formed before the exit test of this one is
// resolved: the 18 cycles from FSETP to the branch run in the shadow of the two float->int conversions instead of
// ahead of them (tools/probe_chain.cu: cycles per step of a warp that is alone; tools/sass_ctrl.py: the schedule).
// When the loop ends that address is garbage (t is +inf, NaN or beyond max_range) and is not used.  The cell of the
// LAST sample, needed for the hit distance, is read off the pointer that sample was loaded through -- keeping the
// previous t or cell in a register of its own costs one or two MOVs in every step of the unrolled loop, a division
// of the offset by the row stride (multiply-high) costs half a dozen instructions once per ray.
template <bool COUNT>
__device__ __forceinline__ float march_padded(const MarchParams &P, float x0, float y0, float dx, float dy,
                                              uint32_t &steps, float t)
{
    const float HIT = __int_as_float(0x7f800000);
    float fx = fmaf(dx, t, x0), fy = fmaf(dy, t, y0), s;
    int idx = __float2int_rz(fx) * P.stride + __float2int_rz(fy);
    const float *cell;
    bool tail = true;
#pragma unroll
    for (int it = 1; it < TAIL_AFTER; ++it) {
        cell = P.dist + idx;
        s = __ldg(cell);
        if (COUNT) ++steps;
        t = __fadd_rn(t, s);
        fx = fmaf(dx, t, x0);
        fy = fmaf(dy, t, y0);
        idx = trunc_here(fx) * P.stride + trunc_here(fy);
        if (!(t < P.max_range)) { tail = false; break; }
    }
    if (tail) {
        const uint32_t sink = touch_sink_address();
        const float adx = __fmul_rn(dx, (float)TAIL_AHEAD), ady = __fmul_rn(dy, (float)TAIL_AHEAD);
        for (;;) {
            cell = P.dist + idx;
            s = __ldg(cell);
            if (COUNT) ++steps;
            touch(sink, P.dist + (__float2int_rz(__fadd_rn(fx, adx)) * P.stride + __float2int_rz(__fadd_rn(fy, ady))));
            t = __fadd_rn(t, s);
            fx = fmaf(dx, t, x0);
            fy = fmaf(dy, t, y0);
            idx = trunc_here(fx) * P.stride + trunc_here(fy);
            if (!(t < P.max_range)) break;
        }
        touches_done();
    }
    if (COUNT && s != s) --steps;   // the NaN that ended the ray was read from the padding: the reference reads nothing there
    if (s == HIT) {   // an occupied cell: inside the map, so its offset from cell (0, 0) is row * stride + column
        const uint32_t off = ((uint32_t)(uintptr_t)cell - (uint32_t)(uintptr_t)P.dist) >> 2;   // the padded field is below 4 GiB (rl_marcher_create)
        const uint32_t px = __umulhi(off, P.stride_magic) >> P.stride_shift, py = off - px * (uint32_t)P.stride;
        const float xd = __fsub_rn((float)(int)px, x0);
        const float yd = __fsub_rn((float)(int)py, y0);
        return sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
    }
    return P.max_range;
}

// P.dist is the MARCH FIELD (common.h: march_step_of): s = +inf on an occupied cell, max(0.999 d, 1)
// elsewhere -- exactly the value the reference's loop adds to t after sampling the cell, computed once
// per cell by the ingest with the same two fp32 operations.  A hit therefore shows up as t + inf failing
// `t < max_range`: one exit test per step, and which exit it was is decided once, after the loop.
//
// PADDED: the marcher's copy of the field is surrounded by P.pad cells of NaN.  Every sample of the loop
// is taken at a parameter t < max_range from a pose inside the map, i.e. at most max_range (+1 for the
// truncation, + TAIL_AHEAD for the look-ahead touch) cells outside it -- inside the padding.  Leaving the
// map then needs no test of its own: t + NaN = NaN fails `t < max_range` like a hit does, and NaN != +inf
// sorts it with the misses afterwards.  Three more instructions gone from every step (two ISETP, one BRA).
// EARLY: the padded loop in its early-address form (march_padded).  It shortens the dependent chain of a step by
// 17 % and costs about ten instructions more per ray at the exit: the latency-bound plain kernels take it (config 2:
// single launch 0.0866 -> 0.0824 ms), the issue-bound territory and crash kernels keep the form below (config 5 lost
// 2-3 % with it).  Same samples, same bits either way.
template <bool COUNT, bool PADDED, bool EARLY = false>
__device__ __forceinline__ float march_ray(const MarchParams &P, float x0, float y0, float dx,
                                           float dy, uint32_t &steps, const FirstSample &f0)
{
    const float HIT = __int_as_float(0x7f800000);
    if (!f0.inside || !(dx == dx) || !(dy == dy)) return P.max_range;   // NaN pose/heading or pose outside the map
    if (COUNT) ++steps;
    if (f0.s == HIT) {   // pose inside an occupied cell: distance to that cell's corner
        const float xd = __fsub_rn((float)f0.px, x0);
        const float yd = __fsub_rn((float)f0.py, y0);
        return sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
    }
    float t = f0.s;   // 0 + step
    if (!(t < P.max_range)) return P.max_range;
    if (PADDED && EARLY) return march_padded<COUNT>(P, x0, y0, dx, dy, steps, t);
    int px, py, it = 1;
    float s;
    bool tail = false;
    for (;;) {
        px = __float2int_rz(fmaf(dx, t, x0));
        py = __float2int_rz(fmaf(dy, t, y0));
        if (!PADDED && ((unsigned)px >= (unsigned)P.rows || (unsigned)py >= (unsigned)P.cols)) { s = 0.0f; break; }   // left the map: a miss
        s = __ldg(P.dist + (px * P.stride + py));
        if (COUNT) ++steps;
        t = __fadd_rn(t, s);
        if (!(t < P.max_range)) break;
        if (++it == TAIL_AFTER) { tail = true; break; }
    }
    if (tail) {
        const uint32_t sink = touch_sink_address();
        const float adx = __fmul_rn(dx, (float)TAIL_AHEAD), ady = __fmul_rn(dy, (float)TAIL_AHEAD);
        bool inside = true;
        for (;;) {
            const float fx = fmaf(dx, t, x0), fy = fmaf(dy, t, y0);
            px = __float2int_rz(fx);
            py = __float2int_rz(fy);
            if (!PADDED && ((unsigned)px >= (unsigned)P.rows || (unsigned)py >= (unsigned)P.cols)) { inside = false; break; }
            s = __ldg(P.dist + (px * P.stride + py));
            if (COUNT) ++steps;
            const int ax = __float2int_rz(__fadd_rn(fx, adx)), ay = __float2int_rz(__fadd_rn(fy, ady));
            if (PADDED || ((unsigned)ax < (unsigned)P.rows && (unsigned)ay < (unsigned)P.cols)) touch(sink, P.dist + (ax * P.stride + ay));
            t = __fadd_rn(t, s);
            if (!(t < P.max_range)) break;
        }
        touches_done();
        if (!inside) return P.max_range;
    }
    if (COUNT && PADDED && s != s) --steps;   // the NaN that ended the ray was read from the padding: the reference reads nothing there
    if (s == HIT) {
        const float xd = __fsub_rn((float)px, x0);
        const float yd = __fsub_rn((float)py, y0);
        return sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
    }
    return P.max_range;
}

// Unsigned division by a launch-constant divisor d >= 2, exact for numerators below 2^31:
// n / d == umulhi(n, magic) >> shift with s = ceil(log2 d), magic = ceil(2^(31+s) / d), shift = s-1.
struct FastDiv {
    uint32_t magic, shift, d;
};

__device__ __forceinline__ uint32_t fast_div(uint32_t n, const FastDiv &f) { return __umulhi(n, f.magic) >> f.shift; }

// n / d for ray indices beyond 2^31 (one call with more than 2 G rays): the quotient of the double-precision
// product is off by at most one for n < 2^52, one comparison each way puts it right -- a dozen instructions
// instead of the ~80 of a 64-bit integer division (config 5 in one 4.32 G-ray call: 44.3 -> 43.6 ms).
__device__ __forceinline__ int64_t wide_div(int64_t n, int d, int &rem)
{
    int64_t q = (int64_t)__double2ll_rz(__dmul_rn((double)n, __drcp_rn((double)d)));
    int64_t r = n - q * d;
    if (r < 0) { --q; r += d; }
    if (r >= d) { ++q; r -= d; }
    rem = (int)r;
    return q;
}

}  // namespace rl
