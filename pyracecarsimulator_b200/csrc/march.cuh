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
template <bool COUNT>
__device__ __forceinline__ float march_ray(const MarchParams &P, float x0, float y0, float dx,
                                           float dy, uint32_t &steps)
{
    float t = 0.0f;
    while (t < P.max_range) {
        const float fx = fmaf(dx, t, x0);
        const float fy = fmaf(dy, t, y0);
        // (int) truncates toward zero, so (-1, 0) is cell 0 and in bounds; NaN fails every
        // comparison and leaves the map like x86's cvttss2si INT_MIN does upstream.
        if (!(fx > -1.0f && fx < P.frows && fy > -1.0f && fy < P.fcols)) return P.max_range;
        const int px = __float2int_rz(fx);
        const int py = __float2int_rz(fy);
        const float d = __ldg(P.dist + (px * P.cols + py));
        if (COUNT) ++steps;
        if (d <= 0.0f) {
            const float xd = __fsub_rn((float)px, x0);
            const float yd = __fsub_rn((float)py, y0);
            return sqrtf(fmaf(xd, xd, __fmul_rn(yd, yd)));
        }
        t = __fadd_rn(t, fmaxf(__fmul_rn(d, 0.999f), 1.0f));
    }
    return P.max_range;
}

}  // namespace rl
