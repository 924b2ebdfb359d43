// common.h -- internal types shared by the translation units behind include/rangelib_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <string>

#include "../../include/rangelib_b200.h"

namespace rl {

void set_error(const std::string &msg);
int32_t fail(int32_t code, const std::string &msg);
// [p, p+bytes) against the ranges page-locked by rl_host_register: 1 = inside one of them, 0 = touches none,
// -1 = starts inside one but runs past its end (only partly page-locked)
int host_registered_range(const void *p, size_t bytes);

#define RL_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return rl::fail(_e == cudaErrorMemoryAllocation ? RL_ERR_OOM : RL_ERR_CUDA,        \
                            std::string(#expr) + ": " + cudaGetErrorString(_e));               \
    } while (0)

// Switches to `device` for the lifetime of the object and restores the caller's device.
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != device && cudaSetDevice(device) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

// World <-> grid constants of an OMap, all narrowed to fp32 exactly as
// PyOMap / RangeMethod::numpy_calc_range do on the host.
struct WorldFrame {
    float scale, angle, origin_x, origin_y, sin_angle, cos_angle, inv_scale, rotation_const;
};

// The march field: what a ray adds to its parameter t after sampling a cell.  The reference's loop is
//     d = dist[cell]; if (d <= 0) hit; t += max(0.999 * d, 1); if (!(t < max_range)) miss;
// max(0.999 * d, 1) is a pure function of the cell, so the ingest evaluates it once per cell (the same two
// fp32 operations, same bits) and stores +inf for occupied cells (d == 0): then t + step = inf fails the one
// remaining test `t < max_range` and which exit it was is decided once, after the loop.  Three instructions
// fewer per march step (FMUL, FMNMX, FSETP) in an issue-bound loop; results are bit-identical.
#ifdef __CUDACC__
__device__ __forceinline__ float march_step_of(float d)
{
    return d <= 0.0f ? __int_as_float(0x7f800000) : fmaxf(__fmul_rn(d, 0.999f), 1.0f);
}
#endif

// Passed by value to every march kernel.
struct MarchParams {
    const float *dist;  // the march field (see march_step_of): cell (row, col) is dist[row * stride + col]
    int rows, cols;     // rows = OMap.width (msg.info.height), cols = OMap.height (msg.info.width)
    int stride;         // floats between rows: cols, or cols + 2*pad for the marcher's padded copy
    uint32_t stride_magic, stride_shift;   // offset / stride == umulhi(offset, magic) >> shift (march.cuh: FastDiv)
    int pad;            // > 0: `pad` cells of NaN surround the map on every side (rows -pad .. rows+pad-1 and
                        // columns -pad .. cols+pad-1 are addressable), and pad exceeds max_range + the tail
                        // look-ahead, so no sample of a ray that starts inside the map needs a bounds test
    float frows, fcols;
    float max_range;    // pixels
    WorldFrame w;
};


}  // namespace rl

struct rl_map {
    int device = 0;
    int rows = 0, cols = 0;
    rl::WorldFrame world{};
    uint8_t *d_occ = nullptr;   // rows*cols, 0/1
    int32_t *d_dist2 = nullptr; // rows*cols exact squared distance
    float *d_dist = nullptr;    // rows*cols sqrt
    float *d_step = nullptr;    // rows*cols march field derived from d_dist (rl::march_step_of): what the kernels read
    float ingest_ms = 0.f;
    std::atomic<int> refs{1};
};

void rl_map_retain(const rl_map *m);
void rl_map_release(const rl_map *m);
