// followgap.cu -- batched follow-the-gap action generator: the consumer of a
// scan on every MCTS expansion (followgap/followgap.hpp:104-129 via scripts/mcts.py:262-267).
// One warp per scan: lanes stride over the beams for the clip / closest-return search, the
// "farther than 1.75 m" predicate is packed into ballot words in shared memory, and lane 0 walks
// the words for the first longest run with bit scans.  Same float / double mix as the reference.
#include "common.h"

namespace {

constexpr int FG_WARPS = 4;

__device__ __forceinline__ int next_bit(const uint32_t *w, int pos, int size, bool want_one)
{
    // first index >= pos whose bit equals want_one, or size
    while (pos < size) {
        uint32_t word = w[pos >> 5];
        if (!want_one) word = ~word;
        word &= 0xffffffffu << (pos & 31);
        if (word) {
            const int i = (pos & ~31) + (__ffs(word) - 1);
            return i < size ? i : size;
        }
        pos = (pos & ~31) + 32;
    }
    return size;
}

__global__ void __launch_bounds__(FG_WARPS * 32)
follow_gap_kernel(const float *__restrict__ scans, int64_t n_scans, int size, float max_distance,
                  float max_angle, float angle_inc, float *__restrict__ out)
{
    extern __shared__ uint32_t words_all[];
    const int lane = threadIdx.x & 31;
    const int nwords = (size + 31) >> 5;
    const int64_t scan = (int64_t)blockIdx.x * FG_WARPS + (threadIdx.x >> 5);
    if (scan >= n_scans) return;   // warp-uniform
    uint32_t *w = words_all + (threadIdx.x >> 5) * nwords;
    const float *l = scans + scan * size;
    const int clip_end = size - 10;   // preprocessLidar leaves the last 10 beams alone

    // closest non-zero return, first index on ties (followgap.hpp:111-118)
    float bv = __int_as_float(0x7f800000);
    int bi = 0x7fffffff;
    for (int i = lane; i < size; i += 32) {
        if (i == 0) continue;
        float v = l[i];
        if (i < clip_end && v > max_distance) v = max_distance;
        if (v != 0.0f && v < bv) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    float v0 = l[0];
    if (0 < clip_end && v0 > max_distance) v0 = max_distance;
    const int m = (bi != 0x7fffffff && bv < v0) ? bi : 0;   // NaN v0 compares false: stays 0

    // safety bubble + "gap" predicate packed 32 beams per word (followgap.hpp:66-79, :44)
    for (int base = 0; base < size; base += 32) {
        const int i = base + lane;
        bool far = false;
        if (i < size) {
            float v = l[i];
            if (i < clip_end && v > max_distance) v = max_distance;
            const bool bubble = (i == m) || (i >= m - 5 && i < m + 5 && i > 0 && i < size - 1);
            far = !bubble && v > 1.75f;
        }
        const uint32_t word = __ballot_sync(0xffffffffu, far);
        if (lane == 0) w[base >> 5] = word;
    }
    __syncwarp();
    if (lane != 0) return;

    // first longest run (followgap.hpp:30-64)
    int max_start = 0, max_size = 0, pos = 0;
    while (pos < size) {
        const int start = next_bit(w, pos, size, true);
        if (start >= size) break;
        const int end = next_bit(w, start, size, false);
        if (end - start > max_size) { max_size = end - start; max_start = start; }
        pos = end + 1;
    }
    int best = (max_start + (max_start + max_size + 1)) / 2;
    if (best > size - 1) best = size - 1;   // the reference reads one past the end here
    float angle;
    if (best > size / 2) angle = (float)(-(double)angle_inc * ((size / 2.0) - (double)best));
    else angle = (float)((double)angle_inc * ((double)best - (size / 2.0)));
    angle = 2.0f * __fdiv_rn(angle, l[best]);
    const float lo = (angle < -max_angle) ? -max_angle : angle;   // std::max(angle, -max_angle)
    out[scan] = (max_angle < lo) ? max_angle : lo;                // std::min(lo, max_angle)
}

}  // namespace

extern "C" RL_API int32_t rl_follow_gap(const float *d_scans, int64_t num_scans, int32_t num_rays,
                                        float max_distance, float max_angle, float angle_inc,
                                        float *d_out, void *stream)
{
    if (num_scans < 0 || num_rays < 10 || num_rays > (1 << 20) || (num_scans > 0 && (!d_scans || !d_out)))
        return rl::fail(RL_ERR_BAD_ARG, "rl_follow_gap: bad argument (num_rays must be >= 10)");
    if (num_scans == 0) return RL_OK;
    const int nwords = (num_rays + 31) / 32;
    const size_t smem = (size_t)FG_WARPS * nwords * sizeof(uint32_t);
    const int64_t blocks = (num_scans + FG_WARPS - 1) / FG_WARPS;
    if (blocks > 0x7fffffffLL) return rl::fail(RL_ERR_BAD_ARG, "rl_follow_gap: too many scans for one call");
    if (smem > 48 * 1024)
        RL_CUDA(cudaFuncSetAttribute(follow_gap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    follow_gap_kernel<<<(unsigned)blocks, FG_WARPS * 32, smem, (cudaStream_t)stream>>>(
        d_scans, num_scans, num_rays, max_distance, max_angle, angle_inc, d_out);
    RL_CUDA(cudaGetLastError());
    return RL_OK;
}
