// march.cu -- the ray-marching hot path and its C-ABI entry points (north_star (b)).
// Replaces range_libc's RayMarching / RayMarchingGPU calc_range_many (2-arg and the fork's
// 4-arg fan) and calc_range_repeat_angles; reference call sites scripts/scan_simulator.py:103-106,
// :130-133 and scripts/two_player/scan.py:69-70.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <utility>

#include <cooperative_groups.h>

#include "glibc_trig.cuh"
#include "march.cuh"
#include "marcher.h"

namespace {

using rl::GridPose;
using rl::MarchParams;
using rl::launch_windowed;
using rl::launch_windowed_ex;

constexpr int CTA_THREADS = rl::MARCH_CTA_THREADS;

// Programmatic dependent launch: let the next kernel in the stream start being scheduled as soon as
// every CTA of this one has started (a no-op unless that kernel was launched with the PDL attribute).
__device__ __forceinline__ void release_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }

template <bool COUNT>
__device__ __forceinline__ void flush_steps(uint32_t steps, unsigned long long *counter)
{
    if (COUNT) {
        steps = __reduce_add_sync(0xffffffffu, steps);
        if ((threadIdx.x & 31) == 0 && steps) atomicAdd(counter, (unsigned long long)steps);
    }
}

// ---- one (x, y, theta) row per ray (upstream 2-arg calc_range_many) ----
template <bool COUNT, bool PADDED>
__global__ void __launch_bounds__(CTA_THREADS)
march_many_kernel(MarchParams P, const float *__restrict__ ins, float *__restrict__ outs,
                  int64_t n, unsigned long long *counter)
{
    release_dependents();
    const int64_t i = (int64_t)blockIdx.x * CTA_THREADS + threadIdx.x;
    uint32_t steps = 0;
    if (i < n) {
        const GridPose g = rl::world_to_grid(P.w, ins[3 * i], ins[3 * i + 1], ins[3 * i + 2]);
        const rl::FirstSample f0 = rl::first_sample(P, g.y, g.x);
        float s, c;
        rl::glibc_sincosf(g.theta, &s, &c);
        __stcs(outs + i, __fmul_rn(rl::march_ray<COUNT, PADDED, true>(P, g.y, g.x, c, s, steps, f0), P.w.scale));
    }
    flush_steps<COUNT>(steps, counter);
}

// ---- one thread per ray over the flat (pose, beam) index space: lanes run over adjacent beams ----
// Measured on B200 (tools/tune_march): this finest-grained mapping beats warp-per-pose, beam
// segments per warp, persistent work queues and multi-ray-per-lane variants; the march is bound
// by instruction issue and the latency of its dependent loads (profiles/r01_ncu_full_summary.md),
// and the hardware CTA scheduler is the cheapest load balancer.
// FAN:   beam j heads theta + fmaf(j, fov/num_beams, -fov/2)   (fork's 4-arg calc_range_many)
// !FAN:  beam a heads theta + angles[a]                        (calc_range_repeat_angles)
// Peer output: the fused march + all-gather writes every range straight into the gathered buffer
// of every GPU of the box (its own included) over NVLink, so the transfer overlaps the march ray by
// ray instead of following it as a separate collective.
//   OUT_LOCAL  outs[i] = r
//   OUT_PEERS  one 4-byte store per range and peer (or one multimem.st, replicated by the NVSwitch)
//   OUT_PEERS4 the CTA's 128 ranges are staged in shared memory and leave as 32 16-byte stores per peer
//              (multimem.st.v4.f32 with multicast): a quarter of the NVLink packets for the same bytes
constexpr int MAX_PEERS = 16;
constexpr int OUT_LOCAL = 0, OUT_PEERS = 1, OUT_PEERS4 = 2;
struct PeerOut {
    float *buf[MAX_PEERS];   // gathered buffer of each rank (peer-mapped device pointers)
    int world;
    int multicast;           // buf[0] is an NVLS multicast address: one multimem.st reaches every GPU
    int64_t offset;          // this rank's slot: rank * slot_rays
};

__device__ __forceinline__ void peer_store(const PeerOut &peers, int64_t i, float r)
{
    if (peers.multicast == 2) {   // RL_GATHER_WEAK: no ordering asked of the store; the kernel boundary + barrier publish it
        asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(peers.buf[0] + peers.offset + i), "f"(r) : "memory");
    } else if (peers.multicast) {   // the NVSwitch replicates the store to all GPUs of the multicast group
        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(peers.buf[0] + peers.offset + i), "f"(r)
                     : "memory");
    } else {
#pragma unroll 1
        for (int q = 0; q < peers.world; ++q) peers.buf[q][peers.offset + i] = r;
    }
}

__device__ __forceinline__ void peer_store4(const PeerOut &peers, int64_t i, float4 v)
{
    if (peers.multicast == 2) {
        asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(peers.buf[0] + peers.offset + i),
                     "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                     : "memory");
    } else if (peers.multicast) {
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(peers.buf[0] + peers.offset + i),
                     "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                     : "memory");
    } else {
#pragma unroll 1
        for (int q = 0; q < peers.world; ++q) *reinterpret_cast<float4 *>(peers.buf[q] + peers.offset + i) = v;
    }
}

// Beam j of the pose at `p` (x, y, theta in the world frame): the range in metres.
template <bool FAN, bool COUNT, bool PADDED, bool EARLY>
__device__ __forceinline__ float pose_ray(const MarchParams &P, const float *__restrict__ p, const float *__restrict__ angles,
                                          int j, float fov, float inc, uint32_t &steps)
{
    const float thw = __ldg(p + 2);
    const GridPose g = rl::world_to_grid(P.w, __ldg(p), __ldg(p + 1), thw);
    float thg;
    if (FAN) thg = __fadd_rn(-__fadd_rn(thw, fmaf((float)j, inc, -0.5f * fov)), P.w.rotation_const);
    else thg = __fsub_rn(g.theta, __ldg(angles + j));
    const rl::FirstSample f0 = rl::first_sample(P, g.y, g.x);
    float s, c;
    rl::glibc_sincosf(thg, &s, &c);
    return __fmul_rn(rl::march_ray<COUNT, PADDED, EARLY>(P, g.y, g.x, c, s, steps, f0), P.w.scale);
}

template <bool FAN, bool COUNT, bool SMALL, int OUT, bool PADDED>
__global__ void __launch_bounds__(CTA_THREADS)
march_pose_kernel(MarchParams P, const float *__restrict__ poses, int64_t pose_stride_floats,
                  const float *__restrict__ angles, float *__restrict__ outs, int64_t num_rays_total,
                  int num_beams, rl::FastDiv div, float fov, float inc, unsigned long long *counter,
                  PeerOut peers)
{
    release_dependents();
    __shared__ __align__(16) float stage[OUT == OUT_PEERS4 ? CTA_THREADS : 1];   // read back as float4
    const int64_t i = (int64_t)blockIdx.x * CTA_THREADS + threadIdx.x;
    uint32_t steps = 0;
    if (i < num_rays_total) {
        int64_t k;
        int j;
        if (SMALL) {   // fewer than 2^31 rays and at least 2 beams: multiply-high instead of a divide
            const uint32_t k32 = rl::fast_div((uint32_t)i, div);
            k = k32;
            j = (int)((uint32_t)i - k32 * (uint32_t)num_beams);
        } else {
            k = rl::wide_div(i, num_beams, j);
        }
        const float r = pose_ray<FAN, COUNT, PADDED, true>(P, poses + k * pose_stride_floats, angles, j, fov, inc, steps);
        if (OUT == OUT_PEERS) peer_store(peers, i, r);
        else if (OUT == OUT_PEERS4) stage[threadIdx.x] = r;
        else __stcs(outs + i, r);   // streaming store: the ranges are not read again here (steady state 0.0489 -> 0.0477 ms)
    }
    if (OUT == OUT_PEERS4) {   // peers.offset is a multiple of 4 floats (checked by the host), so is the CTA's base
        __syncthreads();
        if (threadIdx.x < CTA_THREADS / 4) {
            const int64_t b = (int64_t)blockIdx.x * CTA_THREADS + 4 * threadIdx.x;
            if (b + 3 < num_rays_total) {
                peer_store4(peers, b, *reinterpret_cast<const float4 *>(stage + 4 * threadIdx.x));
            } else {
                for (int e = 0; e < 4; ++e)
                    if (b + e < num_rays_total) peer_store(peers, b + e, stage[4 * threadIdx.x + e]);
            }
        }
    }
    flush_steps<COUNT>(steps, counter);
}

// Surround the map's march field with `pad` cells of NaN (see march_ray<.., PADDED>).
__global__ void __launch_bounds__(256)
pad_field_kernel(const float *__restrict__ src, int rows, int cols, int pad, float *__restrict__ dst)
{
    const int stride = cols + 2 * pad;
    const int64_t total = (int64_t)(rows + 2 * pad) * stride;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int r = (int)(i / stride) - pad, c = (int)(i % stride) - pad;
    dst[i] = ((unsigned)r < (unsigned)rows && (unsigned)c < (unsigned)cols) ? src[(size_t)r * cols + c]
                                                                            : __int_as_float(0x7fc00000);
}

// ---- map order + SM territories for large batches ----
// A batch of many poses (a particle set, a pose grid: BASELINE configs 3 and 5 hold one pose per four map cells)
// samples every neighbourhood of the map hundreds of times, yet in the caller's order those samples are spread
// over all SMs and over the whole launch, so nearly every one of them misses L1 (14 % hits on config 3) and the
// kernel runs at the one-sector-per-clock rate at which an SM can send L1 misses to L2 (ncu:
// l1tex__m_l1tex2xbar_req_cycles_active 77 %); with a field larger than L2 (config 5) they miss L2 as well (3.3 TB/s
// of DRAM sector gathers).  Two launches turn that reuse into L1 hits:
//   1. pose_sort_kernel, a counting sort in ONE cooperative launch, puts the poses' indices in MAP ORDER -- Morton
//      order of their cells (16 px or larger, at most 2^16 of them): histogram, scan and scatter phases separated by
//      grid-wide barriers (cub's radix sort: 72 us of kernels in 6 launches, 150 us with their launch gaps);
//   2. march_territory_kernel cuts that order into one contiguous range per SM.  The warps resident on an SM claim 32-ray tasks from
//      their SM's own range (one atomic per TERR_CLAIM tasks), so at any moment an SM works on a few dozen
//      neighbouring poses and sweeps slowly through one compact territory of the map, whose field cells stay in its
//      L1 (74 % hits on config 3, L2 sector reads down 3.2 x).  A warp whose range is used up helps out in the range
//      with the most work left.
// The kernel still reads pose k and writes its ranges at k * num_beams: only the ORDER of the work changes, not a
// bit of any result (tests/test_gpu_round2.py::test_map_order_marching_is_invisible).
namespace cg = cooperative_groups;

constexpr int TERR_CLAIM = 4;          // 32-ray tasks per atomic
constexpr int TERR_CLAIM_STRIDE = 8;   // uint32 between the ranges' counters: one 32-byte sector each
constexpr int SORT_MAX_SIDE_BITS = 8;  // at most 256 x 256 Morton cells: keys fit 16 bits
constexpr int SORT_TILE = 32;          // bins scanned by one CTA in the scan phase

__device__ __forceinline__ uint32_t spread_bits8(uint32_t v)   // abcdefgh -> 0a0b0c0d0e0f0g0h
{
    v &= 0xffu;
    v = (v | (v << 4)) & 0x0f0fu;
    v = (v | (v << 2)) & 0x3333u;
    v = (v | (v << 1)) & 0x5555u;
    return v;
}

struct Territories {
    // counting sort of the poses by map cell
    uint16_t *keys;            // num_poses
    uint32_t *cursor;          // n_bins: count -> exclusive offset inside the bin's tile -> scatter cursor (zeroed before the launch)
    uint32_t *tile_off;        // n_bins / SORT_TILE: poses in all earlier tiles
    uint32_t *perm;            // num_poses: pose indices in map order
    uint32_t n_bins;           // 4^side_bits
    int shift;                 // cell = (row >> shift, col >> shift)
    int64_t num_poses;
    // work distribution
    uint32_t *claims;          // one counter per range (zeroed before the launch): tasks of the range handed out so far
    uint32_t n_ranges;         // = SM count
    uint32_t tasks_per_range;  // 32-ray tasks per range (the last range may be shorter)
    uint32_t n_tasks;
};

__device__ __forceinline__ uint32_t territory_len(const Territories &T, uint32_t r)
{
    const uint64_t base = (uint64_t)r * T.tasks_per_range;
    return base < T.n_tasks ? (uint32_t)min((uint64_t)T.tasks_per_range, (uint64_t)T.n_tasks - base) : 0u;
}

// The range with the most unclaimed tasks (all lanes get the same answer), or 0xffffffff when none is left.
// (Helping in the NEXT range with work left instead -- so that thieves spread out -- measured 5 % slower on uniform
// poses and 24 % slower on a particle cloud, whose ranges differ a lot in cost: the fullest range is the one that
// would finish last.  One task per claim with the next claim issued a task ahead: 3-8 % slower than four per claim.)
__device__ __forceinline__ uint32_t busiest_territory(const Territories &T, unsigned lane)
{
    uint32_t best_left = 0, best_r = 0xffffffffu;
    for (uint32_t q = lane; q < T.n_ranges; q += 32) {
        const uint32_t c = *reinterpret_cast<const volatile uint32_t *>(T.claims + q * TERR_CLAIM_STRIDE);
        const uint32_t len = territory_len(T, q);
        const uint32_t left = c < len ? len - c : 0u;
        if (left > best_left) { best_left = left; best_r = q; }
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
        const uint32_t ol = __shfl_xor_sync(0xffffffffu, best_left, off), orr = __shfl_xor_sync(0xffffffffu, best_r, off);
        if (ol > best_left || (ol == best_left && orr < best_r)) { best_left = ol; best_r = orr; }
    }
    return best_left ? best_r : 0xffffffffu;
}

// Counting with two levels of aggregation.  A particle cloud puts a million poses into a few dozen cells, and
// atomics on one address are served one after the other (a 1 M-pose cloud spent 100 us in this sort with one global
// atomic per warp and key).  So (1) the lanes of a warp that hold the same key act through one leader
// (__match_any_sync), and (2) every CTA counts its own poses in a small shared-memory hash table first and talks to
// the global counters once per distinct key: one atomic per (CTA, key) to add its count, one to reserve its slots
// in the scatter phase.  Keys that find no room in the table (four probes) go to the global counters directly.
constexpr int SORT_THREADS = 1024;   // few, large CTAs: a grid-wide barrier costs one same-address atomic per CTA
constexpr int SORT_SLOT_BITS = 9;    // 512 slots, 6 KB: what is taken from the SM's L1 carve-out stays small
constexpr int SORT_SLOTS = 1 << SORT_SLOT_BITS;
constexpr int SORT_FILL = SORT_SLOTS / 2;   // no new keys beyond this: spread-out poses gain nothing from the table
constexpr int SORT_PROBES = 3;
constexpr uint32_t SLOT_EMPTY = 0xffffffffu;

struct SortTable {
    uint32_t key[SORT_SLOTS];
    uint32_t cnt[SORT_SLOTS];    // phase 1: poses of this CTA with the key; phase 3: handed out so far
    uint32_t base[SORT_SLOTS];   // phase 3: first slot of the block this CTA reserved for the key ...
    uint32_t room[SORT_SLOTS];   // ... and its length (the table's count of phase 1)
    uint32_t used;
};

// A key sits in the first slot of its probe sequence that was empty when it arrived and slots never empty again, so
// a lookup that meets an empty slot knows the key is not in the table.  No new key is admitted once the table is
// half full -- a racy test: a key may be refused for one warp (counted globally) and admitted for a later one, which
// is why the scatter phase fills a key's reserved block by arrival and sends the overflow to the global counter
// (sort_account): per (CTA, key) the table and the global counter hand out exactly as many slots as each counted.
template <bool INSERT>
__device__ __forceinline__ int table_slot(SortTable &tb, uint32_t key)
{
    uint32_t s = (key * 2654435761u) >> (32 - SORT_SLOT_BITS);
#pragma unroll
    for (int p = 0; p < SORT_PROBES; ++p, s = (s + 1) & (SORT_SLOTS - 1)) {
        uint32_t cur = *reinterpret_cast<volatile uint32_t *>(&tb.key[s]);
        if (cur == key) return (int)s;
        if (cur == SLOT_EMPTY) {
            if (!INSERT || *reinterpret_cast<volatile uint32_t *>(&tb.used) >= (uint32_t)SORT_FILL) return -1;
            cur = atomicCAS(&tb.key[s], SLOT_EMPTY, key);
            if (cur == SLOT_EMPTY) { atomicAdd(&tb.used, 1u); return (int)s; }
            if (cur == key) return (int)s;
        }
    }
    return -1;
}

// Phase 1 (SCATTER = false): count this warp's keys.  Phase 3 (SCATTER = true): returns this lane's position
// inside its key's bin.  Warp-synchronous; lanes with !valid only take part in the votes.
template <bool SCATTER>
__device__ __forceinline__ uint32_t sort_account(SortTable &tb, uint32_t *cursor, uint32_t key, bool valid)
{
    const unsigned active = __ballot_sync(0xffffffffu, valid);
    uint32_t pos = 0;
    if (valid) {
        const unsigned lane = threadIdx.x & 31;
        const unsigned peers = __match_any_sync(active, key);
        const int leader = __ffs(peers) - 1;
        const uint32_t n = (uint32_t)__popc(peers);
        uint32_t block_first = 0, block_n = 0, global_first = 0;   // this group's share of the CTA's block, the rest
        if ((int)lane == leader) {
            const int slot = SCATTER ? table_slot<false>(tb, key) : table_slot<true>(tb, key);
            if (slot >= 0) {
                const uint32_t local = atomicAdd(&tb.cnt[slot], n);
                if (SCATTER) {
                    const uint32_t room = tb.room[slot];
                    block_n = room > local ? min(n, room - local) : 0u;
                    block_first = tb.base[slot] + local;
                }
            }
            if (slot < 0 || (SCATTER && block_n < n)) {
                const uint32_t g = atomicAdd(cursor + key, n - block_n);
                if (SCATTER) global_first = g;
            }
        }
        if (SCATTER) {
            block_first = __shfl_sync(peers, block_first, leader);
            block_n = __shfl_sync(peers, block_n, leader);
            global_first = __shfl_sync(peers, global_first, leader);
            const uint32_t rank = (uint32_t)__popc(peers & ((1u << lane) - 1u));
            pos = rank < block_n ? block_first + rank : global_first + (rank - block_n);
        }
    }
    return pos;
}

__global__ void __launch_bounds__(SORT_THREADS)
pose_sort_kernel(MarchParams P, const float *__restrict__ poses, int64_t pose_stride_floats, Territories T)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ SortTable tb;
    __shared__ uint32_t warp_sum[SORT_THREADS / 32];
    const int64_t tid = (int64_t)blockIdx.x * SORT_THREADS + threadIdx.x, nthreads = (int64_t)gridDim.x * SORT_THREADS;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int s = threadIdx.x; s < SORT_SLOTS; s += SORT_THREADS) { tb.key[s] = SLOT_EMPTY; tb.cnt[s] = 0; }
    if (threadIdx.x == 0) tb.used = 0;
    __syncthreads();
    // 1. keys + histogram (whole warps iterate together: the accounting is warp-synchronous)
    for (int64_t k0 = tid - lane; k0 < T.num_poses; k0 += nthreads) {
        const int64_t k = k0 + lane;
        const bool valid = k < T.num_poses;
        uint32_t key = 0;
        if (valid) {
            const float *p = poses + k * pose_stride_floats;
            const GridPose g = rl::world_to_grid(P.w, __ldg(p), __ldg(p + 1), 0.0f);
            // saturating conversions; NaN -> 0: any key is fine, the order is only a cache hint
            const int row = min(max(__float2int_rz(g.y), 0), P.rows - 1), col = min(max(__float2int_rz(g.x), 0), P.cols - 1);
            key = (spread_bits8((uint32_t)(row >> T.shift)) << 1) | spread_bits8((uint32_t)(col >> T.shift));
            T.keys[k] = (uint16_t)key;
        }
        sort_account<false>(tb, T.cursor, key, valid);
    }
    __syncthreads();
    for (int s = threadIdx.x; s < SORT_SLOTS; s += SORT_THREADS)
        if (tb.key[s] != SLOT_EMPTY) atomicAdd(T.cursor + tb.key[s], tb.cnt[s]);
    grid.sync();
    // 2a. every warp scans tiles of SORT_TILE bins: count -> exclusive offset inside the tile; tile total -> tile_off
    const uint32_t n_tiles = T.n_bins / SORT_TILE;
    for (uint32_t tile = (uint32_t)(tid >> 5); tile < n_tiles; tile += (uint32_t)(nthreads >> 5)) {
        const uint32_t c = T.cursor[tile * SORT_TILE + lane];
        uint32_t inc = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, inc, off);
            if ((int)lane >= off) inc += o;
        }
        T.cursor[tile * SORT_TILE + lane] = inc - c;
        if (lane == 31) T.tile_off[tile] = inc;
    }
    grid.sync();
    // 2b. CTA 0 turns the tile totals into exclusive offsets
    if (blockIdx.x == 0) {
        uint32_t carry = 0;
        for (uint32_t base = 0; base < n_tiles; base += SORT_THREADS) {
            const uint32_t i = base + threadIdx.x;
            const uint32_t c = i < n_tiles ? T.tile_off[i] : 0u;
            uint32_t inc = c;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, inc, off);
                if ((int)lane >= off) inc += o;
            }
            if (lane == 31) warp_sum[warp] = inc;
            __syncthreads();
            uint32_t before = carry;
            for (unsigned w = 0; w < warp; ++w) before += warp_sum[w];
            if (i < n_tiles) T.tile_off[i] = before + inc - c;
            for (unsigned w = 0; w < SORT_THREADS / 32; ++w) carry += warp_sum[w];
            __syncthreads();
        }
    }
    // 3. scatter the pose indices to their cells' slots; the CTA first reserves a block per key of its table
    for (int s = threadIdx.x; s < SORT_SLOTS; s += SORT_THREADS) {
        if (tb.key[s] != SLOT_EMPTY) {
            tb.room[s] = tb.cnt[s];
            tb.base[s] = atomicAdd(T.cursor + tb.key[s], tb.cnt[s]);
            tb.cnt[s] = 0;
        }
    }
    grid.sync();   // (also the CTA barrier between the reservation and its use)
    for (int64_t k0 = tid - lane; k0 < T.num_poses; k0 += nthreads) {
        const int64_t k = k0 + lane;
        const bool valid = k < T.num_poses;
        const uint32_t key = valid ? T.keys[k] : 0u;
        const uint32_t pos = sort_account<true>(tb, T.cursor, key, valid);
        if (valid) T.perm[T.tile_off[key / SORT_TILE] + pos] = (uint32_t)k;
    }
}

// PEERS: the fused march + all-gather (OUT_PEERS of march_pose_kernel): every range goes straight to slot `rank` of
// every GPU's gathered buffer, at the caller's index (4-byte stores: a pose's block of ranges starts on a 4-byte
// boundary only, and at 8 GPUs the step is bound by NVLink ingress whatever the store width).
template <bool FAN, bool COUNT, bool SMALL, bool PADDED, bool PEERS>
__global__ void __launch_bounds__(CTA_THREADS, 2048 / CTA_THREADS)   // 32 registers, like the plain kernel: every warp slot of the SM
march_territory_kernel(MarchParams P, const float *__restrict__ poses, int64_t pose_stride_floats,
                       const float *__restrict__ angles, float *__restrict__ outs, int64_t num_rays_total,
                       int num_beams, rl::FastDiv div, float fov, float inc, unsigned long long *counter, Territories T,
                       PeerOut peers)
{
    release_dependents();
    const unsigned lane = threadIdx.x & 31;
    uint32_t smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    uint32_t r = smid % T.n_ranges;
    uint32_t steps = 0;
    while (r != 0xffffffffu) {
        const uint32_t len = territory_len(T, r);
        uint32_t c = 0;
        if (lane == 0) c = atomicAdd(T.claims + r * TERR_CLAIM_STRIDE, (uint32_t)TERR_CLAIM);
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c >= len) {   // this range is used up: help where the most work is left
            r = busiest_territory(T, lane);
            continue;
        }
        const uint32_t first = r * T.tasks_per_range + c, last = first + min((uint32_t)TERR_CLAIM, len - c);
#pragma unroll 1
        for (uint32_t task = first; task < last; ++task) {
            const int64_t i = (int64_t)task * 32 + lane;
            if (i < num_rays_total) {
                int64_t k;
                int j;
                if (SMALL) {
                    const uint32_t k32 = rl::fast_div((uint32_t)i, div);
                    k = k32;
                    j = (int)((uint32_t)i - k32 * (uint32_t)num_beams);
                } else {
                    k = rl::wide_div(i, num_beams, j);
                }
                if (T.perm) k = __ldg(T.perm + k);   // null: the caller's order by territories (RL_TERRITORY_IDENTITY, measurements)
                const float rng = pose_ray<FAN, COUNT, PADDED, false>(P, poses + k * pose_stride_floats, angles, j, fov, inc, steps);
                if (PEERS) peer_store(peers, k * num_beams + j, rng);
                else __stcs(outs + (k * num_beams + j), rng);   // streaming store: the ranges are not read again here
            }
        }
    }
    flush_steps<COUNT>(steps, counter);
}

// Ranges that already exist on this GPU -> slot `rank` of every GPU's gathered buffer (16-byte stores).  The
// same stores as the fused march without the march: what the NVLink / NVSwitch path alone takes for these
// bytes (bench.py reports the fused step against it), and the gather for ranges produced by other means.
__global__ void __launch_bounds__(256)
gather_copy_kernel(const float *__restrict__ src, int64_t n, PeerOut peers)
{
    const int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 + 3 < n) {
        peer_store4(peers, i4, *reinterpret_cast<const float4 *>(src + i4));
    } else {
        for (int64_t i = i4; i < n; ++i) peer_store(peers, i, src[i]);
    }
}

__global__ void trig_probe_kernel(const float *__restrict__ in, float *__restrict__ s,
                                  float *__restrict__ c, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rl::glibc_sincosf(in[i], s + i, c + i);
}

}  // namespace

namespace rl {
const MarchParams &marcher_params(const rl_marcher *m) { return m->P; }
int marcher_device(const rl_marcher *m) { return m->map->device; }
}  // namespace rl

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

bool is_pinned_byte(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// Page-locked over its WHOLE extent?  A caller may hand a longer view over the start of a buffer that was
// page-locked (by rl_host_register or by itself) for a shorter length: both ends are asked, and a range the
// library registered must lie inside one registration.  Anything else takes the staged path.
bool is_pinned_range(const void *p, size_t bytes)
{
    if (bytes == 0) return false;
    if (!is_pinned_byte(p) || !is_pinned_byte(static_cast<const char *>(p) + bytes - 1)) return false;
    return rl::host_registered_range(p, bytes) >= 0;
}

int32_t launch_many(rl_marcher *m, const float *d_ins, float *d_outs, int64_t n, cudaStream_t s, bool pdl = false)
{
    if (n == 0) return RL_OK;
    const int64_t blocks = (n + CTA_THREADS - 1) / CTA_THREADS;
    if (blocks > 0x7fffffffLL) return rl::fail(RL_ERR_BAD_ARG, "calc_range_many: too many rays for one call");
    unsigned long long *ctr = m->count ? m->d_steps : nullptr;
    const bool padded = m->P.pad > 0;
    if (m->count) {
        if (padded) RL_CUDA(launch_windowed(m, march_many_kernel<true, true>, (unsigned)blocks, s, pdl, m->P, d_ins, d_outs, n, ctr));
        else RL_CUDA(launch_windowed(m, march_many_kernel<true, false>, (unsigned)blocks, s, pdl, m->P, d_ins, d_outs, n, ctr));
    } else {
        if (padded) RL_CUDA(launch_windowed(m, march_many_kernel<false, true>, (unsigned)blocks, s, pdl, m->P, d_ins, d_outs, n, ctr));
        else RL_CUDA(launch_windowed(m, march_many_kernel<false, false>, (unsigned)blocks, s, pdl, m->P, d_ins, d_outs, n, ctr));
    }
    RL_CUDA(cudaGetLastError());
    return RL_OK;
}

rl::FastDiv make_fast_div(int d)
{
    rl::FastDiv f{0, 0, (uint32_t)d};
    if (d >= 2) {
        int s = 1;
        while ((1u << s) < (uint32_t)d) ++s;
        f.magic = (uint32_t)((((uint64_t)1 << (31 + s)) + d - 1) / d);
        f.shift = (uint32_t)(s - 1);
    }
    return f;
}

// One launch of march_pose_kernel.  peers == nullptr: ranges to d_outs; otherwise to the gathered buffers.
template <bool FAN>
int32_t launch_pose(rl_marcher *m, const float *d_poses, int64_t stride_rows, const float *d_angles,
                    float *d_outs, int64_t num_poses, int num_beams, float fov, cudaStream_t s,
                    bool pdl = false, const PeerOut *peers = nullptr)
{
    if (num_poses == 0 || num_beams == 0) return RL_OK;
    const int64_t total = num_poses * num_beams;
    const int64_t blocks = (total + CTA_THREADS - 1) / CTA_THREADS;
    if (blocks > 0x7fffffffLL) return rl::fail(RL_ERR_BAD_ARG, "calc_range: too many rays for one call");
    const rl::FastDiv div = make_fast_div(num_beams);
    const float inc = fov / (float)num_beams;   // IEEE division, same bits as the oracle's
    const bool small = num_beams >= 2 && total < ((int64_t)1 << 31);
    unsigned long long *ctr = m->count ? m->d_steps : nullptr;
    const int64_t stride_floats = stride_rows * 3;
    const PeerOut po = peers ? *peers : PeerOut{};
    // large batches are marched in map order, by SM territories (see march_territory_kernel); scratch is stream-ordered
    // Worth it when the poses are dense enough to share field cells (L2-resident field: at least one pose per 16 map
    // cells and 24 M rays -- 1 M x 60 on a 2049^2 map gains 19 %, 65 536 x 1080 gains 3 %, smaller batches lose to
    // the sort's ~50 us) or when the field is larger than L2 and locality saves DRAM sector gathers (config 5: +80 %).
    bool by_territories = false;
    // With peer output only between two GPUs and plain peer stores: the territory kernel stores 4 bytes at a time, and
    // multimem.st is bound by the number of stores (config 5's gathered pieces at 4 GPUs: 98.8 Grays/s against 157 for
    // the plain kernel's 16-byte multimem stores; between 2 GPUs with peer stores 165 against 82).
    const bool peers_ok = !peers || (m->gather_territories && po.world <= 2 && po.multicast == 0);
    if (m->sort_poses && peers_ok && num_poses < ((int64_t)1 << 32) && blocks * (CTA_THREADS / 32) < ((int64_t)1 << 32)) {
        if (m->sort_forced) by_territories = num_poses >= m->sort_min_poses;
        else if (m->field_beyond_l2) by_territories = num_poses >= 16384 && total >= ((int64_t)16 << 20);
        else by_territories = num_poses * 16 >= (int64_t)m->P.rows * m->P.cols && total >= ((int64_t)24 << 20);
    }
    if (by_territories) {
        Territories terr{};
        int side = m->P.rows > m->P.cols ? m->P.rows : m->P.cols, side_bits = 1;
        terr.shift = m->sort_shift;
        while (((side - 1) >> terr.shift) >= (1 << SORT_MAX_SIDE_BITS)) ++terr.shift;
        while ((1 << side_bits) <= ((side - 1) >> terr.shift)) ++side_bits;
        if (side_bits < 3) side_bits = 3;   // at least SORT_TILE * 2 bins
        terr.n_bins = 1u << (2 * side_bits);
        terr.num_poses = num_poses;
        auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
        const size_t keys_b = up((size_t)num_poses * sizeof(uint16_t)), perm_b = up((size_t)num_poses * sizeof(uint32_t)),
                     tile_b = up((size_t)terr.n_bins / SORT_TILE * sizeof(uint32_t)),
                     zero_b = up((size_t)terr.n_bins * sizeof(uint32_t)) + up((size_t)m->sm_count * TERR_CLAIM_STRIDE * sizeof(uint32_t));
        void *scratch = nullptr;
        // the marcher's own pool keeps the scratch between calls (the device's default pool hands its memory back to
        // the driver at every synchronisation: 0.2 ms per call to get it again, measured)
        if (m->scratch_pool && cudaMallocFromPoolAsync(&scratch, zero_b + keys_b + perm_b + tile_b, m->scratch_pool, s) == cudaSuccess) {
            char *base = static_cast<char *>(scratch);
            terr.cursor = reinterpret_cast<uint32_t *>(base);
            terr.claims = reinterpret_cast<uint32_t *>(base + up((size_t)terr.n_bins * sizeof(uint32_t)));
            terr.keys = reinterpret_cast<uint16_t *>(base + zero_b);
            terr.perm = reinterpret_cast<uint32_t *>(base + zero_b + keys_b);
            terr.tile_off = reinterpret_cast<uint32_t *>(base + zero_b + keys_b + perm_b);
            terr.n_ranges = (uint32_t)m->sm_count;
            terr.n_tasks = (uint32_t)((total + 31) / 32);
            terr.tasks_per_range = (terr.n_tasks + terr.n_ranges - 1) / terr.n_ranges;
            // whole claims per range, so that only a range's last claim can be short
            terr.tasks_per_range = (terr.tasks_per_range + TERR_CLAIM - 1) / TERR_CLAIM * TERR_CLAIM;
            bool sorted = cudaMemsetAsync(scratch, 0, zero_b, s) == cudaSuccess;
            if (m->territory_identity) terr.perm = nullptr;
            else if (sorted) {
                static const int sort_per_sm = [] {
                    int v = 0;
                    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, pose_sort_kernel, SORT_THREADS, 0) != cudaSuccess) cudaGetLastError();
                    return v < 1 ? 1 : v;
                }();
                int64_t grid = (int64_t)sort_per_sm * m->sm_count;   // co-resident: grid-wide barriers
                const int64_t need = (num_poses + SORT_THREADS - 1) / SORT_THREADS;
                if (grid > need) grid = need;
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3((unsigned)grid);
                cfg.blockDim = dim3(SORT_THREADS);
                cfg.stream = s;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeCooperative;
                attr[0].val.cooperative = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                sorted = cudaLaunchKernelEx(&cfg, pose_sort_kernel, m->P, d_poses, stride_floats, terr) == cudaSuccess;
            }
            if (!sorted) {   // e.g. no room for a cooperative launch: the caller's order will do
                cudaGetLastError();
                cudaFreeAsync(scratch, s);
                scratch = nullptr;
            }
#define RL_TERR_FAIL(expr)                                                                                   \
            do {                                                                                             \
                const cudaError_t e_ = (expr);                                                               \
                if (e_ != cudaSuccess) {                                                                     \
                    cudaFreeAsync(scratch, s);                                                               \
                    return rl::fail(RL_ERR_CUDA, std::string("march_territory_kernel: ") + cudaGetErrorString(e_)); \
                }                                                                                            \
            } while (0)
#define RL_TERR2(COUNT, SMALL, PADDED, PEERS)                                                                \
            do {                                                                                             \
                auto kern = march_territory_kernel<FAN, COUNT, SMALL, PADDED, PEERS>;                        \
                static const int per_sm = [&] {                                                              \
                    int v = 0;                                                                               \
                    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kern, CTA_THREADS, 0) != cudaSuccess) cudaGetLastError(); \
                    return v < 1 ? 1 : v;                                                                    \
                }();                                                                                         \
                int64_t grid = (int64_t)per_sm * m->sm_count;                                                \
                if (grid > blocks) grid = blocks;                                                            \
                RL_TERR_FAIL(launch_windowed_ex(m, kern, (unsigned)grid, s, false, false, m->P, d_poses, stride_floats, d_angles, \
                                                d_outs, total, num_beams, div, fov, inc, ctr, terr, po));    \
            } while (0)
#define RL_TERR(COUNT, SMALL, PEERS) do { if (m->P.pad > 0) RL_TERR2(COUNT, SMALL, true, PEERS); else RL_TERR2(COUNT, SMALL, false, PEERS); } while (0)
            if (sorted) {
                if (peers) { if (small) RL_TERR(false, true, true); else RL_TERR(false, false, true); }
                else if (m->count) { if (small) RL_TERR(true, true, false); else RL_TERR(true, false, false); }
                else { if (small) RL_TERR(false, true, false); else RL_TERR(false, false, false); }
                cudaFreeAsync(scratch, s);
                RL_CUDA(cudaGetLastError());
                return RL_OK;
            }
#undef RL_TERR
#undef RL_TERR2
#undef RL_TERR_FAIL
        }
        cudaGetLastError();   // no scratch: march in the caller's order
    }
#define RL_LAUNCH2(COUNT, SMALL, OUT, PADDED)                                                      \
    RL_CUDA(launch_windowed(m, march_pose_kernel<FAN, COUNT, SMALL, OUT, PADDED>, (unsigned)blocks, s, pdl, m->P, d_poses, \
                            stride_floats, d_angles, d_outs, total, num_beams, div, fov, inc, ctr, po))
#define RL_LAUNCH(COUNT, SMALL, OUT)                                                               \
    do { if (m->P.pad > 0) RL_LAUNCH2(COUNT, SMALL, OUT, true); else RL_LAUNCH2(COUNT, SMALL, OUT, false); } while (0)
    if (peers) {
        // 16-byte stores need this rank's slot to start on a 16-byte boundary; RL_GATHER_VEC=0 forces 4-byte stores
        static const bool vec_ok = [] { const char *e = std::getenv("RL_GATHER_VEC"); return !(e && e[0] == '0'); }();
        bool vec = vec_ok && (po.offset % 4) == 0;
        for (int q = 0; q < (po.multicast ? 1 : po.world); ++q) vec = vec && (reinterpret_cast<uintptr_t>(po.buf[q]) & 15) == 0;
        if (vec) { if (small) RL_LAUNCH(false, true, OUT_PEERS4); else RL_LAUNCH(false, false, OUT_PEERS4); }
        else { if (small) RL_LAUNCH(false, true, OUT_PEERS); else RL_LAUNCH(false, false, OUT_PEERS); }
    } else if (m->count) {
        if (small) RL_LAUNCH(true, true, OUT_LOCAL); else RL_LAUNCH(true, false, OUT_LOCAL);
    } else {
        if (small) RL_LAUNCH(false, true, OUT_LOCAL); else RL_LAUNCH(false, false, OUT_LOCAL);
    }
#undef RL_LAUNCH2
#undef RL_LAUNCH
    RL_CUDA(cudaGetLastError());
    return RL_OK;
}

// Host-pointer calls run as a software pipeline over sub-chunks on two streams: while one
// sub-chunk's ranges travel device->host its successor is already marching (and, for the
// row-per-ray form, its inputs are travelling host->device on the other copy engine).
// `units` are poses (fan / repeat_angles) or rays (many); every unit has `in_floats` input and
// `out_floats` output floats.  `launch(first_unit, count, d_in, d_out, stream)` enqueues the kernel.
template <typename Launch>
int32_t host_pipeline_run(rl_marcher *m, const float *ins, int64_t in_stride_floats, int in_floats,
                          float *outs, int64_t units, int64_t out_floats, Launch launch)
{
    if (units == 0 || out_floats == 0) return RL_OK;
    int64_t big = (int64_t)(HOST_CHUNK_RAYS / (size_t)out_floats);   // units per staged chunk
    if (big < 1) big = 1;
    const bool in_pinned = in_stride_floats == in_floats && is_pinned_range(ins, (size_t)units * in_floats * sizeof(float));
    const bool out_pinned = is_pinned_range(outs, (size_t)units * out_floats * sizeof(float));
    cudaStream_t st[2] = {m->stream, m->stream2};
    for (int64_t b = 0; b < units; b += big) {
        const int64_t c = (units - b < big) ? units - b : big;
        int32_t rc = ensure(&m->h_in, &m->d_in, &m->cap_in, (size_t)c * in_floats);
        if (rc != RL_OK) return rc;
        const float *src = ins + b * in_stride_floats;
        if (!in_pinned) {   // gather (strided fork layout) or stage (pageable) into pinned memory
            if (in_stride_floats == in_floats) std::memcpy(m->h_in, src, (size_t)c * in_floats * sizeof(float));
            else for (int64_t k = 0; k < c; ++k) std::memcpy(m->h_in + k * in_floats, src + k * in_stride_floats, in_floats * sizeof(float));
            src = m->h_in;
        }
        // Zero-copy output: when the caller's buffer is page-locked (mapped under unified addressing) the
        // march kernel stores its ranges straight into it over PCIe -- no device->host copy to wait for,
        // the transfer overlaps the march warp by warp.  RL_HOST_ZEROCOPY=0 falls back to the staged path.
        // Buffers page-locked by rl_host_register (pageable memory pinned after the fact) take the copy-engine
        // pipeline below instead: kernel stores into them measured 604 us against 411 us for the DMA.
        static const bool zero_copy_ok = [] { const char *e = std::getenv("RL_HOST_ZEROCOPY"); return !(e && e[0] == '0'); }();
        if (out_pinned && zero_copy_ok && rl::host_registered_range(outs, (size_t)units * out_floats * sizeof(float)) == 0) {
            float *d_alias = nullptr;
            if (cudaHostGetDevicePointer((void **)&d_alias, outs + b * out_floats, 0) == cudaSuccess && d_alias) {
                // a handful of poses (the single scan of the 50 ms tick): the kernel reads them from the page-locked
                // staging buffer itself -- one API call and one DMA round trip less than copying 12 bytes first
                const float *d_src = m->d_in;
                float *mapped = nullptr;
                if ((size_t)c * in_floats <= 256 && cudaHostGetDevicePointer((void **)&mapped, const_cast<float *>(src), 0) == cudaSuccess && mapped)
                    d_src = mapped;
                else {
                    cudaGetLastError();
                    RL_CUDA(cudaMemcpyAsync(m->d_in, src, (size_t)c * in_floats * sizeof(float), cudaMemcpyHostToDevice, st[0]));
                }
                rc = launch(0, c, d_src, d_alias, st[0]);
                if (rc != RL_OK) return rc;
                RL_CUDA(cudaStreamSynchronize(st[0]));
                continue;
            }
            cudaGetLastError();
        }
        // staged / copy-engine path: device ranges, plus a pinned bounce buffer only for pageable outputs
        rc = ensure(nullptr, &m->d_out, &m->cap_out, (size_t)c * out_floats);
        if (rc != RL_OK) return rc;
        if (!out_pinned && m->cap_hout < (size_t)c * out_floats) {
            cudaFreeHost(m->h_out);
            m->h_out = nullptr;
            m->cap_hout = 0;
            RL_CUDA(cudaMallocHost(&m->h_out, (size_t)c * out_floats * sizeof(float)));
            m->cap_hout = (size_t)c * out_floats;
        }
        float *dst = out_pinned ? outs + b * out_floats : m->h_out;
        // Sub-chunks of >= 512K ranges (at most 4: every sub-chunk costs ~12 us of host API time and a few us of
        // cross-stream latency, measured in profiles/r02_e2e_probe.txt; RL_HOST_SUBCHUNKS overrides the count).
        // Kernels go back to back on the compute stream, the copies back to back on the copy stream, each
        // waiting only for its own kernel's event: once the first sub-chunk is marched the PCIe link never
        // idles.  (Kernel and copy of a sub-chunk on ONE stream, alternating between two streams -- the round-1
        // scheme -- queues kernel i+2 behind copy i and leaves the link idle while it runs.)  For buffers from
        // cudaHostAlloc the zero-copy stores above remain faster (368 vs 389 us for 17.7 MB); this path serves
        // pageable buffers and buffers page-locked after the fact (1783 -> 1487 us and 415 -> 404 us per call).
        int64_t bounds[33];
        int64_t want = (c * out_floats) >> 19;
        want = want < 1 ? 1 : (want > 4 ? 4 : want);
        if (const char *e = std::getenv("RL_HOST_SUBCHUNKS")) { const long v = std::atol(e); if (v >= 1 && v <= 32) want = v; }
        if (want > c) want = c;
        const int nsub = (int)want;
        for (int i = 0; i <= nsub; ++i) bounds[i] = c * i / nsub;
        for (int i = 0; i < nsub; ++i) {
            if (!m->sub_ev[i]) RL_CUDA(cudaEventCreateWithFlags(&m->sub_ev[i], cudaEventDisableTiming));
            if (!out_pinned && !m->sub_done[i]) RL_CUDA(cudaEventCreateWithFlags(&m->sub_done[i], cudaEventDisableTiming));
        }
        const bool one_h2d = (size_t)c * in_floats * sizeof(float) <= ((size_t)1 << 20);
        cudaStream_t compute = st[0], copy = st[1];
        if (one_h2d)
            RL_CUDA(cudaMemcpyAsync(m->d_in, src, (size_t)c * in_floats * sizeof(float), cudaMemcpyHostToDevice, compute));
        for (int i = 0; i < nsub; ++i) {
            const int64_t u = bounds[i], n = bounds[i + 1] - bounds[i];
            if (n <= 0) continue;
            if (!one_h2d)
                RL_CUDA(cudaMemcpyAsync(m->d_in + u * in_floats, src + u * in_floats, (size_t)n * in_floats * sizeof(float),
                                        cudaMemcpyHostToDevice, compute));
            rc = launch(u, n, m->d_in + u * in_floats, m->d_out + u * out_floats, compute);
            if (rc != RL_OK) return rc;
            RL_CUDA(cudaEventRecord(m->sub_ev[i], compute));
            RL_CUDA(cudaStreamWaitEvent(copy, m->sub_ev[i], 0));
            RL_CUDA(cudaMemcpyAsync(dst + u * out_floats, m->d_out + u * out_floats, (size_t)n * out_floats * sizeof(float),
                                    cudaMemcpyDeviceToHost, copy));
            if (!out_pinned) RL_CUDA(cudaEventRecord(m->sub_done[i], copy));
        }
        if (!out_pinned) {   // pageable caller buffer: unstage each sub-chunk as soon as it has landed
            for (int i = 0; i < nsub; ++i) {
                const int64_t u = bounds[i], n = bounds[i + 1] - bounds[i];
                if (n <= 0) continue;
                RL_CUDA(cudaEventSynchronize(m->sub_done[i]));
                std::memcpy(outs + (b + u) * out_floats, m->h_out + u * out_floats, (size_t)n * out_floats * sizeof(float));
            }
        }
        RL_CUDA(cudaStreamSynchronize(copy));
        RL_CUDA(cudaStreamSynchronize(compute));
    }
    return RL_OK;
}

// A call that fails half way must not leave copies or kernels in flight that still write the caller's buffers or
// read staging memory the next call may reallocate: drain both streams before reporting the error.
template <typename Launch>
int32_t host_pipeline(rl_marcher *m, const float *ins, int64_t in_stride_floats, int in_floats,
                      float *outs, int64_t units, int64_t out_floats, Launch launch)
{
    const int32_t rc = host_pipeline_run(m, ins, in_stride_floats, in_floats, outs, units, out_floats, launch);
    if (rc != RL_OK) {
        cudaStreamSynchronize(m->stream);
        cudaStreamSynchronize(m->stream2);
        cudaGetLastError();
    }
    return rc;
}

// The persisting-L2 carve-out is a device-wide limit: remember what it was before the first marcher on a
// device raised it and put it back when the last such marcher goes.
struct L2Carve { size_t before = 0; int users = 0; };
std::mutex g_l2_mu;
L2Carve g_l2[64];

int32_t fill_peers(PeerOut &po, void *const *peer_bufs, int32_t world, int32_t rank, int64_t slot_rays, uint32_t flags,
                   const char *who)
{
    for (int q = 0; q < world; ++q) {
        if (!peer_bufs[q]) return rl::fail(RL_ERR_BAD_ARG, std::string(who) + ": null peer buffer");
        po.buf[q] = static_cast<float *>(peer_bufs[q]);
    }
    po.world = world;
    po.multicast = (flags & RL_GATHER_MULTICAST) ? ((flags & RL_GATHER_WEAK) ? 2 : 1) : 0;
    po.offset = (int64_t)rank * slot_rays;
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
    if (flags & ~(uint32_t)(RL_FLAG_NO_L2_WINDOW | RL_FLAG_NO_PADDED_FIELD | RL_FLAG_NO_POSE_SORT)) return rl::fail(RL_ERR_BAD_ARG, "rl_marcher_create: unknown flag");
    rl::DeviceGuard guard(map->device);
    if (!guard.ok) return rl::fail(RL_ERR_CUDA, "rl_marcher_create: cudaSetDevice failed");
    rl_marcher *m = new (std::nothrow) rl_marcher();
    if (!m) return rl::fail(RL_ERR_OOM, "rl_marcher_create: host allocation failed");
    rl_map_retain(map);
    m->map = map;
    m->flags = flags;
    m->P.dist = map->d_step;
    m->P.rows = map->rows;
    m->P.cols = map->cols;
    m->P.stride = map->cols;
    m->P.pad = 0;
    m->P.frows = (float)map->rows;
    m->P.fcols = (float)map->cols;
    m->P.max_range = max_range_px;
    m->P.w = map->world;
    // The marcher's own copy of the march field, surrounded by NaN far enough that no sample of a ray can fall
    // outside it (march_ray<.., PADDED>): every sample lies within max_range (+1 for the truncation) of a pose
    // inside the map, the tail look-ahead TAIL_AHEAD further.  Very long ranges keep the bounds-tested kernels.
    if (!(flags & RL_FLAG_NO_PADDED_FIELD) && max_range_px <= 2048.0f) {
        const int pad = (int)std::ceil(max_range_px) + rl::TAIL_AHEAD + 4;
        const int64_t stride = (int64_t)map->cols + 2 * pad, prow = (int64_t)map->rows + 2 * pad;
        if (prow * stride < ((int64_t)1 << 30) && cudaMalloc(&m->d_field, (size_t)(prow * stride) * sizeof(float)) == cudaSuccess) {
            const int64_t total = prow * stride;
            pad_field_kernel<<<(unsigned)((total + 255) / 256), 256>>>(map->d_step, map->rows, map->cols, pad, m->d_field);
            if (cudaDeviceSynchronize() == cudaSuccess) {
                m->P.dist = m->d_field + (int64_t)pad * stride + pad;   // cell (0, 0)
                m->P.stride = (int)stride;
                const rl::FastDiv sd = make_fast_div((int)stride);
                m->P.stride_magic = sd.magic;
                m->P.stride_shift = sd.shift;
                m->P.pad = pad;
                m->field_bytes = (size_t)total * sizeof(float);
            } else {
                cudaFree(m->d_field);
                m->d_field = nullptr;
            }
        }
        cudaGetLastError();
    }
    cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, map->device);
    if (!(flags & RL_FLAG_NO_L2_WINDOW) && map->device < 64) {
        // persisting-L2 carve-out large enough for the distance field: a device-wide limit, raised here and put
        // back by rl_marcher_destroy when the last marcher that needed it goes
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, map->device);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, map->device);
        const size_t field = m->field_bytes ? m->field_bytes : (size_t)map->rows * map->cols * sizeof(float);
        std::lock_guard<std::mutex> lock(g_l2_mu);
        L2Carve &cv = g_l2[map->device];
        size_t cur = 0;
        cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize);
        // only when the whole field fits: pinning a fraction of a larger field measured slightly slower
        if (max_persist > 0 && field <= (size_t)max_persist && field <= (size_t)max_window) {
            if (cv.users == 0) cv.before = cur;
            if (cur >= field || cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, field) == cudaSuccess) {
                m->l2_window_bytes = field;
                m->l2_hit_ratio = 1.0f;
                m->l2_limit_raised = true;
                ++cv.users;
            }
        }
        cudaGetLastError();
    }
    {   // large batches are marched in map order by SM territories (RL_SORT_* / RL_TERRITORY override, for measurements)
        m->sort_poses = !(flags & RL_FLAG_NO_POSE_SORT);
        if (const char *e = std::getenv("RL_SORT_POSES")) {   // 1: every batch of at least RL_SORT_MIN_POSES poses, 0: none
            m->sort_poses = e[0] == '1' && !(flags & RL_FLAG_NO_POSE_SORT);
            m->sort_forced = m->sort_poses;
        }
        int l2_bytes = 0;
        cudaDeviceGetAttribute(&l2_bytes, cudaDevAttrL2CacheSize, map->device);
        m->field_beyond_l2 = (m->field_bytes ? m->field_bytes : (size_t)map->rows * map->cols * sizeof(float)) > (size_t)l2_bytes;
        if (const char *e = std::getenv("RL_SORT_SHIFT")) { const int v = std::atoi(e); if (v >= 0 && v <= 12) m->sort_shift = v; }
        if (const char *e = std::getenv("RL_GATHER_TERRITORIES")) m->gather_territories = e[0] != '0';
        if (const char *e = std::getenv("RL_TERRITORY_IDENTITY")) m->territory_identity = e[0] == '1';
        if (const char *e = std::getenv("RL_SORT_MIN_POSES")) { const long v = std::atol(e); if (v >= 1) m->sort_min_poses = v; }
    }
    if (m->sort_poses) {   // stream-ordered scratch for the sort, kept by the pool between calls
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = map->device;
        if (cudaMemPoolCreate(&m->scratch_pool, &props) == cudaSuccess) {
            uint64_t keep = ~(uint64_t)0;
            cudaMemPoolSetAttribute(m->scratch_pool, cudaMemPoolAttrReleaseThreshold, &keep);
        } else {
            m->scratch_pool = nullptr;   // no pool: batches are marched in the caller's order
            m->sort_poses = false;
        }
        cudaGetLastError();
    }
    cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->stream2, cudaStreamNonBlocking);
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
        if (m->stream2) { cudaStreamSynchronize(m->stream2); cudaStreamDestroy(m->stream2); }
        for (int i = 0; i < 2; ++i) {
            if (m->pipe[i]) { cudaStreamSynchronize(m->pipe[i]); cudaStreamDestroy(m->pipe[i]); }
            if (m->fork_ev[i]) cudaEventDestroy(m->fork_ev[i]);
            if (m->done_ev[i]) cudaEventDestroy(m->done_ev[i]);
        }
        if (m->scratch_pool) { cudaDeviceSynchronize(); cudaMemPoolDestroy(m->scratch_pool); }
        cudaFreeHost(m->h_in); cudaFreeHost(m->h_out);
        cudaFree(m->d_in); cudaFree(m->d_out); cudaFree(m->d_angles); cudaFree(m->d_steps); cudaFree(m->d_field);
        for (int i = 0; i < 32; ++i) {
            if (m->sub_ev[i]) cudaEventDestroy(m->sub_ev[i]);
            if (m->sub_done[i]) cudaEventDestroy(m->sub_done[i]);
        }
        if (m->l2_limit_raised) {
            std::lock_guard<std::mutex> lock(g_l2_mu);
            L2Carve &cv = g_l2[m->map->device];
            if (--cv.users == 0) {   // last user on this device: un-pin the field's lines, give the carve-out back
                cudaCtxResetPersistingL2Cache();
                cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, cv.before);
            }
        }
        cudaGetLastError();
    }
    rl_map_release(m->map);
    delete m;
    return RL_OK;
}

int32_t rl_marcher_set_pipelined(rl_marcher *m, int32_t mode)
{
    if (!m || mode < RL_PIPELINE_OFF || mode > RL_PIPELINE_PDL)
        return rl::fail(RL_ERR_BAD_ARG, "rl_marcher_set_pipelined: bad argument");
    rl::DeviceGuard guard(m->map->device);
    std::lock_guard<std::mutex> lock(m->pipe_mu);
    if (mode == RL_PIPELINE_STREAMS && !m->pipe[0]) {
        for (int i = 0; i < 2; ++i) {
            RL_CUDA(cudaStreamCreateWithFlags(&m->pipe[i], cudaStreamNonBlocking));
            RL_CUDA(cudaEventCreateWithFlags(&m->fork_ev[i], cudaEventDisableTiming));
            RL_CUDA(cudaEventCreateWithFlags(&m->done_ev[i], cudaEventDisableTiming));
        }
    }
    if (m->pipelined == RL_PIPELINE_STREAMS && mode != RL_PIPELINE_STREAMS && (m->done_pending[0] || m->done_pending[1]))
        return rl::fail(RL_ERR_BAD_ARG, "rl_marcher_set_pipelined: call rl_marcher_join before leaving the two-stream mode");
    m->pipelined = mode;
    return RL_OK;
}

int32_t rl_marcher_join(rl_marcher *m, void *stream)
{
    if (!m) return rl::fail(RL_ERR_BAD_ARG, "rl_marcher_join: null marcher");
    rl::DeviceGuard guard(m->map->device);
    std::lock_guard<std::mutex> lock(m->pipe_mu);
    for (int i = 0; i < 2; ++i) {
        if (!m->done_pending[i]) continue;
        RL_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, m->done_ev[i], 0));
        m->done_pending[i] = false;
    }
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
    rl::PipeScope ps(m, (cudaStream_t)stream);
    const int32_t rc = launch_many(m, d_ins, d_outs, n, ps.run, ps.pdl);
    RL_CUDA(ps.finish());
    return rc;
}

int32_t rl_calc_range_fan(rl_marcher *m, const float *d_poses, int64_t pose_stride_rows, float *d_outs,
                          int64_t num_poses, int32_t num_rays, float fov, void *stream)
{
    if (!m || num_poses < 0 || num_rays <= 0 || pose_stride_rows < 1 ||
        (num_poses > 0 && (!d_poses || !d_outs)))
        return rl::fail(RL_ERR_BAD_ARG, "rl_calc_range_fan: bad argument");
    rl::DeviceGuard guard(m->map->device);
    rl::PipeScope ps(m, (cudaStream_t)stream);
    const int32_t rc = launch_pose<true>(m, d_poses, pose_stride_rows, nullptr, d_outs, num_poses, num_rays, fov,
                                         ps.run, ps.pdl);
    RL_CUDA(ps.finish());
    return rc;
}

int32_t rl_calc_range_repeat_angles(rl_marcher *m, const float *d_poses, const float *d_angles,
                                    float *d_outs, int64_t num_poses, int32_t num_angles, void *stream)
{
    if (!m || num_poses < 0 || num_angles < 0 ||
        (num_poses > 0 && num_angles > 0 && (!d_poses || !d_angles || !d_outs)))
        return rl::fail(RL_ERR_BAD_ARG, "rl_calc_range_repeat_angles: bad argument");
    rl::DeviceGuard guard(m->map->device);
    rl::PipeScope ps(m, (cudaStream_t)stream);
    const int32_t rc = launch_pose<false>(m, d_poses, 1, d_angles, d_outs, num_poses, num_angles, 0.0f, ps.run, ps.pdl);
    RL_CUDA(ps.finish());
    return rc;
}

// ---- peer memory for the fused march + all-gather (one process per GPU) ----
int32_t rl_peer_alloc(int32_t device, int64_t bytes, void **d_ptr, uint8_t *handle64)
{
    if (!d_ptr || !handle64 || bytes <= 0) return rl::fail(RL_ERR_BAD_ARG, "rl_peer_alloc: bad argument");
    rl::DeviceGuard guard(device);
    if (!guard.ok) return rl::fail(RL_ERR_NO_DEVICE, "rl_peer_alloc: bad device");
    void *p = nullptr;
    RL_CUDA(cudaMalloc(&p, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return rl::fail(RL_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(handle64, &h, 64);
    *d_ptr = p;
    return RL_OK;
}

int32_t rl_peer_open(int32_t device, const uint8_t *handle64, void **d_ptr)
{
    if (!d_ptr || !handle64) return rl::fail(RL_ERR_BAD_ARG, "rl_peer_open: bad argument");
    rl::DeviceGuard guard(device);
    if (!guard.ok) return rl::fail(RL_ERR_NO_DEVICE, "rl_peer_open: bad device");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    RL_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return RL_OK;
}

int32_t rl_peer_close(int32_t device, void *d_ptr)
{
    rl::DeviceGuard guard(device);
    RL_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return RL_OK;
}

int32_t rl_peer_free(int32_t device, void *d_ptr)
{
    rl::DeviceGuard guard(device);
    RL_CUDA(cudaFree(d_ptr));
    return RL_OK;
}

// Fan march of this rank's poses with every range stored into slot `rank` of the gathered buffer
// of all `world` GPUs: peer_bufs[q] is rank q's buffer (world * slot_rays floats) as mapped in THIS
// process (own buffer for q == rank).  The caller synchronises the ranks afterwards (any
// stream-ordered collective, e.g. a barrier) before anyone reads the gathered ranges.
int32_t rl_calc_range_fan_allgather(rl_marcher *m, const float *d_poses, int64_t pose_stride_rows,
                                    void *const *peer_bufs, int32_t world, int32_t rank, int64_t slot_rays,
                                    int64_t num_poses, int32_t num_rays, float fov, uint32_t flags,
                                    void *stream)
{
    if (!m || !peer_bufs || world < 1 || world > MAX_PEERS || rank < 0 || rank >= world || num_poses < 0 ||
        num_rays <= 0 || pose_stride_rows < 1 || num_poses * num_rays > slot_rays || (num_poses > 0 && !d_poses))
        return rl::fail(RL_ERR_BAD_ARG, "rl_calc_range_fan_allgather: bad argument");
    if (num_poses == 0) return RL_OK;
    rl::DeviceGuard guard(m->map->device);
    PeerOut po{};
    const int32_t rc = fill_peers(po, peer_bufs, world, rank, slot_rays, flags, "rl_calc_range_fan_allgather");
    if (rc != RL_OK) return rc;
    return launch_pose<true>(m, d_poses, pose_stride_rows, nullptr, nullptr, num_poses, num_rays, fov,
                             (cudaStream_t)stream, false, &po);
}

// calc_range_repeat_angles (the particle-filter shape, BASELINE config 3) with the same fused all-gather.
int32_t rl_calc_range_repeat_angles_allgather(rl_marcher *m, const float *d_poses, const float *d_angles,
                                              void *const *peer_bufs, int32_t world, int32_t rank,
                                              int64_t slot_rays, int64_t num_poses, int32_t num_angles,
                                              uint32_t flags, void *stream)
{
    if (!m || !peer_bufs || world < 1 || world > MAX_PEERS || rank < 0 || rank >= world || num_poses < 0 ||
        num_angles <= 0 || num_poses * num_angles > slot_rays || (num_poses > 0 && (!d_poses || !d_angles)))
        return rl::fail(RL_ERR_BAD_ARG, "rl_calc_range_repeat_angles_allgather: bad argument");
    if (num_poses == 0) return RL_OK;
    rl::DeviceGuard guard(m->map->device);
    PeerOut po{};
    const int32_t rc = fill_peers(po, peer_bufs, world, rank, slot_rays, flags, "rl_calc_range_repeat_angles_allgather");
    if (rc != RL_OK) return rc;
    return launch_pose<false>(m, d_poses, 1, d_angles, nullptr, num_poses, num_angles, 0.0f, (cudaStream_t)stream,
                              false, &po);
}

// All-gather of ranges that already exist: d_src[0..n) -> slot `rank` of every gathered buffer.
int32_t rl_allgather_ranges(int32_t device, const float *d_src, void *const *peer_bufs, int32_t world, int32_t rank,
                            int64_t slot_rays, int64_t n, uint32_t flags, void *stream)
{
    if (!peer_bufs || world < 1 || world > MAX_PEERS || rank < 0 || rank >= world || n < 0 || n > slot_rays ||
        (n > 0 && !d_src))
        return rl::fail(RL_ERR_BAD_ARG, "rl_allgather_ranges: bad argument");
    if (n == 0) return RL_OK;
    if ((reinterpret_cast<uintptr_t>(d_src) & 15) || ((int64_t)rank * slot_rays) % 4)
        return rl::fail(RL_ERR_BAD_ARG, "rl_allgather_ranges: source and slot must be 16-byte aligned");
    rl::DeviceGuard guard(device);
    if (!guard.ok) return rl::fail(RL_ERR_NO_DEVICE, "rl_allgather_ranges: bad device");
    PeerOut po{};
    const int32_t rc = fill_peers(po, peer_bufs, world, rank, slot_rays, flags, "rl_allgather_ranges");
    if (rc != RL_OK) return rc;
    for (int q = 0; q < (po.multicast ? 1 : world); ++q)
        if (reinterpret_cast<uintptr_t>(po.buf[q]) & 15) return rl::fail(RL_ERR_BAD_ARG, "rl_allgather_ranges: gathered buffers must be 16-byte aligned");
    const int64_t blocks = ((n + 3) / 4 + 255) / 256;
    if (blocks > 0x7fffffffLL) return rl::fail(RL_ERR_BAD_ARG, "rl_allgather_ranges: too many ranges for one call");
    gather_copy_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_src, n, po);
    RL_CUDA(cudaGetLastError());
    return RL_OK;
}

int32_t rl_calc_range_many_host(rl_marcher *m, const float *ins, float *outs, int64_t n)
{
    if (!m || n < 0 || (n > 0 && (!ins || !outs))) return rl::fail(RL_ERR_BAD_ARG, "rl_calc_range_many_host: bad argument");
    rl::DeviceGuard guard(m->map->device);
    std::lock_guard<std::mutex> lock(m->mu);
    return host_pipeline(m, ins, 3, 3, outs, n, 1,
                         [&](int64_t, int64_t cnt, const float *d_in, float *d_out, cudaStream_t s) {
                             return launch_many(m, d_in, d_out, cnt, s);
                         });
}

int32_t rl_calc_range_fan_host(rl_marcher *m, const float *poses, int64_t pose_stride_rows, float *outs,
                               int64_t num_poses, int32_t num_rays, float fov)
{
    if (!m || num_poses < 0 || num_rays <= 0 || pose_stride_rows < 1 || (num_poses > 0 && (!poses || !outs)))
        return rl::fail(RL_ERR_BAD_ARG, "rl_calc_range_fan_host: bad argument");
    rl::DeviceGuard guard(m->map->device);
    std::lock_guard<std::mutex> lock(m->mu);
    // only row k*pose_stride_rows of each pose's block is meaningful: the pipeline gathers them
    return host_pipeline(m, poses, 3 * pose_stride_rows, 3, outs, num_poses, num_rays,
                         [&](int64_t, int64_t cnt, const float *d_in, float *d_out, cudaStream_t s) {
                             return launch_pose<true>(m, d_in, 1, nullptr, d_out, cnt, num_rays, fov, s);
                         });
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
    RL_CUDA(cudaMemcpy(m->d_angles, angles, (size_t)num_angles * sizeof(float), cudaMemcpyHostToDevice));
    return host_pipeline(m, poses, 3, 3, outs, num_poses, num_angles,
                         [&](int64_t, int64_t cnt, const float *d_in, float *d_out, cudaStream_t s) {
                             return launch_pose<false>(m, d_in, 1, m->d_angles, d_out, cnt, num_angles, 0.0f, s);
                         });
}

}  // extern "C"
