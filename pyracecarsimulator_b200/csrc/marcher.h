// marcher.h -- the marcher handle and the one way every march-type kernel is launched
// (march.cu: plain / gathered scans, car.cu: scan + crash epilogue).
#pragma once
#include <utility>

#include "common.h"

struct rl_marcher {
    const rl_map *map = nullptr;
    rl::MarchParams P{};
    uint32_t flags = 0;
    int sm_count = 148;
    float *d_field = nullptr;    // this marcher's NaN-padded copy of the map's march field (P.pad > 0), or null
    size_t field_bytes = 0;
    // large batches are marched in map order, by SM territories (march.cu: march_territory_kernel)
    bool sort_poses = false;
    bool sort_forced = false;     // RL_SORT_POSES=1: every batch of at least sort_min_poses poses (tests, measurements)
    bool field_beyond_l2 = false;
    bool territory_identity = false;   // RL_TERRITORY_IDENTITY=1: territories over the caller's order, no sort (measurements)
    bool gather_territories = true;   // the fused march + all-gather between 2 GPUs (plain peer stores) takes the territory kernel too (RL_GATHER_TERRITORIES=0: never)
    int sort_shift = 4;           // Morton cells of at least 16 x 16 px (larger when the map has more than 256 of them a side)
    int64_t sort_min_poses = 1;
    cudaMemPool_t scratch_pool = nullptr;   // the sort's stream-ordered scratch (release threshold: never)
    // host-variant staging (guarded by mu)
    std::mutex mu;
    cudaStream_t stream = nullptr, stream2 = nullptr;   // double-buffered H2D -> march -> D2H pipeline
    float *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr, *d_angles = nullptr;
    size_t cap_in = 0, cap_out = 0, cap_hout = 0, cap_angles = 0;  // floats
    cudaEvent_t sub_ev[32] = {}, sub_done[32] = {};   // per sub-chunk: kernel finished / copy landed (created on first use)
    // L2 persistence: the distance field is the one buffer every ray of every call re-reads, so each march
    // launch carries an access-policy window over it (persisting hits) and it stays L2-resident between
    // calls whatever else streams through the cache (0 = window unavailable / RL_FLAG_NO_L2_WINDOW)
    size_t l2_window_bytes = 0;
    float l2_hit_ratio = 1.0f;
    bool l2_limit_raised = false;   // this marcher raised cudaLimitPersistingL2CacheSize (restored on destroy)
    // optional step counter
    bool count = false;
    unsigned long long *d_steps = nullptr;
    // pipelined launches (rl_marcher_set_pipelined): consecutive device-pointer marches alternate between
    // two internal streams so that launch i+1's bulk covers launch i's drain tail
    std::mutex pipe_mu;
    int pipelined = RL_PIPELINE_OFF;
    cudaStream_t pipe[2] = {nullptr, nullptr};
    cudaEvent_t fork_ev[2] = {nullptr, nullptr}, done_ev[2] = {nullptr, nullptr};
    bool done_pending[2] = {false, false};
    int pipe_next = 0;
};

namespace rl {

constexpr int MARCH_CTA_THREADS = 128;   // 128-thread CTAs measured best (profiles/r01_tuning.md section 2)

// Launch with the distance-field access-policy window attached to this launch only (no stream state
// of the caller is touched).  `pdl`: programmatic dependent launch -- the launch may begin once the
// preceding kernel in the stream has let its dependents go (every march kernel does so as its first
// instruction) instead of after it has drained.
template <typename... KArgs, typename... Args>
cudaError_t launch_windowed_ex(const rl_marcher *m, void (*kernel)(KArgs...), unsigned blocks, cudaStream_t s,
                               bool pdl, bool cooperative, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(MARCH_CTA_THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[3];
    unsigned n = 0;
    if (m->l2_window_bytes) {
        attr[n].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[n].val.accessPolicyWindow.base_ptr = const_cast<float *>(m->d_field ? m->d_field : m->P.dist);
        attr[n].val.accessPolicyWindow.num_bytes = m->l2_window_bytes;
        attr[n].val.accessPolicyWindow.hitRatio = m->l2_hit_ratio;
        attr[n].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[n].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        ++n;
    }
    if (pdl) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cooperative) {   // all CTAs co-resident: the kernel uses grid-wide barriers
        attr[n].id = cudaLaunchAttributeCooperative;
        attr[n].val.cooperative = 1;
        ++n;
    }
    cfg.attrs = n ? attr : nullptr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

template <typename... KArgs, typename... Args>
cudaError_t launch_windowed(const rl_marcher *m, void (*kernel)(KArgs...), unsigned blocks, cudaStream_t s,
                            bool pdl, Args &&...args)
{
    return launch_windowed_ex(m, kernel, blocks, s, pdl, false, std::forward<Args>(args)...);
}

// Where a device-pointer march of marcher `m` requested on `caller` actually runs.
//   RL_PIPELINE_OFF      on `caller`, ordinary stream order.
//   RL_PIPELINE_STREAMS  alternately on two internal streams: the launch waits for everything enqueued on
//                        `caller` so far (so it sees its inputs) but NOT for the previous march, and
//                        `caller` is made to wait for the PREVIOUS march's completion only -- the
//                        ranges of a call are therefore valid in `caller`'s order after the NEXT march
//                        call on this marcher or after rl_marcher_join().
//   RL_PIPELINE_PDL      on `caller` with programmatic dependent launch.
struct PipeScope {
    rl_marcher *m;
    cudaStream_t caller, run;
    int slot = -1;
    bool pdl = false;
    cudaError_t err = cudaSuccess;
    PipeScope(rl_marcher *m_, cudaStream_t caller_) : m(m_), caller(caller_), run(caller_)
    {
        if (m->pipelined == RL_PIPELINE_PDL) pdl = true;
        if (m->pipelined != RL_PIPELINE_STREAMS) return;
        m->pipe_mu.lock();
        slot = m->pipe_next;
        run = m->pipe[slot];
        err = cudaEventRecord(m->fork_ev[slot], caller);
        if (err == cudaSuccess) err = cudaStreamWaitEvent(run, m->fork_ev[slot], 0);
    }
    // after the launch: publish its completion, join the previous one into the caller's stream
    cudaError_t finish()
    {
        if (slot < 0) return cudaSuccess;
        cudaError_t e = err;
        if (e == cudaSuccess) e = cudaEventRecord(m->done_ev[slot], run);
        const int other = slot ^ 1;
        if (e == cudaSuccess && m->done_pending[other]) {
            e = cudaStreamWaitEvent(caller, m->done_ev[other], 0);
            m->done_pending[other] = false;
        }
        if (e == cudaSuccess) {
            m->done_pending[slot] = true;
            m->pipe_next = other;
        }
        return e;
    }
    ~PipeScope()
    {
        if (slot >= 0) m->pipe_mu.unlock();
    }
};

}  // namespace rl
