#!/usr/bin/env python
"""bench.py -- rays/s of the batched lidar scan path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

A step is one pass of the hot path (the fork's 4-arg calc_range_many fan: scanMany) over one
batch of synthetic poses.  The headline workload is BASELINE.json configs[1]: 4096 poses x 1080
beams (fov 4.71, max range 300 px) on the 2049^2 stand-in for the missing maps/map.pgm
(synth_map(2049, 1234)), per GPU (weak scaling).

  N = 1   `value` = the march alone, device time per step (CUDA events, L2 flushed and the pinned
          distance field un-pinned before every step).
  N > 1   `value` = north_star's multi-GPU path: every rank marches its own 4096-pose shard against its
          replica of the map AND the ranges are delivered to every GPU -- the all-gather fused into the
          march kernel (rl_calc_range_fan_allgather: multimem.st / peer stores over NVLink) plus the
          barrier that ends it.  The same K steps are also timed without any exchange (`sharded`) and as
          march + NCCL all_gather (`gather_nccl`); `roofline.nvlink` puts the gathered step against the
          bytes every GPU has to receive.  `--gather none|allgather` moves `value` to those variants.

Next to `value`: `steady_state` (K back-to-back launches under ONE event pair, no flush, consecutive
launches overlapped through the marcher's pipelined mode), `e2e` (ScanSimulator2D.scanMany with host
buffers, H2D + kernel + D2H per step, wall clock), `roofline` (HBM per the driver contract, plus the L2
sector and instruction-issue views that actually bound this kernel), `cpu_baseline` (the oracle on the
host cores, N = 1 only) and `configs`: BASELINE.json configs 1, 3, 4 and 5 at their stated shapes,
sharded over the N GPUs (strong scaling), with the range gather where north_star names one.

`--impl reference` times the CPU oracle (the restated range_libc RayMarching; the original is not in
the reference checkout) on all host threads over the SAME pose batches (seeds, sizes) as the native arm.
One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FOV = 4.71
MAX_RANGE_PX = 300
MAP_N, MAP_SEED = 2049, 1234
METRIC = "rays/sec"
N_SETS = 4          # pose batches rotated through the steps
NVLINK_GBS = 900.0  # per direction per GPU, nominal (NVLink 5)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--poses", type=int, default=4096, help="poses per GPU per step")
    ap.add_argument("--beams", type=int, default=1080)
    ap.add_argument("--gather", default="p2p", choices=["p2p", "allgather", "none"],
                    help="N>1, what `value` times: p2p (default) = the march kernel stores into every GPU's gathered "
                         "buffer over NVLink (fused all-gather); allgather = march then NCCL all_gather; none = ranges "
                         "stay sharded.  All three are measured and reported either way.")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs 1/3/4/5 block")
    ap.add_argument("--no-flush", action="store_true", help="keep L2 warm between steps")
    return ap.parse_args()


def workload_name(args):
    return (f"MCTS rollout batch: {args.poses} poses x {args.beams} beams, fov {FOV}, max_range "
            f"{MAX_RANGE_PX}px, synth_map({MAP_N},{MAP_SEED}) stand-in for maps/map.pgm")


def pose_seed(rank, s):
    return 1000 + 17 * rank + s


def build_map_cpu(oracle):
    from pyracecarsimulator_b200 import maps
    img = maps.synth_map(MAP_N, MAP_SEED)
    y = maps.synth_yaml(MAP_N)
    grid = oracle.mapserver_occupancy(img, y.negate, y.occupied_thresh, y.free_thresh)
    occ = oracle.omap_from_grid(grid, True)
    dist = oracle.sqrt_dist2(oracle.edt_exact(occ))
    return y, dist


def time_oracle(marcher, poses, beams, threads, reps):
    out = np.empty(poses.shape[0] * beams, np.float32)
    best = []
    t_w = time.perf_counter()
    while time.perf_counter() - t_w < (3.0 if threads != 1 else 0.5):   # let the host threads spread out
        marcher.calc_range_fan(poses, beams, FOV, outs=out, threads=threads)
    for _ in range(reps):
        t0 = time.perf_counter()
        marcher.calc_range_fan(poses, beams, FOV, outs=out, threads=threads)
        best.append(time.perf_counter() - t0)
    return float(np.median(best)), out


# ------------------------------------------------------------------------------- reference arm
def run_reference(args, rank):
    """The CPU implementation of the path on the host cores, same workload as the native arm: the same
    N_SETS pose batches (rank 0's seeds), args.poses poses x args.beams beams per step, rotated."""
    if rank != 0:
        return
    import oracle
    from pyracecarsimulator_b200 import maps
    y, dist = build_map_cpu(oracle)
    m = oracle.Marcher(dist, MAX_RANGE_PX, y.resolution, y.origin)
    cores = oracle.max_threads()
    pose_sets = [maps.sample_free_poses(dist, args.poses, pose_seed(0, s), y.resolution, y.origin) for s in range(N_SETS)]
    out = np.empty(args.poses * args.beams, np.float32)
    # host threads on these VMs take a second or two of sustained load to spread over the cores
    t_w, n_w = time.perf_counter(), 0
    while n_w < max(1, args.warmup) and time.perf_counter() - t_w < 20.0 or time.perf_counter() - t_w < 3.0:
        m.calc_range_fan(pose_sets[n_w % N_SETS], args.beams, FOV, outs=out, threads=0)
        n_w += 1
    steps = max(1, args.steps)
    t0 = time.perf_counter()
    for i in range(steps):
        m.calc_range_fan(pose_sets[i % N_SETS], args.beams, FOV, outs=out, threads=0)
    dt = time.perf_counter() - t0
    rays = args.poses * args.beams * steps
    value = rays / dt
    sample = (f"the native arm's own batches: {args.poses} poses x {args.beams} beams per step (seeds "
              f"{pose_seed(0, 0)}..{pose_seed(0, N_SETS - 1)}, rotated), {steps} steps, all {cores} host threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "poses_per_gpu": args.poses, "beams": args.beams,
                   "host_threads": cores,
                   "implementation": "oracle/rangelib_oracle.c (CPU restatement of range_libc "
                                     "RayMarching; range_libc itself is not in the reference checkout)",
                   "note": "one host runs one batch per step whatever --gpus is; the speed-up over it depends on the "
                           "host's core count"},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Polls NVML for SM clock and throttle reasons while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def bind_cpus(local_rank, world):
    """Pin this rank to the host cores next to its GPU BEFORE any page-locked buffer is allocated, so that
    the pinned pages land on that NUMA node and the ranks do not migrate across each other.  GPUs that
    report the same affinity mask (one NUMA node for the whole box) split it evenly."""
    info = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        avail = sorted(os.sched_getaffinity(0))
        words = (max(avail) // 64) + 1

        def mask_of(i):
            h = pynvml.nvmlDeviceGetHandleByIndex(i)
            m = pynvml.nvmlDeviceGetCpuAffinity(h, words)
            return frozenset(c for c in avail if (m[c // 64] >> (c % 64)) & 1)

        mine = mask_of(local_rank)
        if not mine:
            return info
        peers = [r for r in range(world) if mask_of(r) == mine] if world > 1 else [local_rank]
        cpus = sorted(mine)
        if len(peers) > 1 and len(cpus) >= len(peers):
            per = len(cpus) // len(peers)
            k = peers.index(local_rank)
            cpus = cpus[k * per:(k + 1) * per]
        os.sched_setaffinity(0, cpus)
        info = {"bound": True, "cpus": f"{cpus[0]}-{cpus[-1]}" if cpus else "", "n_cpus": len(cpus),
                "gpu_affinity_cpus": len(mine), "gpus_sharing_mask": len(peers)}
    except Exception as e:   # noqa: BLE001 - affinity is an optimisation, never a failure
        info["error"] = repr(e)[:120]
    return info


# ------------------------------------------------------------------------------- native arm
class Ctx:
    """Everything the measurement functions share."""


def gpu_map(n, seed, device):
    from pyracecarsimulator_b200 import maps, range_libc
    img = maps.synth_map(n, seed)
    y = maps.synth_yaml(n)
    path = f"/tmp/_rl_bench_map_{n}_{os.getpid()}.pgm"
    maps.write_pgm(path, img)
    y.image = path
    omap = range_libc.PyOMap(y, device=device)
    os.unlink(path)
    return omap, y


def events(torch, n):
    return [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]


def run_native(args, rank, world, local_rank):
    import torch
    from pyracecarsimulator_b200 import _native, maps, range_libc
    from pyracecarsimulator_b200.scan_simulator import ScanSimulator2D

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU path")
    affinity = bind_cpus(local_rank, world)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist_on = world > 1
    # Libraries (NCCL's version / debug lines) write to stdout: point fd 1 at stderr while the benchmark
    # runs and give it back only for the one JSON line.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    tdist = None
    if dist_on:
        import torch.distributed as tdist
        tdist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist_on:
            tdist.barrier()
        torch.cuda.synchronize()

    def reduce_max(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if dist_on:
            tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def reduce_sum(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if dist_on:
            tdist.all_reduce(t, op=tdist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    L = _native.lib()
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def cold_l2(i=0):
        """256 MiB write > 126 MB L2, then -- once that has really happened -- demote the persisting lines
        of the pinned distance field (cudaCtxResetPersistingL2Cache is not stream-ordered)."""
        if flush is None:
            return
        flush.fill_(i & 0xFF)
        torch.cuda.synchronize()
        L.rl_l2_reset_persisting(local_rank)

    c = Ctx()
    c.args, c.rank, c.world, c.local_rank, c.dev, c.torch, c.tdist = args, rank, world, local_rank, dev, torch, tdist
    c.barrier, c.reduce_max, c.reduce_sum, c.cold_l2, c.dist_on = barrier, reduce_max, reduce_sum, cold_l2, dist_on
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        c.hbm_peak, c.peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    else:
        c.hbm_peak, c.peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"

    # ---- map replica on this GPU (ingest kernels), poses for this rank's shard ----
    omap, y = gpu_map(MAP_N, MAP_SEED, local_rank)
    dist_field = omap.dist()
    c.omap2, c.y2, c.dist2 = omap, y, dist_field
    P, B = args.poses, args.beams
    n_rays = P * B
    pose_sets = [maps.sample_free_poses(dist_field, P, pose_seed(rank, s), y.resolution, y.origin) for s in range(N_SETS)]
    d_poses = [torch.from_numpy(p).to(dev) for p in pose_sets]
    d_out = torch.empty(n_rays, dtype=torch.float32, device=dev)
    d_all = torch.empty(world * n_rays, dtype=torch.float32, device=dev) if dist_on else None
    rm = range_libc.PyRayMarchingGPU(omap, MAX_RANGE_PX)
    c.rm2 = rm
    peer = None
    stream_ptr = int(torch.cuda.current_stream(local_rank).cuda_stream)
    if dist_on:
        from pyracecarsimulator_b200.sharded import PeerGather
        peer = PeerGather(local_rank, n_rays)
    mode = args.gather if dist_on else "none"

    def step(i, ev0=None, ev1=None, how=None):
        how = mode if how is None else how
        cold_l2(i)
        if ev0 is not None:
            ev0.record()
        if how == "p2p":
            peer.march(rm, d_poses[i % N_SETS], FOV, B, stream_ptr)   # fused march + all-gather
            peer.sync()
        else:
            rm.calc_range_fan(d_poses[i % N_SETS], d_out, FOV, B)
            if how == "allgather":
                tdist.all_gather_into_tensor(d_all, d_out)
        if ev1 is not None:
            ev1.record()

    K = args.steps

    def timed(how):
        evs = events(torch, K)
        for i in range(3):
            step(i, how=how)
        barrier()
        for i in range(K):
            step(i, *evs[i], how=how)
        barrier()
        return sum(a.elapsed_time(b) for a, b in evs)

    for i in range(args.warmup):
        step(i)
    barrier()
    ev = events(torch, K)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(K):
        step(i, *ev[i])
    barrier()
    t_wall = time.perf_counter() - t_wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    gather_check = None
    if peer is not None:   # correctness of the fused gather: slot r of every buffer == rank r's own scan, twice (both buffer sets)
        ok = True
        for rep in range(2):
            peer.march(rm, d_poses[rep], FOV, B, stream_ptr)
            peer.sync()
            rm.calc_range_fan(d_poses[rep], d_out, FOV, B)
            mine = [torch.empty_like(d_out) for _ in range(world)]
            tdist.all_gather(mine, d_out)
            torch.cuda.synchronize()
            ok = ok and torch.equal(peer.tensor(), torch.cat(mine))
        flag = reduce_sum([0.0 if ok else 1.0])[0]
        if flag:
            raise SystemExit("bench.py: fused p2p gather differs from march + NCCL all_gather")
        gather_check = "bit-identical"
    launches = K  # one march kernel per step (flush fills and NCCL kernels are not ours)

    # ---- N>1: the variants `value` does not time, same K steps each ----
    other_ms = {}
    copy_ms = 0.0
    if dist_on:
        for how in ("none", "p2p", "allgather"):
            other_ms[how] = dev_ms if how == mode else timed(how)
        # what the NVLink path alone takes for one step's bytes: the same stores without the march
        rm.calc_range_fan(d_poses[0], d_out, FOV, B)
        evs = events(torch, K)
        for i in range(3):
            peer.gather(d_out, stream_ptr)
            peer.sync()
        barrier()
        for i in range(K):
            cold_l2(i)
            evs[i][0].record()
            peer.gather(d_out, stream_ptr)
            peer.sync()
            evs[i][1].record()
        barrier()
        copy_ms = sum(a.elapsed_time(b) for a, b in evs)
        ok = torch.equal(peer.tensor()[rank * n_rays:(rank + 1) * n_rays], d_out)
        if reduce_sum([0.0 if ok else 1.0])[0]:
            raise SystemExit("bench.py: rl_allgather_ranges left a wrong slot")

    # ---- steady state: K launches back to back under one event pair, no flush between them ----
    outs4 = [torch.empty(n_rays, dtype=torch.float32, device=dev) for _ in range(N_SETS)]
    want4 = []
    for s in range(N_SETS):
        rm.calc_range_fan(d_poses[s], outs4[s], FOV, B)
        want4.append(outs4[s].clone())
    steady = {}
    for pm in ("off", "streams", "pdl"):
        rm.set_pipelined(pm)
        per = []
        for rep in range(3):
            for o in outs4:
                o.zero_()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(K):
                rm.calc_range_fan(d_poses[i % N_SETS], outs4[i % N_SETS], FOV, B)
            rm.join()
            b.record()
            torch.cuda.synchronize()
            per.append(a.elapsed_time(b) / K)
        rm.set_pipelined("off")
        same = all(torch.equal(o, w) for o, w in zip(outs4, want4))
        steady[pm] = (float(np.median(per)), same)
    del outs4, want4
    steady_ms = reduce_max([steady[pm][0] for pm in ("off", "streams", "pdl")])
    steady_same = reduce_sum([0.0 if steady[pm][1] else 1.0 for pm in ("off", "streams", "pdl")])

    # ---- e2e through the reference-facing API with host buffers ----
    sim = ScanSimulator2D(B, FOV, 0.01, batch_size=P)
    sim.setMap(omap, MAX_RANGE_PX, y.resolution, y.origin)
    sim.setRaytracingMethod("RMGPU")
    for i in range(3):
        sim.scanMany(pose_sets[i % N_SETS])
    e2e_steps = max(3, min(K, 50))
    barrier()
    t0 = time.perf_counter()
    checksum = 0.0
    for i in range(e2e_steps):
        out = sim.scanMany(pose_sets[i % N_SETS])
        checksum += float(out[0])
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.result()

    # ---- the same call with plain (pageable) numpy buffers, as a user of the raw range_libc API passes them:
    # first call staged, later calls into the array the shim page-locked on its second sighting ----
    np_out = np.zeros(n_rays, dtype=np.float32)
    np_poses = [np.array(pose_sets[i % N_SETS], dtype=np.float32) for i in range(2)]
    for i in range(3):
        rm.calc_range_fan(np_poses[i % 2], np_out, FOV, B)
    np_steps = max(3, min(K, 20))
    t0 = time.perf_counter()
    for i in range(np_steps):
        rm.calc_range_fan(np_poses[i % 2], np_out, FOV, B)
        checksum += float(np_out[0])
    np_s = time.perf_counter() - t0
    range_libc.release_host_buffers()

    # ---- the call MCTS actually makes on this batch: checkCollisionMany (scan + isCrashed), host poses in,
    # one int back -- scan and crash test fused, the ranges never leave the GPU ----
    from pyracecarsimulator_b200.racecar import BatchedCar
    car = BatchedCar(device=local_rank)
    car.setCarEdgeDistances(B, -FOV / 2.0, FOV / B, 0.275)
    h_poses = torch.from_numpy(pose_sets[0]).pin_memory()
    d_tmp = torch.empty((P, 3), dtype=torch.float32, device=dev)

    def check_many(i):
        d_tmp.copy_(h_poses, non_blocking=True)
        first, _ = car.scan_crash(rm, d_tmp, 1, P, FOV)
        return int(first.item())

    for i in range(3):
        check_many(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        crash_idx = check_many(i)
    fused_s = time.perf_counter() - t0

    # ---- PCIe ceiling of the e2e path: pinned D2H of one step's ranges, this rank alone and all ranks at once ----
    h_pin = torch.empty(n_rays, dtype=torch.float32, pin_memory=True)

    def d2h_rate(reps=5):
        h_pin.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        tp = time.perf_counter()
        for _ in range(reps):
            h_pin.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        return reps * n_rays * 4 / (time.perf_counter() - tp) / 1e9

    barrier()
    d2h_all = d2h_rate()
    d2h_alone = d2h_all
    if dist_on:
        for r in range(world):
            barrier()
            if r == rank:
                d2h_alone = d2h_rate()
        barrier()
    d2h_sum_all = reduce_sum([d2h_all])[0]
    d2h_min_alone = -reduce_max([-d2h_alone])[0]
    d2h_min_all = -reduce_max([-d2h_all])[0]

    dev_ms, e2e_s, t_wall, ms_none, ms_p2p, ms_nccl, copy_ms = reduce_max(
        [dev_ms, e2e_s, t_wall] + [other_ms.get(h, 0.0) for h in ("none", "p2p", "allgather")] + [copy_ms])

    # ---- roofline inputs (rank 0's launch): algorithmic bytes of ONE launch (4 B/step + 4 B/ray + 12 B/pose) ----
    roofline = None
    cpu_baseline = None
    ingest_ms = omap.ingest_ms
    if rank == 0:
        rm.count_steps(True)
        steps_per_set = []
        for s in range(N_SETS):
            rm.calc_range_fan(d_poses[s], d_out, FOV, B)
            steps_per_set.append(rm.last_steps())
        rm.count_steps(False)
        mean_steps = float(np.mean(steps_per_set))
        alg_bytes = 4.0 * mean_steps + 4.0 * n_rays + 12.0 * P

        def kernel_time(prepare):
            ks = events(torch, min(K, 50))
            for i, (a, b) in enumerate(ks):
                prepare(i)
                a.record()
                rm.calc_range_fan(d_poses[i % N_SETS], d_out, FOV, B)
                b.record()
            torch.cuda.synchronize()
            return float(np.mean([a.elapsed_time(b) for a, b in ks]))

        kernel_ms = kernel_time(cold_l2)                       # L2 flushed, field un-pinned
        warm_kernel_ms = kernel_time(lambda i: None)           # nothing done between launches
        pinned_kernel_ms = kernel_time(lambda i: flush is not None and flush.fill_(i & 0xFF))   # flushed, field left pinned
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        gather_gbs = _native.gather_bandwidth(local_rank, dist_field.nbytes, 64, 10)
        sector_gps = _native.l2_sector_bandwidth(local_rank, dist_field.nbytes, 64, 10)
        tj = {}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
        traffic, warp_insts = tj.get("dram_bytes_per_launch"), tj.get("warp_insts_per_launch")
        lts_bytes, lts_sectors = tj.get("lts_bytes_per_launch"), tj.get("lts_sectors_per_launch")
        sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        issue_peak = 4 * sms * (clocks["sm_mhz"] or 0) * 1e6
        roofline = {"bound": "hbm", "achieved": achieved, "peak": c.hbm_peak, "unit": "GB/s",
                    "frac": achieved / c.hbm_peak, "traffic": traffic, "peak_source": c.peak_src,
                    "kernel": "march_pose_kernel<FAN>", "kernel_ms": kernel_ms,
                    "kernel_ms_warm_l2": warm_kernel_ms,
                    "kernel_ms_flushed_field_pinned": pinned_kernel_ms,
                    "algorithmic_bytes_per_launch": alg_bytes, "march_steps_per_ray": mean_steps / n_rays,
                    "note": "driver-contract view (algorithmic bytes over the HBM copy peak).  The march is an "
                            "L2-resident dependent gather: DRAM traffic is the cold distance field only; `l2` and "
                            "`issue` are the views that bound it",
                    "l2": None if not lts_sectors else {
                        "bound": "L2 gather rate as the SMs see it: random full-sector reads, measured live.  It equals ONE 32-byte sector "
                                 "per clock per SM (SMs x SM clock, `one_sector_per_clk_per_sm`): the rate at which an SM's L1 can "
                                 "send misses to L2 (ncu l1tex__m_l1tex2xbar_req_cycles_active), not an L2-slice limit -- so L1 hits "
                                 "are what relieves it (profiles/r02_territories_ncu.csv)",
                        "one_sector_per_clk_per_sm_gsectors_per_s": sms * (clocks["sm_mhz"] or 0) * 1e-3,
                        "achieved": lts_sectors * 32.0 / (kernel_ms * 1e-3) / 1e9, "peak": sector_gps * 32.0, "unit": "GB/s",
                        "frac": lts_sectors / (kernel_ms * 1e-3) / (sector_gps * 1e9),
                        "lts_sectors_per_launch": lts_sectors, "lts_bytes_per_launch": lts_bytes,
                        "l2_bytes_per_algorithmic_byte": (lts_bytes or lts_sectors * 32.0) / alg_bytes,
                        "peak_gsectors_per_s": sector_gps,
                        "source": tj.get("source"),
                        "note": "traffic = ncu lts__t_sectors.sum of the shipped kernel (profiles/traffic.json, recomputable from "
                                "profiles/); peak = rl_l2_sector_bandwidth over a buffer of the field's size"},
                    "issue": None if not (warp_insts and issue_peak) else {
                        "achieved": warp_insts / (kernel_ms * 1e-3) / 1e9, "peak": issue_peak / 1e9,
                        "unit": "G warp-instructions/s", "frac": warp_insts / (kernel_ms * 1e-3) / issue_peak,
                        "note": "warp instructions per launch (ncu smsp__inst_executed.sum, profiles/traffic.json) over 4 "
                                "issue slots per SM per clock"},
                    "gather4_calibration": {"gbs_at_4B_per_gather": gather_gbs, "gsectors_per_s": gather_gbs / 4.0,
                                            "note": "random 4-byte gathers (1 useful word per fetched sector): the sector rate "
                                                    "an L1-missing scalar gather reaches; NOT a ceiling for the march, whose "
                                                    "warps share sectors and hit L1 (round 1 reported achieved/this = 1.42)"}}
        st_ms = steady_ms[1]
        roofline["at_steady_state"] = {
            "ms_per_launch": st_ms,
            "l2_frac": None if not lts_sectors else lts_sectors / (st_ms * 1e-3) / (sector_gps * 1e9),
            "issue_frac": None if not (warp_insts and issue_peak) else warp_insts / (st_ms * 1e-3) / issue_peak,
            "achieved_algorithmic_gbs": alg_bytes / (st_ms * 1e-3) / 1e9,
            "note": "the same per-launch traffic and instruction counts over the steady-state time per launch (consecutive "
                    "launches overlapped, `steady_state`): what the kernel sustains when scans are issued back to back"}
        if dist_on:
            recv = (world - 1) * n_rays * 4.0
            roofline["nvlink"] = {"bytes_received_per_gpu_per_step": recv, "step_ms": ms_p2p / K,
                                  "achieved": recv / (ms_p2p / K * 1e-3) / 1e9, "peak": NVLINK_GBS, "unit": "GB/s",
                                  "frac": recv / (ms_p2p / K * 1e-3) / 1e9 / NVLINK_GBS,
                                  "floor_ms": recv / (NVLINK_GBS * 1e9) * 1e3,
                                  "stores_only_ms": copy_ms / K,
                                  "frac_of_stores_only": (copy_ms / K) / (ms_p2p / K) if ms_p2p else None,
                                  "stores_only_note": "rl_allgather_ranges: the same NVLink stores of one step's ranges without "
                                                      "the march (+ the same barrier), measured in this run: the practical floor "
                                                      "of the gathered step on this box",
                                  "note": "fused gather step against NVLink ingress: every GPU must receive the other "
                                          f"{world - 1} shards; peak = nominal {NVLINK_GBS:.0f} GB/s per direction"}

        if not args.no_cpu_baseline and world == 1:
            import oracle
            m = oracle.Marcher(dist_field, MAX_RANGE_PX, y.resolution, y.origin)
            cores = oracle.max_threads()
            n1 = min(P, 256)
            t1, ref1 = time_oracle(m, pose_sets[0][:n1], B, 1, 3)
            tn, refn = time_oracle(m, pose_sets[0], B, 0, 3)
            rm.calc_range_fan(d_poses[0], d_out, FOV, B)
            got = d_out.cpu().numpy()
            tol = np.maximum(1e-4 * np.abs(refn), 0.5 * y.resolution)
            cpu_baseline = {"value": n_rays / tn, "unit": "rays/s", "cores": cores, "kind": "port",
                            "sample": f"{P} poses x {B} beams (one full step), all {cores} host threads, median of 3; "
                                      f"single thread on {n1} poses: {n1 * B / t1:.4g} rays/s",
                            "single_thread_value": n1 * B / t1,
                            "parity_vs_gpu": {"bit_identical_frac": float(np.mean(got == refn)),
                                              "within_tolerance": bool(np.all(np.abs(got - refn) <= tol))}}

    if peer is not None:
        peer.close()
        peer = None
    del d_all, sim, h_pin
    configs = None
    if not args.no_configs:
        c.flush = flush
        c.sm_mhz = clocks.get("sm_mhz")
        configs = run_configs(c)

    if rank == 0:
        total_rays = world * n_rays
        line = {
            "metric": METRIC, "value": total_rays * K / (dev_ms * 1e-3), "unit": "rays/s", "n_gpus": world,
            "steps": K, "warmup": args.warmup, "ms_per_step": dev_ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "poses_per_gpu": P, "beams": B,
                       "global_rays_per_step": total_rays, "map": f"{MAP_N}x{MAP_N} fp32 distance field "
                       f"({dist_field.nbytes >> 20} MiB, replicated per GPU)",
                       "parallelism": f"pose-sharded x{world}, map replicated" +
                                      {"none": ", no exchange step (ranges stay on their GPUs)" if dist_on else "",
                                       "allgather": ", NCCL all_gather of ranges inside the step",
                                       "p2p": ", ranges delivered to every GPU inside the step: stored into all gathered buffers "
                                              "over NVLink by the march kernel (fused all-gather) + stream-ordered barrier"}[mode],
                       "l2": "no flush" if args.no_flush else "flushed between steps (256 MiB fill, synchronize, cudaCtxResetPersistingL2Cache: "
                             "the distance field the product pins in L2 is evicted too), outside the per-step events",
                       "timing": "CUDA events per step on the launching stream, summed, max over ranks",
                       "trig": "exact: glibc's sinf/cosf algorithm evaluated per beam on the device (bit parity with the host libm)",
                       "cpu_affinity": affinity},
            "wall_ms_per_step_incl_flush": t_wall / K * 1e3,
            "steady_state": {"value": total_rays / (steady_ms[1] * 1e-3), "unit": "rays/s", "ms_per_launch": steady_ms[1],
                             "mode": "rl_marcher_set_pipelined(RL_PIPELINE_STREAMS): consecutive launches alternate between two "
                                     "internal streams, so launch i+1's bulk covers launch i's drain tail",
                             "bit_identical": steady_same[1] == 0.0, "launches": K,
                             "stream_order": {"value": total_rays / (steady_ms[0] * 1e-3), "ms_per_launch": steady_ms[0],
                                              "bit_identical": steady_same[0] == 0.0},
                             "pdl": {"value": total_rays / (steady_ms[2] * 1e-3), "ms_per_launch": steady_ms[2],
                                     "bit_identical": steady_same[2] == 0.0,
                                     "mode": "RL_PIPELINE_PDL: same stream, programmatic dependent launch"},
                             "note": f"{K} launches under ONE event pair, no L2 flush, {N_SETS} rotating pose batches and output "
                                     "buffers, per rank (ranges stay sharded), max over ranks"},
            "e2e": {"value": total_rays * e2e_steps / e2e_s, "unit": "rays/s",
                    "h2d_bytes_per_step": P * 12, "d2h_bytes_per_step": n_rays * 4, "steps": e2e_steps,
                    "api": "ScanSimulator2D.scanMany(host poses) -> host ranges (pinned), per rank",
                    "pinned_d2h_gbs": d2h_min_all, "pinned_d2h_gbs_rank_alone_min": d2h_min_alone,
                    "pinned_d2h_gbs_all_ranks_sum": d2h_sum_all,
                    "frac_of_d2h_ceiling": (total_rays * e2e_steps / e2e_s) * 4.0 / (d2h_sum_all * 1e9),
                    "numpy_pageable_buffers": {"value": n_rays * np_steps / np_s, "unit": "rays/s per GPU",
                                               "ms_per_call": np_s / np_steps * 1e3,
                                               "api": "PyRayMarchingGPU.calc_range_fan(np.ndarray poses, np.zeros outs): the "
                                                      "caller's array is page-locked on its second sighting (rl_host_register)"},
                    "pcie_bound_rays_per_s": d2h_sum_all * 1e9 / 4.0,
                    "note": "ranges must reach host memory: the step is bound by pinned D2H bandwidth, measured in the same run "
                            "for this rank alone and for all ranks copying at once (the box's aggregate ceiling)"},
            "e2e_fused_crash": {"value": n_rays * e2e_steps / fused_s, "unit": "nominal rays/s per GPU",
                                "ms_per_call": fused_s / e2e_steps * 1e3, "first_crash_index": crash_idx,
                                "api": "checkCollisionMany semantics (scripts/racecar_simulator_v2.py:146-167): host poses in, "
                                       "index of the first crashed pose out; rl_scan_crash skips every pose after it",
                                "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "ingest_ms": ingest_ms,
        }
        if dist_on:
            def entry(ms, note):
                return {"value": total_rays * K / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms / K, "note": note}
            line["sharded"] = entry(ms_none, "ranges left on their GPUs: no exchange step")
            line["gather"] = entry(ms_p2p, "fused: the march kernel stores every range into all GPUs' gathered buffers over "
                                           "NVLink peer memory (rl_calc_range_fan_allgather), barrier after the kernel; "
                                           f"every GPU receives {(world - 1) * n_rays * 4 / 1e6:.0f} MB per step")
            line["gather_nccl"] = entry(ms_nccl, "march, then NCCL all_gather_into_tensor of the ranges")
            line["gather_check"] = gather_check
            line["gather_backend"] = c.gather_backend if hasattr(c, "gather_backend") else None
            line["gather_mode"] = os.environ.get("RL_GATHER_MODE", "auto (peer stores between 2 GPUs, NVLS multicast from 3 up)")
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        if configs is not None:
            line["configs"] = configs
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if dist_on:
        tdist.barrier()
        tdist.destroy_process_group()


# ------------------------------------------------------------------------------- BASELINE configs 1, 3, 4, 5
def median_ms(c, fn, reps=3, warm=1, flush=True):
    """Device time of fn() (CUDA events on the current stream): median of `reps`, max over ranks."""
    torch = c.torch
    for _ in range(warm):
        fn()
    ts = []
    for i in range(reps):
        if flush:
            c.cold_l2(i)
        c.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return c.reduce_max([float(np.median(ts))])[0]


def counter_views(c, cfg_key, launches, ms):
    """Issue / SM->L2 request / DRAM views of a config at N = 1: ncu counters of one launch of this shape
    (profiles/traffic_configs.json, by territories, sort included) x `launches`, over the time measured here."""
    path = os.path.join(ROOT, "profiles", "traffic_configs.json")
    if c.world != 1 or not os.path.exists(path):
        return None
    tj = json.load(open(path)).get(cfg_key)
    if not tj:
        return None
    torch = c.torch
    sms = torch.cuda.get_device_properties(c.local_rank).multi_processor_count
    mhz = float(getattr(c, "sm_mhz", 0.0) or 1965.0)   # the SM clock sampled under load in the timed region of this run
    t, srt, plain = tj["territories"], tj["sort"], tj["caller_order"]
    insts = (t["warp_insts"] + srt["warp_insts"]) * launches
    l2 = (t["l2_read_sectors_from_sm"] + srt["l2_read_sectors_from_sm"]) * launches
    dram = (t["dram_bytes_read"] + t["dram_bytes_written"] + srt["dram_bytes_read"] + srt["dram_bytes_written"]) * launches
    sec = ms * 1e-3
    return {"issue_frac": insts / sec / (4 * sms * mhz * 1e6), "sm_to_l2_request_frac": l2 / sec / (sms * mhz * 1e6),
            "dram_gbs": dram / sec / 1e9, "dram_frac": dram / sec / 1e9 / c.hbm_peak,
            "l1_hit_rate": t["l1_sector_hits"] / t["l1_sectors_requested"],
            "l1_hit_rate_callers_order": plain["l1_sector_hits"] / plain["l1_sectors_requested"],
            "dram_bytes_per_launch": {"territories": t["dram_bytes_read"] + t["dram_bytes_written"],
                                      "callers_order": plain["dram_bytes_read"] + plain["dram_bytes_written"]},
            "sm_clock_mhz": mhz,
            "note": "per-launch ncu counters of the map-order path (profiles/traffic_configs.json, from profiles/r02_territories_ncu.csv) "
                    "over the time measured in this run; issue = 4 warp instructions per SM per clock, SM->L2 = one sector per SM per "
                    "clock (DESIGN.md 4a): the kernel is issue-bound once its field cells come from L1"}


def run_configs(c):
    out = {"note": "BASELINE.json configs at their stated shapes; total work fixed and sharded over the N GPUs (strong "
                   "scaling), device time = median of 3 after a warm-up, max over ranks; algorithmic bytes = 4 B per march "
                   "step (counted on the device) + 4 B per stored range + 12 B per pose; frac_hbm = algorithmic bytes / "
                   "kernel time / measured HBM copy peak"}
    for name, fn in (("config1", config1), ("config3", config3), ("config4", config4), ("config5", config5)):
        try:
            out[name] = fn(c)
        except Exception as e:   # noqa: BLE001 - a failing side config must not take the headline down
            import traceback
            traceback.print_exc()
            out[name] = {"error": repr(e)[:300]}
        c.barrier()
    return out


def config1(c):
    """maps/colombia single-pose 1080-beam scan through ScanSimulator2D.scan (host floats in, host ranges out)."""
    if c.rank != 0:
        return None
    from pyracecarsimulator_b200 import maps, range_libc
    from pyracecarsimulator_b200.scan_simulator import ScanSimulator2D
    z = np.load(os.path.join(ROOT, "tests", "golden", "colombia_map.npz"))
    path = f"/tmp/_rl_bench_colombia_{os.getpid()}.pgm"
    maps.write_pgm(path, z["img"])
    yc = maps.MapYaml(path, float(z["resolution"]), tuple(float(v) for v in z["origin"]))
    omap = range_libc.PyOMap(yc, device=c.local_rank)
    os.unlink(path)
    sim = ScanSimulator2D(1080, FOV, 0.01, batch_size=200)
    sim.setMap(omap, MAX_RANGE_PX, yc.resolution, yc.origin)
    sim.setRaytracingMethod("RMGPU")
    for _ in range(20):
        sim.scan(0.275, 0.0, 0.0)
    n = 300
    t0 = time.perf_counter()
    for _ in range(n):
        sim.scan(0.275, 0.0, 0.0)
    us = (time.perf_counter() - t0) / n * 1e6
    return {"workload": "maps/colombia, 1 pose x 1080 beams, ScanSimulator2D.scan (host in, host out), one GPU",
            "us_per_scan": us, "rays_per_s": 1080 / (us * 1e-6), "ingest_ms": omap.ingest_ms,
            "note": "latency case: the reference's real-time budget is 50 000 us per scan (20 Hz)"}


def config3(c):
    """Particle filter: 1M poses x 60 angles (calc_range_repeat_angles) on the 2049^2 stand-in for maps/big-teach."""
    torch = c.torch
    from pyracecarsimulator_b200 import maps
    from pyracecarsimulator_b200.sharded import PeerGather, shard_bounds
    n_total, A = 1_000_000, 60
    lo, hi = shard_bounds(n_total, c.world, c.rank)
    per = -(-n_total // c.world)
    poses = maps.sample_free_poses(c.dist2, n_total, 303, c.y2.resolution, c.y2.origin)[lo:hi]
    d_p = torch.from_numpy(np.ascontiguousarray(poses)).to(c.dev)
    angles = torch.from_numpy(np.linspace(-FOV / 2, FOV / 2, A, endpoint=False).astype(np.float32)).to(c.dev)
    out = torch.empty(max(1, (hi - lo) * A), dtype=torch.float32, device=c.dev)
    rm = c.rm2
    ms = median_ms(c, lambda: rm.calc_range_repeat_angles(d_p, angles, out))
    rm.count_steps(True)
    rm.calc_range_repeat_angles(d_p, angles, out)
    steps = c.reduce_sum([float(rm.last_steps())])[0]
    rm.count_steps(False)
    rays = n_total * A
    alg = 4.0 * steps + 4.0 * rays + 12.0 * n_total
    res = {"workload": f"1M poses x 60 angles, calc_range_repeat_angles, synth_map({MAP_N},{MAP_SEED}), sharded x{c.world}",
           "rays": rays, "kernel_ms": ms, "rays_per_s": rays / (ms * 1e-3), "march_steps_per_ray": steps / rays,
           "algorithmic_bytes": alg, "achieved_gbs": alg / c.world / (ms * 1e-3) / 1e9,
           "frac_hbm": alg / c.world / (ms * 1e-3) / 1e9 / c.hbm_peak,
           "bound": "L2-resident gather (16 MiB field): latency / issue, not HBM",
           "order": "map order by SM territories when this rank's share is dense and large enough (>= one pose per 16 map cells "
                    "and >= 24 M rays: 1 GPU and 2 GPUs here), the caller's order otherwise"}
    views = counter_views(c, "config3", 1, ms)
    if views:
        res["counters"] = views
    if c.dist_on:
        peer = PeerGather(c.local_rank, per * A)
        sp = int(torch.cuda.current_stream(c.local_rank).cuda_stream)

        def fused():
            peer.march_angles(rm, d_p, angles, sp)
            peer.sync()

        gms = median_ms(c, fused)
        fused()
        rm.calc_range_repeat_angles(d_p, angles, out)
        torch.cuda.synchronize()
        own = peer.tensor()[c.rank * per * A: c.rank * per * A + (hi - lo) * A]
        ok = c.reduce_sum([0.0 if torch.equal(own, out[:(hi - lo) * A]) else 1.0])[0] == 0.0
        recv = (c.world - 1) * per * A * 4.0
        res["with_gather"] = {"ms": gms, "rays_per_s": rays / (gms * 1e-3), "backend": peer.backend,
                              "multicast": peer.multicast, "own_slot_check": "bit-identical" if ok else "MISMATCH",
                              "bytes_received_per_gpu": recv, "nvlink_floor_ms": recv / (NVLINK_GBS * 1e9) * 1e3,
                              "nvlink_frac": recv / (gms * 1e-3) / 1e9 / NVLINK_GBS,
                              "api": "rl_calc_range_repeat_angles_allgather (fused) + barrier"}
        c.gather_backend = peer.backend + (" + NVLS multicast (multimem.st)" if peer.multicast else "")
        peer.close()
    return res


def config4(c):
    """Fused rollout: 65536 cars x 50 bicycle-model steps, a 1080-beam scan per step, maps/colombia."""
    torch = c.torch
    from pyracecarsimulator_b200 import maps, range_libc
    from pyracecarsimulator_b200.racecar import BatchedCar
    from pyracecarsimulator_b200.sharded import ShardedRollout, gpu_rollout_fn, shard_bounds
    z = np.load(os.path.join(ROOT, "tests", "golden", "colombia_map.npz"))
    path = f"/tmp/_rl_bench_colombia4_{os.getpid()}.pgm"
    maps.write_pgm(path, z["img"])
    yc = maps.MapYaml(path, float(z["resolution"]), tuple(float(v) for v in z["origin"]))
    omap = range_libc.PyOMap(yc, device=c.local_rank)
    os.unlink(path)
    rm = range_libc.PyRayMarchingGPU(omap, MAX_RANGE_PX)
    car = BatchedCar(device=c.local_rank)
    car.setCarEdgeDistances(1080, -FOV / 2.0, FOV / 1080, 0.275)
    ncars, steps, R = 65536, 50, 1080
    start = maps.sample_free_poses(omap.dist(), ncars, 404, yc.resolution, yc.origin, min_clear_px=6.0)
    s0 = np.zeros((ncars, 11))
    s0[:, :3] = start
    s0[:, 3] = 2.0
    states = torch.from_numpy(s0).to(c.dev)
    sr = ShardedRollout(gpu_rollout_fn(car, rm, FOV), c.dev)
    res = {}

    def roll():
        res["crash"], res["reward"] = sr.rollout(states, steps, seed=42)

    ms = median_ms(c, roll, flush=False)
    crash = res["crash"]
    needed = int(torch.where(crash >= 0, crash + 1, torch.full_like(crash, steps)).sum().item())
    # algorithmic bytes: 4 B per march step of the poses that had to be scanned (counted by re-scanning this rank's
    # needed poses with the step counter on) + 12 B per scanned pose; no range is stored
    lo, hi = shard_bounds(ncars, c.world, c.rank)
    st = states[lo:hi].clone()
    o = car.rollout(rm, st, None, steps, FOV, seed=42, car_offset=lo)
    cr = o["crash_index"]
    last = torch.where(cr >= 0, cr, torch.full_like(cr, steps - 1))
    buf = torch.empty((hi - lo) * R, dtype=torch.float32, device=c.dev)
    rm.count_steps(True)
    for s in range(steps):
        sel = o["poses"][s][last >= s].contiguous()
        if sel.shape[0]:
            rm.calc_range_fan(sel, buf, FOV, R)
    msteps = c.reduce_sum([float(rm.last_steps())])[0]
    rm.count_steps(False)
    alg = 4.0 * msteps + 12.0 * needed
    nominal = ncars * steps * R
    return {"workload": f"65536 cars x 50 steps x 1080 beams on maps/colombia, cars sharded x{c.world} (ShardedRollout), "
                        "action schedule drawn on the device (Philox seed 42), all-gather of (crash_index, reward) inside",
            "nominal_rays": nominal, "kernel_ms": ms, "nominal_rays_per_s": nominal / (ms * 1e-3),
            "crashed_frac": float((crash >= 0).float().mean().item()), "poses_needed": needed, "poses_total": ncars * steps,
            "rays_needed": needed * R, "rays_needed_per_s": needed * R / (ms * 1e-3),
            "algorithmic_bytes": alg, "achieved_gbs": alg / c.world / (ms * 1e-3) / 1e9,
            "frac_hbm": alg / c.world / (ms * 1e-3) / 1e9 / c.hbm_peak,
            "launches_per_rollout": 4, "bound": "L2-resident gather (0.6 MiB field): latency / issue, not HBM",
            "note": "kernel_ms covers rl_rollout_actions + car_rollout_kernel + march_crash_kernel + finalize and the 12 B/car "
                    "gather; rays after a car's first crash are skipped, nominal counts them"}


def config5(c):
    """Synthetic 8192^2 map (256 MiB fp32 field: NOT L2-resident), 16M poses x 270 beams, map replicated,
    ranges gathered to every GPU -- in pieces, through two alternating peer buffer sets, so that the 17.3 GB of
    ranges never need one buffer."""
    torch = c.torch
    from pyracecarsimulator_b200 import maps, range_libc
    from pyracecarsimulator_b200.sharded import PeerGather, shard_bounds
    n_total, R, n_map, seed = 16_000_000, 270, 8192, 5678
    omap, y = gpu_map(n_map, seed, c.local_rank)
    dist = omap.dist()
    rm = range_libc.PyRayMarchingGPU(omap, MAX_RANGE_PX)
    lo, hi = shard_bounds(n_total, c.world, c.rank)
    per = -(-n_total // c.world)
    mine = hi - lo
    # every rank draws only its own poses (seeded per rank): 16M x 12 B would be 192 MB of host work per rank
    poses = maps.sample_free_poses(dist, mine, 505 + 31 * c.rank, y.resolution, y.origin)
    del dist
    d_p = torch.from_numpy(poses).to(c.dev)
    chunk = min(per, 1 << 18)                 # gathered pieces: 262 144 poses x 270 beams x 4 B = 283 MB per rank and piece
    n_chunks = -(-per // chunk)
    chunk_s = min(per, 1 << 21)               # sharded pieces: 2 M poses (2.2 GB of ranges), two alternating buffers
    n_chunks_s = -(-per // chunk_s)
    sp = int(torch.cuda.current_stream(c.local_rank).cuda_stream)
    ring = [torch.empty(chunk_s * R, dtype=torch.float32, device=c.dev) for _ in range(2)]

    def sharded():
        for k in range(n_chunks_s):
            a, b = min(k * chunk_s, mine), min((k + 1) * chunk_s, mine)
            if b > a:
                rm.calc_range_fan(d_p[a:b], ring[k & 1], FOV, R)

    ms = median_ms(c, sharded, flush=False)
    rm.count_steps(True)
    sharded()
    msteps = c.reduce_sum([float(rm.last_steps())])[0]
    rm.count_steps(False)
    rays = n_total * R
    alg = 4.0 * msteps + 4.0 * rays + 12.0 * n_total
    res = {"workload": f"synth_map({n_map},{seed}) (256 MiB fp32 field, replicated), 16M poses x 270 beams sharded x{c.world}, "
                       f"marched in {n_chunks_s} pieces of {chunk_s} poses per rank, each piece in map order by SM territories "
                       "(march_territory_kernel: the field is twice the L2, neighbouring poses share their field cells in L1 / L2)",
           "rays": rays, "kernel_ms": ms, "rays_per_s": rays / (ms * 1e-3), "march_steps_per_ray": msteps / rays,
           "algorithmic_bytes": alg, "achieved_gbs": alg / c.world / (ms * 1e-3) / 1e9,
           "frac_hbm": alg / c.world / (ms * 1e-3) / 1e9 / c.hbm_peak, "ingest_ms": omap.ingest_ms,
           "bound": "the one HBM-side case: the field is twice the L2, misses are 32-byte sector gathers from HBM (3.3 TB/s of "
                    "them in caller order, profiles/r02_cfg35_metrics.csv; map order + SM territories turn most into L1 / L2 hits)"}
    views = counter_views(c, "config5_share", n_total / 2.0e6, ms)   # counters were taken on a 2 M-pose piece
    if views:
        res["counters"] = views
    if c.dist_on:
        peer = PeerGather(c.local_rank, chunk * R, nbuf=2)

        def fused():
            for k in range(n_chunks):
                a, b = min(k * chunk, mine), min((k + 1) * chunk, mine)
                if b > a:
                    peer.march(rm, d_p[a:b], FOV, R, sp)
                else:
                    peer._take(None)
                peer.sync()

        gms = median_ms(c, fused, flush=False)
        # correctness: the last piece as gathered on this GPU == every rank's own scan of that piece (checksums)
        fused()
        k = n_chunks - 1
        a, b = min(k * chunk, mine), min((k + 1) * chunk, mine)
        rm.calc_range_fan(d_p[a:b], ring[0], FOV, R)
        own = ring[0][:(b - a) * R]
        sums = torch.zeros(c.world, dtype=torch.float64, device=c.dev)
        sums[c.rank] = own.double().sum()
        c.tdist.all_reduce(sums)
        got = peer.tensor().view(c.world, chunk * R)[:, :(b - a) * R].double().sum(dim=1)
        ok = bool(torch.equal(got, sums)) and bool(torch.equal(peer.tensor().view(c.world, chunk * R)[c.rank, :(b - a) * R], own))
        ok = c.reduce_sum([0.0 if ok else 1.0])[0] == 0.0
        recv = (c.world - 1) * per * R * 4.0
        res["with_gather"] = {"ms": gms, "rays_per_s": rays / (gms * 1e-3), "backend": peer.backend, "multicast": peer.multicast,
                              "pieces": n_chunks, "piece_bytes_per_rank": chunk * R * 4, "check": "bit-identical own slot + "
                              "per-rank checksums" if ok else "MISMATCH",
                              "bytes_received_per_gpu": recv, "nvlink_floor_ms": recv / (NVLINK_GBS * 1e9) * 1e3,
                              "nvlink_frac": recv / (gms * 1e-3) / 1e9 / NVLINK_GBS,
                              "api": "rl_calc_range_fan_allgather per piece into two alternating peer buffer sets + barrier"}
        peer.close()
    return res


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_native(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
