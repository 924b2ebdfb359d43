#!/usr/bin/env python
"""bench.py -- rays/s of the batched lidar scan path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

A step is one pass of the hot path (the fork's 4-arg calc_range_many fan: scanMany) over one
batch of synthetic poses.  At N=1 the workload is BASELINE.json configs[1]: 4096 poses x 1080
beams (fov 4.71, max range 300 px) on the 2049^2 stand-in for the missing maps/map.pgm
(synth_map(2049, 1234)).  For N>1 every rank marches its own 4096-pose
shard against its own replica of the map (weak scaling; rays are independent, so the path has no
exchange step and `value` times the sharded march).  The optional delivery of all ranges to every
GPU -- north_star's "final gather of ranges over NVLink" -- is timed in the same run and reported
under `gather` (fused: the march kernel stores into every GPU's buffer over NVLink peer memory)
and `gather_nccl` (march + NCCL all_gather); `--gather p2p|allgather` puts it inside `value`.

One JSON line is printed by rank 0; keys follow the driver contract plus `roofline` and
`cpu_baseline`.  `value` is device time (CUDA events per step on the launching stream, L2
flushed between steps, max over ranks); `e2e` is the same metric through
ScanSimulator2D.scanMany with host buffers (H2D + kernel + D2H per step, wall clock).
`--impl reference` times the CPU oracle (the restated range_libc RayMarching; the original is
not in the reference checkout) on all host threads instead.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--poses", type=int, default=4096, help="poses per GPU per step")
    ap.add_argument("--beams", type=int, default=1080)
    ap.add_argument("--gather", default="none", choices=["p2p", "allgather", "none"],
                    help="N>1, what `value` times: none = ranges stay sharded (default); p2p = march kernel stores into "
                         "every GPU's gathered buffer over NVLink (fused); allgather = march then NCCL all_gather. "
                         "The other variants are still measured and reported under `gather` / `gather_nccl`.")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="keep L2 warm between steps")
    return ap.parse_args()


def workload_name(args):
    return (f"MCTS rollout batch: {args.poses} poses x {args.beams} beams, fov {FOV}, max_range "
            f"{MAX_RANGE_PX}px, synth_map({MAP_N},{MAP_SEED}) stand-in for maps/map.pgm")


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
    if rank != 0:
        return
    import oracle
    from pyracecarsimulator_b200 import maps
    y, dist = build_map_cpu(oracle)
    m = oracle.Marcher(dist, MAX_RANGE_PX, y.resolution, y.origin)
    cores = oracle.max_threads()
    sample_poses = min(args.poses, 512)
    poses = maps.sample_free_poses(dist, sample_poses, 4242, y.resolution, y.origin)
    out = np.empty(sample_poses * args.beams, np.float32)
    # host threads on these VMs take a second or two of sustained load to spread over the cores
    t_w, n_w = time.perf_counter(), 0
    while n_w < max(1, args.warmup) or time.perf_counter() - t_w < 3.0:
        m.calc_range_fan(poses, args.beams, FOV, outs=out, threads=0)
        n_w += 1
    steps = max(1, args.steps)   # one step = the 512-pose sample: a few ms on a multi-core host
    t0 = time.perf_counter()
    for _ in range(steps):
        m.calc_range_fan(poses, args.beams, FOV, outs=out, threads=0)
    dt = time.perf_counter() - t0
    rays = sample_poses * args.beams * steps
    value = rays / dt
    sample = f"{sample_poses} of {args.poses} poses x {args.beams} beams per step, {steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(args), "host_threads": cores,
                   "implementation": "oracle/rangelib_oracle.c (CPU restatement of range_libc "
                                     "RayMarching; range_libc itself is not in the reference checkout)"},
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


# ------------------------------------------------------------------------------- native arm
def run_native(args, rank, world, local_rank):
    import torch
    from pyracecarsimulator_b200 import _native, maps, range_libc
    from pyracecarsimulator_b200.scan_simulator import ScanSimulator2D

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist_on = world > 1
    # Libraries (NCCL's version / debug lines) write to stdout: point fd 1 at stderr while the benchmark
    # runs and give it back only for the one JSON line.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if dist_on:
        import torch.distributed as tdist
        tdist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist_on:
            tdist.barrier()
        torch.cuda.synchronize()

    # ---- map replica on this GPU (ingest kernels), poses for this rank's shard ----
    img = maps.synth_map(MAP_N, MAP_SEED)
    y = maps.synth_yaml(MAP_N)
    path = f"/tmp/_rl_bench_map_{os.getpid()}.pgm"
    maps.write_pgm(path, img)
    y.image = path
    omap = range_libc.PyOMap(y, device=local_rank)
    os.unlink(path)
    dist_field = omap.dist()
    P, B = args.poses, args.beams
    n_rays = P * B
    n_sets = 4  # rotate pose batches so consecutive steps do not repeat the same rays
    pose_sets = [maps.sample_free_poses(dist_field, P, 1000 + 17 * rank + s, y.resolution, y.origin)
                 for s in range(n_sets)]
    d_poses = [torch.from_numpy(p).to(dev) for p in pose_sets]
    d_out = torch.empty(n_rays, dtype=torch.float32, device=dev)
    d_all = torch.empty(world * n_rays, dtype=torch.float32, device=dev) if dist_on else None
    rm = range_libc.PyRayMarchingGPU(omap, MAX_RANGE_PX)
    peer = None
    if dist_on:
        from pyracecarsimulator_b200.sharded import PeerGather
        peer = PeerGather(local_rank, n_rays)
        stream_ptr = int(torch.cuda.current_stream(local_rank).cuda_stream)
    mode = args.gather if dist_on else "none"
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(i, ev0=None, ev1=None, how=None):
        how = mode if how is None else how
        if flush is not None:
            flush.fill_(i & 0xFF)          # 256 MiB write > 126 MB L2 ...
            _native.lib().rl_l2_reset_persisting(local_rank)   # ... and un-pin the distance field, so it is evicted too
        if ev0 is not None:
            ev0.record()
        if how == "p2p":
            peer.march(rm, d_poses[i % n_sets], FOV, B, stream_ptr)   # fused march + all-gather
            peer.sync()
        else:
            rm.calc_range_fan(d_poses[i % n_sets], d_out, FOV, B)
            if how == "allgather":
                tdist.all_gather_into_tensor(d_all, d_out)
        if ev1 is not None:
            ev1.record()

    def timed(how):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for i in range(3):
            step(i, how=how)
        barrier()
        for i in range(K):
            step(i, *evs[i], how=how)
        barrier()
        return sum(a.elapsed_time(b) for a, b in evs)

    K = args.steps
    for i in range(args.warmup):
        step(i)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(K):
        step(i, *ev[i])
    barrier()
    t_wall = time.perf_counter() - t_wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    if peer is not None:   # correctness of the fused gather: slot r of every buffer == rank r's own scan
        peer.march(rm, d_poses[0], FOV, B, stream_ptr)
        peer.sync()
        rm.calc_range_fan(d_poses[0], d_out, FOV, B)
        mine = [torch.empty_like(d_out) for _ in range(world)]
        tdist.all_gather(mine, d_out)
        torch.cuda.synchronize()
        if not torch.equal(peer.tensor(), torch.cat(mine)):
            raise SystemExit("bench.py: fused p2p gather differs from march + NCCL all_gather")
    launches = K  # one march kernel per step (flush fills and NCCL kernels are not ours)

    # ---- N>1: the variants `value` does not time, same K steps each ----
    other_ms = {}
    if dist_on:
        for how in ("none", "p2p", "allgather"):
            other_ms[how] = dev_ms if how == mode else timed(how)

    # ---- e2e through the reference-facing API with host buffers ----
    sim = ScanSimulator2D(B, FOV, 0.01, batch_size=P)
    sim.setMap(omap, MAX_RANGE_PX, y.resolution, y.origin)
    sim.setRaytracingMethod("RMGPU")
    for i in range(3):
        sim.scanMany(pose_sets[i % n_sets])
    e2e_steps = max(3, min(K, 50))
    barrier()
    t0 = time.perf_counter()
    checksum = 0.0
    for i in range(e2e_steps):
        out = sim.scanMany(pose_sets[i % n_sets])
        checksum += float(out[0])
    barrier()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.result()

    # ---- the same call with plain (pageable) numpy buffers, as a user of the raw range_libc API passes them:
    # first call staged, later calls into the array the shim page-locked on its second sighting ----
    np_out = np.zeros(n_rays, dtype=np.float32)
    np_poses = [np.array(pose_sets[i % n_sets], dtype=np.float32) for i in range(2)]
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

    # ---- max over ranks ----
    # PCIe ceiling of the e2e path: pinned D2H of one step's ranges
    h_pin = torch.empty(n_rays, dtype=torch.float32, pin_memory=True)
    h_pin.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    tp = time.perf_counter()
    for _ in range(5):
        h_pin.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    d2h_gbs = 5 * n_rays * 4 / (time.perf_counter() - tp) / 1e9

    t = torch.tensor([dev_ms, e2e_s, t_wall] + [other_ms.get(h, 0.0) for h in ("none", "p2p", "allgather")],
                     dtype=torch.float64, device=dev)
    if dist_on:
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
    dev_ms, e2e_s, t_wall, ms_none, ms_p2p, ms_nccl = (float(v) for v in t.tolist())

    if rank == 0:
        # ---- roofline inputs: algorithmic bytes of ONE launch (4 B/step + 4 B/ray + 12 B/pose) ----
        rm.count_steps(True)
        steps_per_set = []
        for s in range(n_sets):
            rm.calc_range_fan(d_poses[s], d_out, FOV, B)
            steps_per_set.append(rm.last_steps())
        rm.count_steps(False)
        mean_steps = float(np.mean(steps_per_set))
        alg_bytes = 4.0 * mean_steps + 4.0 * n_rays + 12.0 * P
        # kernel-only time: same loop, gather off, events around the launch alone
        k_ms = []
        for i in range(min(K, 50)):
            if flush is not None:
                flush.fill_(i & 0xFF)
                _native.lib().rl_l2_reset_persisting(local_rank)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rm.calc_range_fan(d_poses[i % n_sets], d_out, FOV, B)
            b.record()
            k_ms.append((a, b))
        torch.cuda.synchronize()
        kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in k_ms]))
        warm_ms = []
        for i in range(min(K, 50)):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rm.calc_range_fan(d_poses[i % n_sets], d_out, FOV, B)
            b.record()
            warm_ms.append((a, b))
        torch.cuda.synchronize()
        warm_kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in warm_ms]))
        pinned_ms = []
        for i in range(min(K, 50)):   # L2 flushed, distance field left pinned by the product's access-policy window
            if flush is not None:
                flush.fill_(i & 0xFF)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rm.calc_range_fan(d_poses[i % n_sets], d_out, FOV, B)
            b.record()
            pinned_ms.append((a, b))
        torch.cuda.synchronize()
        pinned_kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in pinned_ms]))
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            hbm_peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
        else:
            hbm_peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        gather_gbs = _native.gather_bandwidth(local_rank, dist_field.nbytes, 64, 10)
        traffic, warp_insts = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic, warp_insts = tj.get("dram_bytes_per_launch"), tj.get("warp_insts_per_launch")
        roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                    "kernel": "march_pose_kernel<FAN>", "kernel_ms": kernel_ms,
                    "kernel_ms_warm_l2": warm_kernel_ms,
                    "kernel_ms_flushed_field_pinned": pinned_kernel_ms,
                    "algorithmic_bytes_per_launch": alg_bytes, "march_steps_per_ray": mean_steps / n_rays,
                    "note": "the march is an L2-resident gather, not an HBM stream; l2_gather is the bound "
                            "BASELINE.json names",
                    "issue": None if not (warp_insts and clocks.get("sm_mhz")) else {
                        "achieved": warp_insts / (kernel_ms * 1e-3) / 1e9,
                        "peak": 4 * torch.cuda.get_device_properties(local_rank).multi_processor_count * clocks["sm_mhz"] * 1e6 / 1e9,
                        "unit": "G warp-instructions/s",
                        "frac": warp_insts / (kernel_ms * 1e-3) / (4 * torch.cuda.get_device_properties(local_rank).multi_processor_count * clocks["sm_mhz"] * 1e6),
                        "note": "what actually bounds the kernel: warp instructions per launch (ncu smsp__inst_executed.sum, "
                                "profiles/traffic.json) over 4 issue slots per SM per clock"},
                    "l2_gather": {"achieved": achieved, "peak": gather_gbs, "unit": "GB/s",
                                  "frac": achieved / gather_gbs,
                                  "peak_source": "rl_gather_bandwidth: random 4-B gathers from a "
                                                 f"{dist_field.nbytes >> 20} MiB L2-resident buffer, measured live"}}

        cpu_baseline = None
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

        total_rays = world * n_rays
        line = {
            "metric": METRIC, "value": total_rays * K / (dev_ms * 1e-3), "unit": "rays/s", "n_gpus": world,
            "steps": K, "warmup": args.warmup, "ms_per_step": dev_ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "poses_per_gpu": P, "beams": B,
                       "global_rays_per_step": total_rays, "map": f"{MAP_N}x{MAP_N} fp32 distance field "
                       f"({dist_field.nbytes >> 20} MiB, replicated per GPU)",
                       "parallelism": f"pose-sharded x{world}, map replicated" +
                                      {"none": ", no exchange step (rays are independent)" if dist_on else "",
                                       "allgather": ", NCCL all_gather of ranges inside the step",
                                       "p2p": ", ranges stored into every GPU's gathered buffer over NVLink by the march "
                                              "kernel (fused all-gather) + 4-byte all_reduce as barrier, inside the step"}[mode],
                       "l2": "no flush" if args.no_flush else "flushed between steps (256 MiB fill + cudaCtxResetPersistingL2Cache, so the "
                             "distance field the product pins in L2 is evicted too), excluded from the per-step events",
                       "timing": "CUDA events per step on the launching stream, summed, max over ranks",
                       "trig": "exact: glibc's sinf/cosf algorithm evaluated per beam on the device (bit parity with the host libm)"},
            "wall_ms_per_step_incl_flush": t_wall / K * 1e3,
            "e2e": {"value": total_rays * e2e_steps / e2e_s, "unit": "rays/s",
                    "h2d_bytes_per_step": P * 12, "d2h_bytes_per_step": n_rays * 4, "steps": e2e_steps,
                    "api": "ScanSimulator2D.scanMany(host poses) -> host ranges (pinned), per rank",
                    "pinned_d2h_gbs": d2h_gbs,
                    "numpy_pageable_buffers": {"value": n_rays * np_steps / np_s, "unit": "rays/s per GPU",
                                               "ms_per_call": np_s / np_steps * 1e3,
                                               "api": "PyRayMarchingGPU.calc_range_fan(np.ndarray poses, np.zeros outs): the "
                                                      "caller's array is page-locked on its second sighting (rl_host_register)"},
                    "pcie_bound_rays_per_s": world * d2h_gbs * 1e9 / 4.0},
            "e2e_fused_crash": {"value": n_rays * e2e_steps / fused_s, "unit": "nominal rays/s per GPU",
                                "ms_per_call": fused_s / e2e_steps * 1e3, "first_crash_index": crash_idx,
                                "api": "checkCollisionMany semantics (scripts/racecar_simulator_v2.py:146-167): host poses in, "
                                       "index of the first crashed pose out; rl_scan_crash skips every pose after it",
                                "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "ingest_ms": omap.ingest_ms,
        }
        if dist_on:
            def entry(ms, note):
                return {"value": total_rays * K / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms / K, "note": note}
            line["sharded"] = entry(ms_none, "ranges left on their GPUs: no exchange step")
            line["gather"] = entry(ms_p2p, "fused: the march kernel stores every range into all GPUs' gathered buffers over "
                                           "NVLink peer memory (rl_calc_range_fan_allgather; backend " + peer.backend +
                                           (", one multimem.st per range replicated by the NVSwitch" if peer.multicast else
                                            ", one store per peer") + (f" [symmetric memory unavailable: {getattr(peer, '_symm_error', '')[:200]}]"
                                                                      if peer.backend == "ipc" else "") +
                                           "), barrier after the kernel; "
                                           f"every GPU receives {(world - 1) * n_rays * 4 / 1e6:.0f} MB per step")
            line["gather_nccl"] = entry(ms_nccl, "march, then NCCL all_gather_into_tensor of the ranges")
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if peer is not None:
        peer.close()
    if dist_on:
        tdist.barrier()
        tdist.destroy_process_group()


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
