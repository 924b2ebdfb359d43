"""Where the multi-GPU end-to-end number goes: per-GPU device->host bandwidth for the
17.7 MB of ranges one scanMany step returns, each rank ALONE vs ALL ranks at once, for
  ce        copy engine into cudaHostAlloc'ed pinned memory (torch pin_memory)
  zc        the product's zero-copy path: the march kernel stores straight into the pinned buffer
  zc_wc     the same into write-combined pinned memory (cudaHostAllocWriteCombined)
  scanMany  ScanSimulator2D.scanMany end to end (H2D poses + kernel + ranges on the host)
with and without binding the rank to the CPUs next to its GPU before the pinned allocation (RL_BIND=0/1).

    RL_BIND=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29561 tools/d2h_probe.py
Rank 0 prints one JSON line per measurement (min / mean over ranks of GB/s per GPU, and the sum)."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (bind_cpus, gpu_map)
from pyracecarsimulator_b200 import maps, range_libc  # noqa: E402
from pyracecarsimulator_b200.scan_simulator import ScanSimulator2D  # noqa: E402

P, B, FOV = 4096, 1080, 4.71


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    bind = os.environ.get("RL_BIND", "1") != "0"
    aff = bench.bind_cpus(local, world) if bind else {"bound": False}
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    omap, y = bench.gpu_map(2049, 1234, local)
    rm = range_libc.PyRayMarchingGPU(omap, 300)
    poses = maps.sample_free_poses(omap.dist(), P, 1000 + rank, y.resolution, y.origin)
    n = P * B
    d_out = torch.empty(n, dtype=torch.float32, device=dev)
    h_pin = torch.empty(n, dtype=torch.float32, pin_memory=True)
    h_np = h_pin.numpy()
    poses_pin = torch.from_numpy(poses).pin_memory().numpy()
    # write-combined pinned buffer through the CUDA runtime
    wc_np = None
    try:
        from cuda.bindings import runtime as rt
        err, ptr = rt.cudaHostAlloc(n * 4, rt.cudaHostAllocWriteCombined | rt.cudaHostAllocMapped | rt.cudaHostAllocPortable)
        if int(err) == 0:
            wc_np = np.ctypeslib.as_array((ctypes.c_float * n).from_address(int(ptr)))
    except Exception as e:   # noqa: BLE001
        if rank == 0:
            print(json.dumps({"probe": "d2h", "wc_error": repr(e)[:200]}), flush=True)
    sim = ScanSimulator2D(B, FOV, 0.01, batch_size=P)
    sim.setMap(omap, 300, y.resolution, y.origin)
    sim.setRaytracingMethod("RMGPU")

    def ce():
        h_pin.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()

    def zc():
        rm.calc_range_fan(poses_pin, h_np, FOV, B)

    def zc_wc():
        rm.calc_range_fan(poses_pin, wc_np, FOV, B)

    def scan_many():
        sim.scanMany(poses)

    def rate(fn, reps=8):
        fn()
        fn()
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return reps * n * 4 / (time.perf_counter() - t) / 1e9

    tests = [("ce", ce), ("zc", zc), ("scanMany", scan_many)]
    if wc_np is not None:
        tests.insert(2, ("zc_wc", zc_wc))
    for name, fn in tests:
        alone = 0.0
        for r in range(world):
            dist.barrier()
            torch.cuda.synchronize()
            if r == rank:
                alone = rate(fn)
        dist.barrier()
        torch.cuda.synchronize()
        together = rate(fn, 16)
        t = torch.tensor([alone, together], dtype=torch.float64, device=dev)
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        if rank == 0:
            a = np.array([v.cpu().numpy() for v in allv])
            print(json.dumps({"probe": "d2h", "path": name, "world": world, "bound_to_gpu_cpus": bind, "affinity": aff,
                              "alone_gbs_per_gpu": [round(float(x), 2) for x in a[:, 0]],
                              "together_gbs_per_gpu": [round(float(x), 2) for x in a[:, 1]],
                              "alone_min": float(a[:, 0].min()), "together_min": float(a[:, 1].min()),
                              "together_sum": float(a[:, 1].sum()),
                              "together_rays_per_s": float(a[:, 1].sum()) * 1e9 / 4}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
