"""Multi-GPU correctness check, spawned by tests/test_gpu_multi.py for world sizes 2 / 4 / 8 when the box has
that many GPUs (or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \\
        --master-port 29555 tests/multigpu_check.py

Every rank builds its replica of the map, then ShardedScanner's three NCCL gather modes and the fused
peer-memory gather are compared bit for bit with the single-GPU scan of the whole batch, and
ShardedRollout (cars sharded, device-drawn actions per global car, gather of crash index + reward)
with the single-GPU rollout of all cars."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyracecarsimulator_b200 import maps, range_libc  # noqa: E402
from pyracecarsimulator_b200.racecar import BatchedCar  # noqa: E402
from pyracecarsimulator_b200.sharded import (ShardedRollout, ShardedScanner, gpu_march_angles_fn, gpu_march_fn,  # noqa: E402
                                             gpu_rollout_fn, shard_bounds)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    img = maps.synth_map(513, 21)
    y = maps.synth_yaml(513)
    path = f"/tmp/_mg_{rank}.pgm"
    maps.write_pgm(path, img)
    y.image = path
    omap = range_libc.PyOMap(y, device=local)
    rm = range_libc.PyRayMarchingGPU(omap, 300)
    fov, R = 4.71, 270
    ok = True
    sc = ShardedScanner(gpu_march_fn(rm, fov, R), R, dev)
    for n in (1001, 64, 7, 1):                       # uneven shards, fewer poses than ranks
        poses = torch.from_numpy(maps.sample_free_poses(omap.dist(), n, 5 + n, y.resolution, y.origin))
        want = torch.empty(n * R, dtype=torch.float32, device=dev)
        rm.calc_range_fan(poses.to(dev), want, fov, R)           # whole batch on this GPU
        got_all = sc.scan(poses, gather="all")
        got_root = sc.scan(poses, gather="root")
        got_none = sc.scan(poses, gather="none")
        lo, hi = shard_bounds(n, world, rank)
        got_fused = sc.scan_fused(poses, rm, fov).clone()
        checks = [torch.equal(got_all, want), (got_root is None) if rank else torch.equal(got_root, want),
                  torch.equal(got_none, want[lo * R:hi * R]), torch.equal(got_fused, want)]
        ok = ok and all(checks)
        if rank == 0:
            print(f"n={n}: all={checks[0]} root={checks[1]} none={checks[2]} fused={checks[3]}")
    # chunk-pipelined NCCL all-gather (piece k-1 travels while piece k is marched)
    sc3 = ShardedScanner(gpu_march_fn(rm, fov, R), R, dev, chunks=3)
    for n in (1001, 7):
        poses = torch.from_numpy(maps.sample_free_poses(omap.dist(), n, 5 + n, y.resolution, y.origin))
        want = torch.empty(n * R, dtype=torch.float32, device=dev)
        rm.calc_range_fan(poses.to(dev), want, fov, R)
        c = torch.equal(sc3.scan(poses, gather="all"), want)
        ok = ok and c
        if rank == 0:
            print(f"n={n}: chunked all-gather={c}")
    # the particle-filter shape (config 3): calc_range_repeat_angles sharded, NCCL and fused gathers
    A = 60
    angles = torch.from_numpy(np.linspace(-fov / 2, fov / 2, A, endpoint=False).astype(np.float32)).to(dev)
    sca = ShardedScanner(gpu_march_angles_fn(rm, angles), A, dev)
    for n in (4099, 64, 3):
        poses = torch.from_numpy(maps.sample_free_poses(omap.dist(), n, 9 + n, y.resolution, y.origin))
        want = torch.empty(n * A, dtype=torch.float32, device=dev)
        rm.calc_range_repeat_angles(poses.to(dev), angles, want)
        checks = [torch.equal(sca.scan(poses, gather="all"), want), torch.equal(sca.scan_fused(poses, rm, angles=angles), want)]
        ok = ok and all(checks)
        if rank == 0:
            print(f"repeat_angles n={n}: nccl={checks[0]} fused={checks[1]}")
    # the fused gathers through the map-order kernel (march_territory_kernel with peer stores), forced on these sizes
    os.environ.update({"RL_SORT_POSES": "1", "RL_SORT_MIN_POSES": "1"})
    rmt = range_libc.PyRayMarchingGPU(omap, 300)
    for k in ("RL_SORT_POSES", "RL_SORT_MIN_POSES"):
        os.environ.pop(k, None)
    sct = ShardedScanner(gpu_march_fn(rmt, fov, R), R, dev)
    scta = ShardedScanner(gpu_march_angles_fn(rmt, angles), A, dev)
    for n in (4099, 64, 3):
        poses = torch.from_numpy(maps.sample_free_poses(omap.dist(), n, 9 + n, y.resolution, y.origin))
        want = torch.empty(n * R, dtype=torch.float32, device=dev)
        rm.calc_range_fan(poses.to(dev), want, fov, R)
        wanta = torch.empty(n * A, dtype=torch.float32, device=dev)
        rm.calc_range_repeat_angles(poses.to(dev), angles, wanta)
        checks = [torch.equal(sct.scan_fused(poses, rmt, fov), want), torch.equal(scta.scan_fused(poses, rmt, angles=angles), wanta),
                  torch.equal(sct.scan(poses, gather="all"), want)]
        ok = ok and all(checks)
        if rank == 0:
            print(f"territories n={n}: fused fan={checks[0]} fused angles={checks[1]} nccl={checks[2]}")
    # back-to-back fused calls: the view of call i stays valid while call i+1 runs (two buffer sets)
    pa = torch.from_numpy(maps.sample_free_poses(omap.dist(), 800, 1, y.resolution, y.origin))
    pb = torch.from_numpy(maps.sample_free_poses(omap.dist(), 800, 2, y.resolution, y.origin))
    wa = torch.empty(800 * R, dtype=torch.float32, device=dev)
    wb = torch.empty(800 * R, dtype=torch.float32, device=dev)
    rm.calc_range_fan(pa.to(dev), wa, fov, R)
    rm.calc_range_fan(pb.to(dev), wb, fov, R)
    scf = ShardedScanner(gpu_march_fn(rm, fov, R), R, dev)
    c = True
    for _ in range(4):
        va = scf.scan_fused(pa, rm, fov)
        vb = scf.scan_fused(pb, rm, fov)
        c = c and torch.equal(va, wa) and torch.equal(vb, wb)
    ok = ok and c
    if rank == 0:
        print(f"back-to-back fused views: {c}")
    # the all-gather alone (ranges that already exist): slot r of every GPU == rank r's shard
    lo, hi = shard_bounds(800, world, rank)
    per = -(-800 // world)
    peer = scf._peer
    mine = wa[lo * R:hi * R].contiguous()
    peer.gather(mine, int(torch.cuda.current_stream(local).cuda_stream))
    peer.sync()
    g = peer.tensor()
    c = all(torch.equal(g[r * peer.slot_rays: r * peer.slot_rays + (min(800, (r + 1) * per) - r * per) * R],
                        wa[r * per * R: min(800, (r + 1) * per) * R]) for r in range(world))
    ok = ok and c
    if rank == 0:
        print(f"all-gather of existing ranges: {c}")
    # a batch gathered in pieces through the alternating buffer sets (config 5's path)
    n, piece = 1003, 100
    poses = torch.from_numpy(maps.sample_free_poses(omap.dist(), n, 99, y.resolution, y.origin))
    want = torch.empty(n * R, dtype=torch.float32, device=dev)
    rm.calc_range_fan(poses.to(dev), want, fov, R)
    per = -(-n // world)
    got = torch.full((n * R,), -1.0, dtype=torch.float32, device=dev)
    for first, counts, gathered in scf.scan_fused_chunks(poses, rm, fov, piece):
        for r in range(world):
            if counts[r] > 0:
                a = (r * per + first) * R
                got[a:a + counts[r] * R] = gathered[r, :counts[r] * R]
    c = torch.equal(got, want)
    ok = ok and c
    if rank == 0:
        print(f"gathered in pieces: {c}")
    for s_ in (sca, scf):
        if getattr(s_, "_peer", None) is not None:
            s_._peer.close()
    # fused rollout, cars sharded (config 4 shape, reduced): 1001 / 5 / 1 cars x 30 steps x 270 beams
    car = BatchedCar(device=local)
    car.setCarEdgeDistances(R, -fov / 2.0, fov / R, 0.275)
    sr = ShardedRollout(gpu_rollout_fn(car, rm, fov), dev)
    for n in (1001, 5, 1):
        s0 = np.zeros((n, 11))
        s0[:, :3] = maps.sample_free_poses(omap.dist(), n, 77 + n, y.resolution, y.origin, min_clear_px=6.0)
        s0[:, 3] = 2.0
        states = torch.from_numpy(s0)
        whole = car.rollout(rm, states.to(dev).clone(), None, 30, fov, seed=42)      # all cars on this GPU
        got_c, got_r = sr.rollout(states, 30, seed=42)
        checks = [torch.equal(got_c, whole["crash_index"]), torch.equal(got_r, whole["reward"])]
        ok = ok and all(checks)
        if rank == 0:
            print(f"rollout n={n}: crash_index={checks[0]} reward={checks[1]} "
                  f"({int((whole['crash_index'] >= 0).sum())} cars crash)")
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIGPU CHECK", "PASSED" if int(flag.item()) else "FAILED", f"(world {world})")
    if getattr(sc, "_peer", None) is not None:
        sc._peer.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
