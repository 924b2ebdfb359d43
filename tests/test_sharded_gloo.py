"""Multi-GPU host logic on CPU: world_size-2 (and 3) `gloo` process groups run the pose
partition, padding and the single range gather of pyracecarsimulator_b200.sharded with the
oracle standing in for the march kernel (test infrastructure: the product binds the CUDA
marcher, see sharded.gpu_march_fn)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from pyracecarsimulator_b200.sharded import shard_bounds


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 5, 4096, 1000003):
        for world in (1, 2, 3, 8):
            got = [shard_bounds(n, world, r) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            for (a, b), (c, d) in zip(got, got[1:]):
                assert b == c and a <= b and c <= d
            per = -(-n // world) if n else 0
            assert all(hi - lo <= per for lo, hi in got)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_poses, num_rays, gather, q, chunks=1):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from pyracecarsimulator_b200 import maps
        from pyracecarsimulator_b200.sharded import ShardedScanner
        img = maps.synth_map(129, 3)
        y = maps.synth_yaml(129)
        occ = oracle.omap_from_grid(oracle.mapserver_occupancy(img), True)
        dfield = oracle.edt_float(occ)                      # "replicated map": every rank builds its own
        m = oracle.Marcher(dfield, 300, y.resolution, y.origin)
        poses = maps.sample_free_poses(dfield, n_poses, 5, y.resolution, y.origin) if n_poses else np.zeros((0, 3), np.float32)

        def march(p, out):                                   # oracle stands in for the CUDA kernel
            if p.shape[0]:
                out.copy_(torch.from_numpy(m.calc_range_fan(p.numpy(), num_rays, 4.71)))

        sc = ShardedScanner(march, num_rays, torch.device("cpu"), chunks=chunks)
        got = sc.scan(torch.from_numpy(poses), gather=gather)
        want = m.calc_range_fan(poses, num_rays, 4.71) if n_poses else np.zeros(0, np.float32)
        if gather == "all":
            ok = np.array_equal(got.numpy(), want)
        elif gather == "root":
            ok = (got is None) if rank != 0 else np.array_equal(got.numpy(), want)
        else:
            lo, hi = shard_bounds(n_poses, world, rank)
            ok = np.array_equal(got.numpy(), want[lo * num_rays:hi * num_rays])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_poses,gather,chunks", [(2, 37, "all", 1), (2, 37, "root", 1), (2, 37, "none", 1),
                                                         (3, 10, "all", 1), (2, 1, "all", 1), (2, 0, "all", 1),
                                                         # chunk-pipelined all-gather: piece k-1 travels while piece k is marched
                                                         (2, 37, "all", 3), (3, 10, "all", 4), (2, 1, "all", 2), (2, 0, "all", 3),
                                                         (2, 64, "all", 64)])
def test_sharded_scan_matches_single_process(world, n_poses, gather, chunks):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_poses, 60, gather, q, chunks)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = dict(q.get(timeout=5) for _ in range(world))
    assert res == {r: True for r in range(world)}


def _rollout_worker(rank, world, port, n_cars, q):
    """ShardedRollout with a CPU stand-in for BatchedCar.rollout: the oracle's vehicle model driven
    by the oracle's copy of the device action schedule, crash = first step whose speed target
    exceeds a bound (any deterministic function of the GLOBAL car index would do)."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from pyracecarsimulator_b200.sharded import ShardedRollout
        steps, every = 30, 10
        p = oracle.car_params()

        def rollout(states, car_offset, steps_, seed, stream_id):
            n = states.shape[0]
            acts = oracle.rollout_actions(n, (steps_ + every - 1) // every, seed=seed, stream_id=stream_id,
                                          car_offset=car_offset)
            crash = np.full(n, -(steps_ + 1), np.int32)
            reward = np.zeros(n)
            for c in range(n):
                s = states[c].numpy().copy()
                vs = []
                for i in range(steps_):
                    oracle.car_step(p, s, acts[c, i // every, 0], acts[c, i // every, 1], 0.01)
                    vs.append(s[3])
                    if crash[c] < 0 and acts[c, i // every, 0] > 6.0:
                        crash[c] = i
                reward[c] = sum(vs[:crash[c]]) if crash[c] >= 0 else sum(vs)
            return torch.from_numpy(crash), torch.from_numpy(reward)

        rng = np.random.default_rng(9)
        s0 = np.zeros((n_cars, 11))
        s0[:, :3] = rng.uniform(-1, 1, (n_cars, 3))
        s0[:, 3] = 2.0
        states = torch.from_numpy(s0)
        sr = ShardedRollout(rollout, torch.device("cpu"))
        got_c, got_r = sr.rollout(states, steps, seed=42)
        want_c, want_r = rollout(states.clone(), 0, steps, 42, 0)       # the whole job in one process
        ok = torch.equal(got_c, want_c) and torch.equal(got_r, want_r) and torch.equal(states, torch.from_numpy(s0))
        loc_c, loc_r = sr.rollout(states, steps, seed=42, gather="none")
        lo, hi = shard_bounds(n_cars, world, rank)
        ok = ok and torch.equal(loc_c, want_c[lo:hi]) and torch.equal(loc_r, want_r[lo:hi])
        ok = ok and bool((want_c >= 0).any()) == (n_cars > 8 or bool((want_c >= 0).any()))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_cars", [(2, 37), (3, 10), (2, 1), (2, 0)])
def test_sharded_rollout_matches_single_process(world, n_cars):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rollout_worker, args=(r, world, port, n_cars, q)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = dict(q.get(timeout=5) for _ in range(world))
    assert res == {r: True for r in range(world)}


def test_gathered_buffer_sets_start_on_16_byte_boundaries():
    """The fused gather stores 16 bytes at a time: every buffer set of a PeerGather allocation must start on a
    16-byte boundary whatever the slot size (an odd slot_rays used to put the second set on a 4- or 8-byte one),
    and shapes that were already aligned keep their layout."""
    from pyracecarsimulator_b200.sharded import gather_set_layout
    for world in (1, 2, 3, 4, 8):
        for slot in (1, 3, 61, 183, 1080, 4096 * 1080, 125000 * 60, 262144 * 270, 7 * 61 + 1):
            for nbuf in (1, 2, 3):
                offs, total = gather_set_layout(world, slot, nbuf)
                assert len(offs) == nbuf and offs[0] == 0
                assert all((o * 4) % 256 == 0 for o in offs)
                assert all(b - a >= world * slot for a, b in zip(offs, offs[1:] + [total]))
                if (world * slot) % 64 == 0:
                    assert offs == [b * world * slot for b in range(nbuf)] and total == nbuf * world * slot
