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


def _worker(rank, world, port, n_poses, num_rays, gather, q):
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

        sc = ShardedScanner(march, num_rays, torch.device("cpu"))
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


@pytest.mark.parametrize("world,n_poses,gather", [(2, 37, "all"), (2, 37, "root"), (2, 37, "none"),
                                                  (3, 10, "all"), (2, 1, "all"), (2, 0, "all")])
def test_sharded_scan_matches_single_process(world, n_poses, gather):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_poses, 60, gather, q)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = dict(q.get(timeout=5) for _ in range(world))
    assert res == {r: True for r in range(world)}
