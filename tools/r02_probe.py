"""Round-2 single-GPU measurements that decide product defaults (not part of bench.py):
  edt      scan vs divide-and-conquer EDT row pass on dense and sparse maps
  steady   K back-to-back config-2 launches under one event pair: stream order vs two-stream vs PDL
  calib    4-byte gather and full-sector L2 calibrations
  cfg      configs 3 and 5 (one GPU's share), device time
  terr     map order + SM territories on configs 3 / 5, per variant
    python tools/r02_probe.py [edt] [steady] [calib] [cfg] [terr]
Prints one JSON line per measurement."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyracecarsimulator_b200 import _native, maps, range_libc  # noqa: E402

FOV = 4.71


def ingest_ms(arg, which, reps=3):
    os.environ["RL_EDT_ROWS"] = which
    best = 1e9
    for _ in range(reps):
        om = range_libc.PyOMap(arg)
        best = min(best, om.ingest_ms)
        del om
    os.environ.pop("RL_EDT_ROWS", None)
    return best


def edt():
    cases = {}
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "colombia_map.npz"))
    import oracle
    g = oracle.mapserver_occupancy(z["img"])
    cases["colombia 435x350"] = maps.OccupancyGrid.make(g.ravel(), 435, 350, 0.05, (0.0, 0.0, 0.0))
    for n, seed in ((2049, 1234), (4096, 5678), (8192, 5678)):
        g = oracle.mapserver_occupancy(maps.synth_map(n, seed))
        cases[f"synth {n}^2"] = maps.OccupancyGrid.make(g.ravel(), n, n, 0.05, (0.0, 0.0, 0.0))
    for n in (2048, 8192):
        lone = np.zeros((n, n), bool)
        lone[n // 3, (2 * n) // 3] = True
        cases[f"lone obstacle {n}^2"] = lone
        border = np.zeros((n, n), bool)
        border[0, :] = border[-1, :] = True
        border[:, 0] = border[:, -1] = True
        cases[f"border only {n}^2"] = border
        rng = np.random.default_rng(n)
        sparse = rng.random((n, n)) < 2e-6
        sparse[0, 0] = True
        cases[f"{int(sparse.sum())} random cells {n}^2"] = sparse
    for name, arg in cases.items():
        row = {"probe": "edt", "map": name}
        for which in ("auto", "dc", "scan"):
            row[which + "_ms"] = ingest_ms(arg, which)
        for b in ("8", "16", "64", "128"):
            os.environ["RL_EDT_BUDGET"] = b
            row["auto_budget" + b + "_ms"] = ingest_ms(arg, "auto")
        os.environ.pop("RL_EDT_BUDGET", None)
        print(json.dumps(row), flush=True)


def config2():
    img = maps.synth_map(2049, 1234)
    y = maps.synth_yaml(2049)
    path = f"/tmp/_rl_probe_{os.getpid()}.pgm"
    maps.write_pgm(path, img)
    y.image = path
    omap = range_libc.PyOMap(y)
    os.unlink(path)
    return omap, y, omap.dist()


def steady():
    omap, y, dist = config2()
    P, B, K = 4096, 1080, 64
    sets = [torch.from_numpy(maps.sample_free_poses(dist, P, 1000 + s, y.resolution, y.origin)).cuda() for s in range(4)]
    outs = [torch.empty(P * B, dtype=torch.float32, device="cuda") for _ in range(4)]
    rm = range_libc.PyRayMarchingGPU(omap, 300)
    want = []
    for s in range(4):
        rm.calc_range_fan(sets[s], outs[s], FOV, B)
        want.append(outs[s].clone())
    torch.cuda.synchronize()
    for mode in ("off", "streams", "pdl"):
        rm.set_pipelined(mode)
        res = []
        for rep in range(5):
            for o in outs:
                o.zero_()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(K):
                rm.calc_range_fan(sets[i % 4], outs[i % 4], FOV, B)
            rm.join()
            b.record()
            torch.cuda.synchronize()
            res.append(a.elapsed_time(b) / K)
        same = all(torch.equal(o, w) for o, w in zip(outs, want))
        rm.set_pipelined("off")
        ms = float(np.median(res))
        print(json.dumps({"probe": "steady", "mode": mode, "ms_per_launch": ms, "grays_per_s": P * B / (ms * 1e-3) / 1e9,
                          "launches": K, "bit_identical": bool(same), "all_ms": res}), flush=True)
    # host launch rate: how long does one call take on the host?
    for mode in ("off", "streams"):
        rm.set_pipelined(mode)
        small = sets[0][:8].contiguous()
        o = torch.empty(8 * B, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(2000):
            rm.calc_range_fan(small, o, FOV, B)
        rm.join()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        rm.set_pipelined("off")
        print(json.dumps({"probe": "host_call_us", "mode": mode, "us": (t1 - t0) / 2000 * 1e6}), flush=True)


def calib():
    for mb in (16.8, 64, 256):
        nbytes = int(mb * 1e6)
        g = _native.gather_bandwidth(0, nbytes, 64, 10)
        s = _native.l2_sector_bandwidth(0, nbytes, 64, 10)
        print(json.dumps({"probe": "calib", "buffer_mb": mb, "gather4_gbs": g, "gather4_gsectors_per_s": g / 4.0,
                          "sector_gsectors_per_s": s, "sector_gbs": s * 32}), flush=True)


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def cfg():
    omap, y, dist = config2()
    rm = range_libc.PyRayMarchingGPU(omap, 300)
    n, a = 1_000_000, 60
    poses = torch.from_numpy(maps.sample_free_poses(dist, n, 303, y.resolution, y.origin)).cuda()
    angles = torch.from_numpy(np.linspace(-FOV / 2, FOV / 2, a, endpoint=False).astype(np.float32)).cuda()
    out = torch.empty(n * a, dtype=torch.float32, device="cuda")
    ms = timeit(lambda: rm.calc_range_repeat_angles(poses, angles, out))
    rm.count_steps(True)
    rm.calc_range_repeat_angles(poses, angles, out)
    steps = rm.last_steps()
    rm.count_steps(False)
    print(json.dumps({"probe": "cfg3", "ms": ms, "grays_per_s": n * a / ms / 1e6, "steps_per_ray": steps / (n * a)}), flush=True)
    del poses, out, rm, omap
    img = maps.synth_map(8192, 5678)
    y5 = maps.synth_yaml(8192)
    path = f"/tmp/_rl_probe5_{os.getpid()}.pgm"
    maps.write_pgm(path, img)
    y5.image = path
    omap5 = range_libc.PyOMap(y5)
    os.unlink(path)
    dist5 = omap5.dist()
    for sort in ("1", "0"):
      os.environ["RL_SORT_POSES"] = sort
      rm5 = range_libc.PyRayMarchingGPU(omap5, 300)
      os.environ.pop("RL_SORT_POSES", None)
      for n5 in (2_000_000,):
        b5 = 270
        poses5 = torch.from_numpy(maps.sample_free_poses(dist5, n5, 505, y5.resolution, y5.origin)).cuda()
        out5 = torch.empty(n5 * b5, dtype=torch.float32, device="cuda")
        ms = timeit(lambda: rm5.calc_range_fan(poses5, out5, FOV, b5))
        rm5.count_steps(True)
        rm5.calc_range_fan(poses5, out5, FOV, b5)
        steps = rm5.last_steps()
        rm5.count_steps(False)
        chk = float(out5.double().sum().item())
        print(json.dumps({"probe": "cfg5_share", "map_order": sort == "1", "poses": n5, "ms": ms, "grays_per_s": n5 * b5 / ms / 1e6,
                          "steps_per_ray": steps / (n5 * b5), "ingest_ms": omap5.ingest_ms, "checksum": chk}), flush=True)
        chunk = 1 << 18
        ms = timeit(lambda: [rm5.calc_range_fan(poses5[a:a + chunk], out5[:chunk * b5], FOV, b5) for a in range(0, n5, chunk)])
        print(json.dumps({"probe": "cfg5_share_in_pieces_of_262144", "map_order": sort == "1", "ms": ms,
                          "grays_per_s": n5 * b5 / ms / 1e6}), flush=True)


VARIANTS = [
    ("caller order", {"RL_SORT_POSES": "0"}),
    ("territories, 16 px cells (default)", {"RL_SORT_POSES": "1"}),
    ("territories, 4 px cells", {"RL_SORT_POSES": "1", "RL_SORT_SHIFT": "2"}),
    ("territories, 64 px cells", {"RL_SORT_POSES": "1", "RL_SORT_SHIFT": "6"}),
]


def with_env(env, make):
    keys = ("RL_SORT_POSES", "RL_SORT_SHIFT", "RL_SORT_MIN_POSES")
    for k in keys:
        os.environ.pop(k, None)
    os.environ.update(env)
    try:
        return make()
    finally:
        for k in keys:
            os.environ.pop(k, None)


def terr():
    """Map order + SM territories: configs 3 and 5 (one GPU's share) and a clustered particle set, per variant."""
    omap, y, dist = config2()
    n, a = 1_000_000, 60
    uniform = maps.sample_free_poses(dist, n, 303, y.resolution, y.origin)
    # a particle filter's cloud: all poses within ~1 m of one place, headings within +-0.3 rad
    rng = np.random.default_rng(9)
    c = uniform[12345]
    cloud = np.empty((n, 3), np.float32)
    cloud[:, 0] = c[0] + rng.normal(0, 0.5, n)
    cloud[:, 1] = c[1] + rng.normal(0, 0.5, n)
    cloud[:, 2] = c[2] + rng.normal(0, 0.3, n)
    angles = torch.from_numpy(np.linspace(-FOV / 2, FOV / 2, a, endpoint=False).astype(np.float32)).cuda()
    out = torch.empty(n * a, dtype=torch.float32, device="cuda")
    ref = {}
    for name, env in VARIANTS:
        rm = with_env(env, lambda: range_libc.PyRayMarchingGPU(omap, 300))
        for label, ps in (("cfg3 uniform", uniform), ("cfg3-shaped particle cloud", cloud)):
            poses = torch.from_numpy(ps).cuda()
            ms = timeit(lambda: rm.calc_range_repeat_angles(poses, angles, out))
            chk = float(out.double().sum().item())
            ref.setdefault(label, out.clone())
            print(json.dumps({"probe": "terr", "case": label, "variant": name, "ms": ms, "grays_per_s": n * a / ms / 1e6,
                              "identical": bool(torch.equal(out, ref[label])), "checksum": chk}), flush=True)
        del rm
    # config 2 must not care (4096 poses < the sort threshold) -- and what if it did sort?
    p2 = torch.from_numpy(maps.sample_free_poses(dist, 4096, 1000, y.resolution, y.origin)).cuda()
    out2 = torch.empty(4096 * 1080, dtype=torch.float32, device="cuda")
    for name, env in (("caller order", {"RL_SORT_POSES": "0"}), ("default", {}),
                      ("forced territories", {"RL_SORT_POSES": "1", "RL_SORT_MIN_POSES": "1"})):
        rm = with_env(env, lambda: range_libc.PyRayMarchingGPU(omap, 300))
        ms = timeit(lambda: rm.calc_range_fan(p2, out2, FOV, 1080), reps=20)
        print(json.dumps({"probe": "terr", "case": "cfg2 warm", "variant": name, "ms": ms, "grays_per_s": 4096 * 1080 / ms / 1e6}), flush=True)
        del rm
    # where should the threshold be?  N poses x 1080 beams and N poses x 60 angles, caller order vs territories
    for nb, kind in ((1080, "fan"), (60, "angles")):
        for npz in (2048, 4096, 8192, 16384, 65536):
            if npz * nb > 80_000_000:
                continue
            pz = torch.from_numpy(maps.sample_free_poses(dist, npz, 77 + npz, y.resolution, y.origin)).cuda()
            oz = torch.empty(npz * nb, dtype=torch.float32, device="cuda")
            row = {"probe": "terr_threshold", "poses": npz, "beams": nb}
            for name, env in (("caller_order", {"RL_SORT_POSES": "0"}), ("territories", {"RL_SORT_POSES": "1", "RL_SORT_MIN_POSES": "1"})):
                rm = with_env(env, lambda: range_libc.PyRayMarchingGPU(omap, 300))
                if kind == "fan":
                    row[name + "_ms"] = timeit(lambda: rm.calc_range_fan(pz, oz, FOV, nb), reps=9)
                else:
                    row[name + "_ms"] = timeit(lambda: rm.calc_range_repeat_angles(pz, angles, oz), reps=9)
                del rm
            print(json.dumps(row), flush=True)
            del pz, oz
    del omap, out, out2
    img = maps.synth_map(8192, 5678)
    y5 = maps.synth_yaml(8192)
    path = f"/tmp/_rl_probe5_{os.getpid()}.pgm"
    maps.write_pgm(path, img)
    y5.image = path
    omap5 = range_libc.PyOMap(y5)
    os.unlink(path)
    dist5 = omap5.dist()
    n5, b5 = 2_000_000, 270
    poses5 = torch.from_numpy(maps.sample_free_poses(dist5, n5, 505, y5.resolution, y5.origin)).cuda()
    out5 = torch.empty(n5 * b5, dtype=torch.float32, device="cuda")
    ref5 = None
    for name, env in VARIANTS:
        rm5 = with_env(env, lambda: range_libc.PyRayMarchingGPU(omap5, 300))
        ms = timeit(lambda: rm5.calc_range_fan(poses5, out5, FOV, b5), reps=3)
        if ref5 is None:
            ref5 = out5.clone()
        print(json.dumps({"probe": "terr", "case": "cfg5 share (2M poses x 270)", "variant": name, "ms": ms,
                          "grays_per_s": n5 * b5 / ms / 1e6, "identical": bool(torch.equal(out5, ref5))}), flush=True)
        del rm5


if __name__ == "__main__":
    what = sys.argv[1:] or ["edt", "steady", "calib", "cfg"]
    for w in what:
        globals()[w]()
