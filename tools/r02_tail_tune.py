"""Tail-mode tuning on BASELINE config 2 (4096 poses x 1080 beams, 2049^2 stand-in map): for each variant library
(csrc built with -DRL_TAIL_AFTER=a -DRL_TAIL_AHEAD=h into tools/variants/) one process measures
  cold    single launches into a flushed L2 with the field un-pinned (bench.py's `value` protocol)
  warm    single launches, L2 left alone
  steady  64 launches back to back on two internal streams (bench.py's `steady_state`)
and a SHA-1 of the ranges (every variant must give the same bits).
    python tools/r02_tail_tune.py <lib.so>[@ENV=VALUE,...] ...      (spawns one child per spec)
"""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(lib):
    import numpy as np
    import torch
    from pyracecarsimulator_b200 import _native
    _native.LIB_PATH = os.path.abspath(lib)
    from pyracecarsimulator_b200 import maps, range_libc
    FOV, P, B = 4.71, 4096, 1080
    img = maps.synth_map(2049, 1234)
    y = maps.synth_yaml(2049)
    path = f"/tmp/_rl_tail_{os.getpid()}.pgm"
    maps.write_pgm(path, img)
    y.image = path
    omap = range_libc.PyOMap(y)
    os.unlink(path)
    dist = omap.dist()
    rm = range_libc.PyRayMarchingGPU(omap, 300.0)
    sets = [torch.from_numpy(maps.sample_free_poses(dist, P, 1000 + s, y.resolution, y.origin)).cuda() for s in range(8)]
    out = torch.empty(P * B, dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    L = _native.lib()

    def one(i):
        rm.calc_range_fan(sets[i % 8], out, FOV, B)

    for i in range(8):
        one(i)
    torch.cuda.synchronize()
    sha = hashlib.sha1()
    for i in range(8):
        one(i)
        sha.update(out.cpu().numpy().tobytes())

    def timed(prepare, n=40):
        ts = []
        for i in range(n):
            prepare()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            one(i)
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2], ts[0]

    def cold():
        flush.fill_(1)
        torch.cuda.synchronize()
        L.rl_l2_reset_persisting(0)

    def warm():
        torch.cuda.synchronize()

    cold_med, cold_min = timed(cold)
    warm_med, warm_min = timed(warm)
    rm.set_pipelined("streams")
    K = 64
    best = 1e9
    for rep in range(5):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(K):
            one(i)
        rm.join()
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b) / K)
    rm.set_pipelined("off")
    print(json.dumps({"probe": "tail_tune", "lib": os.path.basename(os.environ.get("RL_TUNE_LABEL", lib)), "cold_ms_median": round(cold_med, 5),
                      "cold_ms_min": round(cold_min, 5), "warm_ms_median": round(warm_med, 5), "steady_ms_per_launch": round(best, 5),
                      "cold_grays_per_s": round(P * B / cold_med / 1e6, 2), "steady_grays_per_s": round(P * B / best / 1e6, 2),
                      "sha1": sha.hexdigest()[:12]}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--child":
        child(sys.argv[2])
    else:
        for spec in sys.argv[1:]:   # lib.so[@ENV=VALUE[,ENV=VALUE...]]
            lib, _, envs = spec.partition("@")
            env = dict(os.environ)
            for kv in filter(None, envs.split(",")):
                k, _, v = kv.partition("=")
                env[k] = v
            env["RL_TUNE_LABEL"] = spec
            subprocess.run([sys.executable, os.path.abspath(__file__), "--child", lib], check=False, env=env)
