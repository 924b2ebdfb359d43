"""Where the end-to-end step goes: host pipeline depth sweep (RL_HOST_SUBCHUNKS), Python wrapper vs raw
C call, and the bare pinned D2H copy.  Development helper."""
import os
import subprocess
import sys
import time


sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child():
    import torch
    from pyracecarsimulator_b200 import maps, range_libc
    from pyracecarsimulator_b200.scan_simulator import ScanSimulator2D
    img = maps.synth_map(2049, 1234)
    y = maps.synth_yaml(2049)
    path = "/tmp/_probe.pgm"
    maps.write_pgm(path, img)
    y.image = path
    omap = range_libc.PyOMap(y)
    dist = omap.dist()
    P, B = 4096, 1080
    poses = maps.sample_free_poses(dist, P, 1, y.resolution, y.origin)
    sim = ScanSimulator2D(B, 4.71, 0.01, batch_size=P)
    sim.setMap(omap, 300, y.resolution, y.origin)
    sim.setRaytracingMethod("RMGPU")
    for _ in range(5):
        sim.scanMany(poses)

    def loop(fn, n=40):
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t) / n * 1e6

    t_sim = loop(lambda: sim.scanMany(poses))
    rm = sim.scan_method
    t_api = loop(lambda: rm.calc_range_fan(sim.poses_many, sim.output_vector_many, 4.71, B))
    d_out = torch.empty(P * B, dtype=torch.float32, device="cuda")
    h = torch.empty(P * B, dtype=torch.float32, pin_memory=True)
    t_d2h = loop(lambda: h.copy_(d_out, non_blocking=True))
    dp = torch.from_numpy(poses).cuda()
    t_k = loop(lambda: rm.calc_range_fan(dp, d_out, 4.71, B))
    if os.environ.get("RL_PROBE_PAGEABLE"):
        # plain numpy buffers (what a caller of the raw range_libc API passes): first call = staged through the
        # marcher's pinned memory + memcpy; from the second sighting on the buffer is page-locked by the shim
        import numpy as np
        outs = np.zeros(P * B, np.float32)
        t0 = time.perf_counter(); rm.calc_range_fan(poses, outs, 4.71, B); t_first = (time.perf_counter() - t0) * 1e6
        t0 = time.perf_counter(); rm.calc_range_fan(poses, outs, 4.71, B); t_second = (time.perf_counter() - t0) * 1e6
        t_rest = loop(lambda: rm.calc_range_fan(poses, outs, 4.71, B), 20)
        range_libc.release_host_buffers()
        os.environ["RL_HOST_REGISTER"] = "0"
        range_libc._HOST_REGISTRY.enabled = False
        t_staged = loop(lambda: rm.calc_range_fan(poses, outs, 4.71, B), 20)
        print(f"pageable numpy outs: first call {t_first:.0f} us (staged), second {t_second:.0f} us (includes cudaHostRegister), "
              f"then {t_rest:.0f} us per call = {P * B / t_rest / 1e3:.2f} Grays/s; registration disabled: {t_staged:.0f} us per call "
              f"= {P * B / t_staged / 1e3:.2f} Grays/s")
        return
    print(f"zerocopy={os.environ.get('RL_HOST_ZEROCOPY', '1')} subchunks={os.environ.get('RL_HOST_SUBCHUNKS', 'default')}: scanMany {t_sim:.0f} us, raw API {t_api:.0f} us, "
          f"bare pinned D2H {t_d2h:.0f} us, kernel {t_k:.0f} us -> e2e {P * B / t_sim / 1e3:.2f} Grays/s")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    elif len(sys.argv) > 1 and sys.argv[1] == "pageable":
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=dict(os.environ, RL_PROBE_PAGEABLE="1"))
    else:
        env = dict(os.environ, RL_HOST_ZEROCOPY="1")
        env.pop("RL_HOST_SUBCHUNKS", None)
        print("zero-copy stores (kernel writes the pinned host buffer):", flush=True)
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env)
        print("copy-engine pipeline (RL_HOST_ZEROCOPY=0), by sub-chunk count:", flush=True)
        for n in ("1", "2", "4", "6", "8", "12", "16", "24"):
            env = dict(os.environ, RL_HOST_SUBCHUNKS=n, RL_HOST_ZEROCOPY="0")
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env)
