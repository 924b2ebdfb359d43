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
    print(f"subchunks={os.environ.get('RL_HOST_SUBCHUNKS', 'default')}: scanMany {t_sim:.0f} us, raw API {t_api:.0f} us, "
          f"bare pinned D2H {t_d2h:.0f} us, kernel {t_k:.0f} us -> e2e {P * B / t_sim / 1e3:.2f} Grays/s")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for n in ("1", "2", "3", "4", "6", "8", "12"):
            env = dict(os.environ, RL_HOST_SUBCHUNKS=n)
            subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env)
