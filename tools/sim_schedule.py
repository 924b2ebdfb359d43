"""CPU model of how the march kernel's CTAs fill the GPU, to judge work-ordering ideas before spending GPU
time on them.  Development helper, not on the product path (uses the oracle for per-ray step counts).

    python tools/sim_schedule.py            # BASELINE config 2: 4096 poses x 1080 beams, synth_map(2049, 1234)

Model (calibrated on profiles/r01_timeline.md and the DIAG3 table in profiles/r01_tuning.md section 11):
  * a warp costs W = 181 + 17 * S instructions and cannot finish faster than T = 0.6 + 0.127 * S us
    (S = the longest ray of the warp in march steps; 0.127 us = the ~250 cycles one dependent step takes);
  * an SM issues 5300 warp-instructions per us (2.7 IPC, what the bulk phase sustains) shared by its
    resident warps, 16 CTAs of 4 warps per SM, 148 SMs, CTAs dispatched in grid order to the first free slot.
It reproduces the measured 83 us kernel and its 23 us drain.  What it says about ordering:
  * the drain is the last-dispatched long warps, so "longest first" would end at ~62-65 us (-22 %);
  * that needs the step counts in advance; a probe pass that marches every 32nd beam (3 % of the rays,
    capped at 48 steps) and sorts poses by its longest probe gets 71 us in the model, but the probe and the
    sort are two dependent launches in front of the march (~8 us), and a cheap two- or four-class partition of
    the last 40 % of the poses (probe overlapped with the first 60 %) only reaches 77-80 us;
  * the pose's own clearance predicts nothing (86-89 us).
Conclusion recorded in DESIGN.md section 7: not worth its complexity for a 4-7 % gain on this batch shape.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

NSM, SLOTS, CAP = 148, 16, 5300.0


def simulate(wmax, dt=0.05, fixed_instr=181.0, per_step_instr=17.0, tfix=0.6, tstep=0.127):
    """wmax: longest ray (steps) of every warp, in dispatch order, 4 warps per CTA.  Returns the span in us."""
    pad = (-wmax.size) % 4
    wmax = np.concatenate([wmax.astype(np.float64), np.ones(pad)])
    work = (fixed_instr + per_step_instr * wmax).reshape(-1, 4)
    tmin = (tfix + tstep * wmax).reshape(-1, 4)
    ncta = work.shape[0]
    prog = np.ones((NSM * SLOTS, 4))
    sw = np.ones_like(prog)
    stm = np.ones_like(prog)
    nxt, t = 0, 0.0
    while True:
        free = np.flatnonzero((prog >= 1).all(1))
        if nxt < ncta and free.size:
            k = min(free.size, ncta - nxt)
            sl = free[:k]
            prog[sl] = 0
            sw[sl] = work[nxt:nxt + k]
            stm[sl] = tmin[nxt:nxt + k]
            nxt += k
        active = prog < 1
        if not active.any() and nxt >= ncta:
            return t
        demand = np.where(active, sw / stm, 0).reshape(NSM, -1).sum(1)
        scale = np.minimum(1.0, CAP / np.maximum(demand, 1e-9))
        prog = np.where(active, prog + (1.0 / stm) * np.repeat(scale, SLOTS)[:, None] * dt, prog)
        t += dt


def main():
    import oracle
    from pyracecarsimulator_b200 import maps
    img = maps.synth_map(2049, 1234)
    y = maps.synth_yaml(2049)
    dist = oracle.edt_float(oracle.omap_from_grid(oracle.mapserver_occupancy(img), True))
    poses = maps.sample_free_poses(dist, 4096, 1000, y.resolution, y.origin)
    _, steps = oracle.Marcher(dist, 300, y.resolution, y.origin).calc_range_fan(poses, 1080, 4.71, steps=True, threads=0)
    S = steps.reshape(4096, 1080)

    def span(pose_order, extra_front=None):
        w = S[pose_order].reshape(-1, 32).max(1)
        if extra_front is not None:
            w = np.concatenate([extra_front, w])
        return simulate(w)

    nat = np.arange(4096)
    print(f"natural order                          {span(nat):6.1f} us   (measured: 83 us)")
    print(f"poses sorted by their longest ray      {span(np.argsort(-S.max(1), kind='stable')):6.1f} us   (needs the answer)")
    for stride in (4, 8, 16, 32):
        score = np.minimum(S[:, ::stride], 48).max(1)
        print(f"sorted by a stride-{stride:<2d} probe (cap 48)     {span(np.argsort(-score, kind='stable')):6.1f} us"
              f"   + probe and sort launches, {100 / stride:.1f} % extra rays")
    col = ((poses[:, 0] - y.origin[0]) / y.resolution).astype(int)
    row = ((poses[:, 1] - y.origin[1]) / y.resolution).astype(int)
    d0 = dist[row, col]
    print(f"sorted by the pose's clearance         {span(np.argsort(d0)):6.1f} / {span(np.argsort(-d0)):6.1f} us (asc / desc)")
    for head in (0.5, 0.6):
        nh = int(4096 * head)
        tail = np.arange(nh, 4096)
        for stride, edges in ((32, (48, 24, 12)), (16, (48, 24, 12))):
            score = np.minimum(S[tail][:, ::stride], 48).max(1)
            cls = np.zeros(tail.size, int)
            for i, e in enumerate(edges):
                cls[score < e] = i + 1
            order = np.concatenate([np.arange(nh), tail[np.argsort(cls, kind='stable')]])
            print(f"first {head:.0%} natural, rest in {len(edges) + 1} probe classes (stride {stride}) "
                  f"{span(order, extra_front=score.astype(np.float64)):6.1f} us")


if __name__ == "__main__":
    main()
