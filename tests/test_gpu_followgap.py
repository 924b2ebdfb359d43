"""Batched follow-the-gap kernel (csrc/followgap.cu) against the oracle, bit for bit."""
import numpy as np
import pytest

from test_followgap_oracle import scans

pytestmark = pytest.mark.gpu


def test_eval_many_matches_oracle(orc):
    import torch
    from pyracecarsimulator_b200.followgap import PyFollowGap
    fg = PyFollowGap(10, 15.0, 0.4189, 0.004)
    rng = np.random.default_rng(1)
    for n in (10, 33, 270, 1080, 1081, 4096):
        batch = []
        for l in scans(rng, 200):
            batch.append(np.resize(l, n))
        # structured cases: all near, all far, a single far beam at the very end, zeros
        batch += [np.full(n, 1.0, np.float32), np.full(n, 9.0, np.float32), np.full(n, 30.0, np.float32)]
        last = np.full(n, 1.0, np.float32); last[-1] = 5.0
        first = np.full(n, 1.0, np.float32); first[0] = 5.0
        zeros = np.zeros(n, np.float32); zeros[n // 2] = 3.0
        batch += [last, first, zeros]
        arr = np.ascontiguousarray(np.stack(batch).astype(np.float32))
        got = fg.eval_many(torch.from_numpy(arr).cuda()).cpu().numpy()
        want = np.array([orc.followgap_eval(r) for r in arr], np.float32)
        same = (got == want) | (np.isnan(got) & np.isnan(want))
        assert same.all(), (n, np.flatnonzero(~same)[:5], got[~same][:5], want[~same][:5])
    l = np.full(1080, 3.0, np.float32)
    l[200:300] = 1.0
    assert fg.eval(l, 1080) == 0.4000000059604645


def test_scan_then_follow_gap_on_device(orc, colombia, colombia_scan):
    import torch
    from pyracecarsimulator_b200 import maps, range_libc
    from pyracecarsimulator_b200.followgap import PyFollowGap
    binar = np.where(colombia_scan["grid"] > 0, 255, 0).ravel()
    omap = range_libc.PyOMap(maps.OccupancyGrid.make(binar, 435, 350, colombia["resolution"], colombia["origin"]))
    rm = range_libc.PyRayMarchingGPU(omap, 300)
    poses = torch.from_numpy(colombia_scan["poses"]).cuda()
    out = torch.empty(poses.shape[0] * 1080, dtype=torch.float32, device="cuda")
    rm.calc_range_fan(poses, out, 4.71, 1080)
    got = PyFollowGap(10, 15.0, 0.4189, 4.71 / 1080).eval_many(out.reshape(-1, 1080)).cpu().numpy()
    want = np.array([orc.followgap_eval(colombia_scan["fan"][i * 1080:(i + 1) * 1080], 15.0, 0.4189, 4.71 / 1080)
                     for i in range(poses.shape[0])], np.float32)
    assert np.array_equal(got, want)
