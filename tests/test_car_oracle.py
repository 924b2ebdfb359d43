"""Vehicle-model restatement (oracle/car_oracle.c) against the unmodified reference Car
(oracle/_ref, built from /root/reference/racecar/src/racecar.cpp) and the golden vectors it
produced."""
import numpy as np
import pytest


def test_survey_known_answer(orc):
    # params.yaml car, zero state, control(2.0, 0.2), 50 x updatePosition(0.01)
    p = orc.car_params()
    st = np.zeros(11)
    for _ in range(50):
        orc.car_step(p, st, 2.0, 0.2, 0.01)
    want = [2.54472164e-01, 3.19658027e-02, 1.44570218e-01, 8.19454863e-01, 2.0e-01, 4.82234064e-01,
            8.76679543e-02, 1, 2.57302660e-01, 2.65497209e+01, 50]
    assert np.allclose(st, want, rtol=2e-8, atol=0)
    assert np.allclose(orc.car_scan_pose(st, 0.275), [0.52660334, 0.07158427, 0.14457022], atol=5e-9)


def _replay(orc, g, t):
    p = orc.car_params()
    st = g["init"][t].copy()
    out = np.zeros_like(g["states"][t])
    poses = np.zeros_like(g["scan_poses"][t])
    for i in range(out.shape[0]):
        sp, sa = g["actions"][t, i // 10]
        orc.car_step(p, st, sp, sa, 0.01)
        out[i] = st
        poses[i] = orc.car_scan_pose(st, 0.275)
    return out, poses


def test_golden_trajectories_bit_exact(orc, car_golden):
    for t in range(car_golden["init"].shape[0]):
        out, poses = _replay(orc, car_golden, t)
        assert np.array_equal(out, car_golden["states"][t]), t
        assert np.array_equal(poses, car_golden["scan_poses"][t]), t


def test_edge_distances_and_crash_index(orc, car_golden):
    p = orc.car_params()
    fov, n = 4.71, 1080
    edge = orc.car_edge_distances(p, n, -fov / 2.0, fov / n, 0.275)
    # golden: smallest float32 range on beam j that the REFERENCE does not call a crash
    safe = car_golden["first_safe_ray"]
    assert np.all(safe.astype(np.float64) - edge >= 0.001)
    below = np.nextafter(safe, np.float32(-np.inf))
    assert np.all(below.astype(np.float64) - edge < 0.001)
    rays = np.full(3 * n, 5.0, np.float32)
    assert orc.car_is_crashed(rays, edge, n, 3, 0.001) == -4     # "no crash" is -(poses+1)
    rays[n + 500] = 0.05
    assert orc.car_is_crashed(rays, edge, n, 3, 0.001) == 1


def test_against_live_reference(orc):
    if not orc.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    p = orc.car_params()
    ref = orc.RefCar(p)
    rng = np.random.default_rng(3)
    for trial in range(40):
        s0 = np.zeros(11)
        s0[:3] = rng.uniform(-5, 5, 3)
        s0[3] = rng.uniform(-1, 7)
        ref.set_state(s0)
        st = s0.copy()
        for i in range(150):
            if i % 10 == 0:
                sp, sa = rng.uniform(0, 7), rng.uniform(-0.4189, 0.4189)
                ref.control(sp, sa)
            ref.update(0.01)
            orc.car_step(p, st, sp, sa, 0.01)
        assert np.array_equal(ref.get_state(), st)
        assert np.array_equal(ref.scan_pose(0.275), orc.car_scan_pose(st, 0.275))
