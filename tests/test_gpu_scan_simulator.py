"""ScanSimulator2D (scripts/scan_simulator.py:11-135) semantics on the GPU implementation."""
import numpy as np
import pytest

from pyracecarsimulator_b200 import maps, range_libc
from pyracecarsimulator_b200.scan_simulator import ScanSimulator2D
from gpu_util import assert_ranges_match

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sim_env(orc, colombia, colombia_scan):
    binar = np.where(colombia_scan["grid"] > 0, 255, 0).ravel()
    msg = maps.OccupancyGrid.make(binar, 435, 350, colombia["resolution"], colombia["origin"])
    omap = range_libc.PyOMap(msg)
    max_range_px = int(15.0 / colombia["resolution"])           # scripts/racecar_simulator_v2.py:196
    dist = orc.edt_float(colombia_scan["occ"])
    return omap, max_range_px, colombia, orc.Marcher(dist, max_range_px, colombia["resolution"], colombia["origin"])


def make_sim(env, method="RMGPU", batch=16):
    omap, mrx, col, _ = env
    sim = ScanSimulator2D(1080, 4.71, 0.01, batch_size=batch)
    sim.setMap(omap, mrx, col["resolution"], col["origin"])
    sim.setRaytracingMethod(method)
    return sim


@pytest.mark.parametrize("method", ["RM", "RMGPU"])
def test_single_scan_config1(sim_env, colombia_scan, method):
    sim = make_sim(sim_env, method)
    out = sim.scan(0.275, 0.0, 0.0)                              # BASELINE config 1
    assert out is sim.output_vector and out.dtype == np.float32 and out.shape == (1080,)
    assert_ranges_match(out, colombia_scan["fan"][:1080], 0.05)
    again = sim.scan(*colombia_scan["poses"][1])
    assert again is out                                          # aliasing is part of the contract
    assert_ranges_match(again, colombia_scan["fan"][1080:2160], 0.05)


def test_scan_many_reads_exactly_batch_size(sim_env, colombia_scan):
    sim = make_sim(sim_env, batch=16)
    poses = colombia_scan["poses"][:20]
    out = sim.scanMany(poses)
    assert out is sim.output_vector_many and out.shape == (16 * 1080,)
    assert_ranges_match(out, colombia_scan["fan"][:16 * 1080], 0.05)
    # list-of-lists, as MCTS passes np.ndarray rows / python lists
    out2 = sim.scanMany([list(map(float, p)) for p in poses]).copy()
    assert np.array_equal(out2, out)
    wide = sim.input_vector_many
    assert wide.shape == (16 * 1080, 3) and np.array_equal(wide[::1080], poses[:16])
    assert not wide[1].any()


def test_errors_raise_instead_of_exiting(sim_env):
    sim = ScanSimulator2D(1080, 4.71, 0.01, batch_size=4)
    with pytest.raises(RuntimeError):
        sim.setRaytracingMethod("RM")
    with pytest.raises(RuntimeError):
        sim.scan(0, 0, 0)
    omap, mrx, col, _ = sim_env
    sim.setMap(omap, mrx, col["resolution"], col["origin"])
    with pytest.raises(ValueError):
        sim.setRaytracingMethod("CDDT")


def test_concurrent_callers_share_one_marcher(sim_env, colombia_scan):
    # rospy calls runScan from several threads without a lock (SURVEY.md 8b threading)
    import threading
    omap, mrx, col, _ = sim_env
    rm = range_libc.PyRayMarchingGPU(omap, mrx)
    poses = colombia_scan["poses"][:8]
    want = colombia_scan["fan"][:8 * 1080]
    errs = []

    def work():
        try:
            out = np.zeros(8 * 1080, np.float32)
            for _ in range(20):
                rm.calc_range_fan(poses, out, 4.71, 1080)
                assert_ranges_match(out, want, 0.05)
        except Exception as e:  # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=work) for _ in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs


def test_racecar_simulator_facade(orc, sim_env, colombia_scan):
    """RacecarSimulator (scripts/racecar_simulator_v2.py): runScan from the lidar pose, updatePose,
    checkCollision / checkCollisionMany against the oracle car + marcher."""
    from pyracecarsimulator_b200.racecar import DEFAULT_CAR_CONFIG
    from pyracecarsimulator_b200.racecar_simulator import RacecarSimulator
    omap, mrx, col, marcher = sim_env
    cfg = dict(DEFAULT_CAR_CONFIG, scan_dist_to_base=0.275, batch_size=20, scan_beams=1080, scan_fov=4.71,
               scan_std=0.01, scan_max_range=15.0, free_thresh=0.8)
    rcs = RacecarSimulator(cfg)
    rcs.setMap(omap, col["resolution"], col["origin"])
    rcs.setRaytracingMethod("RMGPU")
    assert rcs.scan_simulator.mrx == 300
    # zero-state car: the lidar sits at (0.275, 0, 0)  (BASELINE config 1)
    rcs.runScan()
    assert_ranges_match(rcs.getScan(), colombia_scan["fan"][:1080], 0.05)
    assert rcs.checkCollision() == -2                     # one pose, no crash
    # drive: state follows the oracle
    p = orc.car_params()
    want = np.zeros(11)
    rcs.drive(2.0, 0.2)
    for _ in range(30):
        rcs.updatePose()
        orc.car_step(p, want, 2.0, 0.2, 0.01)
    got = rcs.getState()
    assert np.allclose(got, want, rtol=1e-11, atol=1e-13)
    assert rcs.getTravelDistance() == pytest.approx(want[8], rel=1e-11)
    assert rcs.getMeanVelocity() == pytest.approx(want[9] / want[10], rel=1e-11)
    # checkCollisionMany == scanMany + isCrashed
    poses = maps.sample_free_poses(orc.edt_float(colombia_scan["occ"]), 20, 321, col["resolution"], col["origin"],
                                   min_clear_px=1.0)
    edge = orc.car_edge_distances(p, 1080, -4.71 / 2.0, 4.71 / 1080, 0.275)
    want_idx = orc.car_is_crashed(marcher.calc_range_fan(poses, 1080, 4.71), edge, 1080, 20, 0.001)
    assert rcs.checkCollisionMany(poses) == want_idx
    # the reference's side effect on request: every range of the batch in output_vector_many
    rcs.scan_simulator.output_vector_many[:] = -7.0
    assert rcs.checkCollisionMany(poses, want_ranges=True) == want_idx
    assert np.array_equal(rcs.scan_simulator.output_vector_many, marcher.calc_range_fan(poses, 1080, 4.71))
    rcs.stop()
    assert not rcs.getState().any()
