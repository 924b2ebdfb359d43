"""Seeded differential fuzzing of the whole path against the oracle: random map shapes and densities,
resolutions, origins (including non-zero yaw), max ranges, beam counts, fields of view and poses (inside,
on the border, outside, inside walls).  Everything must be bit-identical: occupancy, d^2, the distance
field and every range, through all three entry points."""
import numpy as np
import pytest

from pyracecarsimulator_b200 import maps, range_libc

pytestmark = pytest.mark.gpu


def random_case(rng):
    rows, cols = int(rng.integers(1, 180)), int(rng.integers(1, 220))
    style = rng.integers(3)
    if style == 0:
        occ = (rng.random((rows, cols)) < rng.choice([0.0, 0.002, 0.02, 0.2, 0.7])).astype(np.uint8)
    elif style == 1:                       # rooms: border + random rectangles
        occ = np.zeros((rows, cols), np.uint8)
        occ[0, :] = occ[-1, :] = occ[:, 0] = occ[:, -1] = 1
        for _ in range(int(rng.integers(0, 8))):
            r, c = int(rng.integers(0, rows)), int(rng.integers(0, cols))
            occ[r:r + int(rng.integers(1, 12)), c:c + int(rng.integers(1, 30))] = 1
    else:                                  # thin diagonal walls (tunnelling with the 1 px minimum step)
        occ = np.zeros((rows, cols), np.uint8)
        for k in range(min(rows, cols)):
            occ[k, (k * 3) % cols] = 1
    res = float(rng.choice([0.05, 0.1, 0.025, 1.0, 0.3]))
    origin = (float(rng.uniform(-20, 20)), float(rng.uniform(-20, 20)), float(rng.choice([0.0, 0.0, 0.4, -1.3, 3.0])))
    max_range = float(rng.choice([300.0, 50.0, 7.5, 1.0, 1000.0]))
    return occ, res, origin, max_range


def random_poses(rng, n, rows, cols, res, origin):
    gx = rng.uniform(-0.2 * cols - 2, 1.2 * cols + 2, n) * res      # some outside the map
    gy = rng.uniform(-0.2 * rows - 2, 1.2 * rows + 2, n) * res
    c, s = np.cos(origin[2]), np.sin(origin[2])
    out = np.stack([origin[0] + c * gx - s * gy, origin[1] + s * gx + c * gy, rng.uniform(-7, 7, n)], axis=1)
    return out.astype(np.float32)


@pytest.mark.parametrize("seed", range(12))
def test_random_maps_and_rays(orc, seed):
    rng = np.random.default_rng(1000 + seed)
    for _ in range(6):
        occ, res, origin, max_range = random_case(rng)
        rows, cols = occ.shape
        msg = maps.OccupancyGrid.make(np.where(occ, 100, 0).astype(np.int8).ravel(), cols, rows, res, origin)
        omap = range_libc.PyOMap(msg)
        d2 = orc.edt_exact(occ)
        dist = orc.edt_float(occ)
        assert np.array_equal(omap.occupancy(), occ)
        assert np.array_equal(omap.dist2(), d2)
        assert np.array_equal(omap.dist(), dist)
        rm = range_libc.PyRayMarchingGPU(omap, max_range)
        m = orc.Marcher(dist, max_range, res, origin)
        n = int(rng.integers(1, 400))
        poses = random_poses(rng, n, rows, cols, res, origin)
        # 2-arg form
        out = np.zeros(n, np.float32)
        rm.calc_range_many(poses, out)
        want = m.calc_range_many(poses)
        assert np.array_equal(out, want), (seed, "many", np.flatnonzero(out != want)[:5])
        # fan
        beams = int(rng.choice([1, 2, 7, 60, 270, 1080]))
        fov = float(rng.choice([4.71, 6.2831853, 1.0, 0.01]))
        k = min(n, 40)
        out = np.zeros(k * beams, np.float32)
        rm.calc_range_fan(poses[:k], out, fov, beams)
        want = m.calc_range_fan(poses[:k], beams, fov)
        assert np.array_equal(out, want), (seed, "fan", beams, fov, np.flatnonzero(out != want)[:5])
        # repeat_angles
        angles = rng.uniform(-3.2, 3.2, int(rng.integers(1, 90))).astype(np.float32)
        out = np.zeros(k * angles.size, np.float32)
        rm.calc_range_repeat_angles(poses[:k], angles, out)
        want = m.calc_range_repeat_angles(poses[:k], angles)
        assert np.array_equal(out, want), (seed, "angles", np.flatnonzero(out != want)[:5])
