"""Checks on the oracle that do not depend on the recalled range_libc details being right
(SURVEY.md section 4 (3)): brute-force and scipy EDT, analytic box room, a DDA caster,
agreement of the three entry points, symmetry."""
import numpy as np
import pytest

from pyracecarsimulator_b200 import maps


def brute_d2(occ):
    rows, cols = occ.shape
    pts = np.argwhere(occ != 0)
    rr, cc = np.mgrid[0:rows, 0:cols]
    if len(pts) == 0:
        return np.full(occ.shape, 0x3FFFFFFF, np.int64)
    d = (rr[..., None] - pts[:, 0]) ** 2 + (cc[..., None] - pts[:, 1]) ** 2
    return d.min(axis=-1)


@pytest.mark.parametrize("shape,density,seed", [((1, 1), 1.0, 0), ((1, 17), 0.2, 1), ((23, 1), 0.2, 2),
                                                ((31, 45), 0.02, 3), ((64, 64), 0.001, 4),
                                                ((40, 33), 0.5, 5), ((50, 60), 0.0, 6)])
def test_edt_exact_vs_brute_force(orc, shape, density, seed):
    rng = np.random.default_rng(seed)
    occ = (rng.random(shape) < density).astype(np.uint8)
    want = brute_d2(occ)
    got = orc.edt_exact(occ).astype(np.int64)
    assert np.array_equal(got, want)
    dist, d2f = orc.edt_float(occ, want_dist2=True)
    if occ.any():
        assert np.array_equal(d2f.astype(np.int64), want)
        assert np.array_equal(dist, np.sqrt(want.astype(np.float32)))
    else:
        assert np.all(dist == np.float32(1e10))  # sqrt(1e20f), SURVEY.md A.3
        assert np.array_equal(orc.sqrt_dist2(got.astype(np.int32)), dist)


def test_edt_vs_scipy_on_synthetic_map(orc):
    ndimage = pytest.importorskip("scipy.ndimage")
    img = maps.synth_map(513, 1234)
    occ = orc.omap_from_grid(orc.mapserver_occupancy(img), True)
    want = ndimage.distance_transform_edt(occ == 0).astype(np.float32)
    assert np.array_equal(orc.edt_float(occ), want)
    assert np.array_equal(orc.sqrt_dist2(orc.edt_exact(occ)), want)


def box_room(n=201, wall=3):
    occ = np.zeros((n, n), np.uint8)
    occ[:wall, :] = occ[-wall:, :] = occ[:, :wall] = occ[:, -wall:] = 1
    return occ


def test_box_room_axis_rays(orc):
    # pose in the middle of a square room, resolution 1, origin 0: axis-aligned rays hit the
    # wall's first cell; range is measured to that cell's corner (A.6)
    occ = box_room()
    m = orc.Marcher(orc.edt_float(occ), 500, 1.0)
    x = y = 100.5
    for th, want_cell in [(0.0, 198), (np.pi, 2), (np.pi / 2, 198), (-np.pi / 2, 2)]:
        r = m.calc_range(x, y, th)
        true_surface = 98.5 if want_cell == 198 else 98.5 - 1  # wall faces at 198.0 and 3.0
        assert abs(r - abs(want_cell - 100.5)) < 1.5, (th, r)
        assert abs(r - true_surface) <= 2.0


def dda_range(occ, x0, y0, th, max_range):
    """Independent caster: walk the ray in 0.01 px steps until an occupied cell."""
    dx, dy = np.cos(th), np.sin(th)
    t = np.arange(0, max_range, 0.01)
    cx = np.floor(x0 + dx * t).astype(int)
    cy = np.floor(y0 + dy * t).astype(int)
    ok = (cx >= 0) & (cy >= 0) & (cx < occ.shape[1]) & (cy < occ.shape[0])
    hit = np.zeros_like(ok)
    hit[ok] = occ[cy[ok], cx[ok]] != 0
    stop = np.flatnonzero(hit | ~ok)
    if len(stop) == 0 or not ok[stop[0]]:
        return max_range
    return t[stop[0]]


def test_against_dda_caster(orc):
    # thick walls only (the 1 px minimum step can tunnel thin diagonal walls, A.6)
    rng = np.random.default_rng(5)
    occ = box_room(301, 4)
    for _ in range(12):
        r, c = rng.integers(20, 270, 2)
        occ[r:r + rng.integers(6, 30), c:c + rng.integers(6, 30)] = 1
    dist = orc.edt_float(occ)
    m = orc.Marcher(dist, 600, 1.0)
    free = np.argwhere(dist > 4)
    bad = 0
    for i in range(300):
        row, col = free[rng.integers(len(free))]
        th = rng.uniform(-np.pi, np.pi)
        got = m.calc_range(col + 0.5, row + 0.5, th)
        want = dda_range(occ, col + 0.5, row + 0.5, th, 600)
        # the last step may land up to 1 px inside the wall (minimum step) and the hit range is
        # measured to that cell's integer corner (up to sqrt(2) px more): under 2.5 px in all
        if abs(got - want) > 2.5:
            bad += 1
    assert bad <= 3, bad  # grazing rays may legitimately differ


def test_entry_points_agree(orc, colombia, colombia_dist, colombia_scan):
    _, dist = colombia_dist
    m = orc.Marcher(dist, 300, colombia["resolution"], colombia["origin"])
    poses = colombia_scan["poses"][:6]
    n, fov = 90, 4.71
    fan = m.calc_range_fan(poses, n, fov).reshape(6, n)
    inc = np.float32(fov) / np.float32(n)
    a = (np.arange(n, dtype=np.float32) * inc + np.float32(-0.5) * np.float32(fov)).astype(np.float32)
    rows = np.repeat(poses, n, axis=0)
    rows[:, 2] = (poses[:, 2:3] + a[None, :]).ravel()
    many = m.calc_range_many(rows).reshape(6, n)
    rep = m.calc_range_repeat_angles(poses, a).reshape(6, n)
    tol = np.maximum(1e-4 * fan, 0.5 * colombia["resolution"])
    # same rays up to the last ulp of the heading: identical except where a cell boundary flips
    assert np.mean(np.abs(many - fan) <= tol) > 0.995
    assert np.mean(np.abs(rep - fan) <= tol) > 0.995
    assert np.mean(many == fan) > 0.98


def test_max_range_and_out_of_map(orc):
    occ = np.zeros((50, 50), np.uint8)
    occ[25, 40] = 1
    m = orc.Marcher(orc.edt_float(occ), 300, 0.05)
    assert m.calc_range(-10.0, 1.0, 0.3) == np.float32(300 * np.float32(0.05))   # starts outside
    assert m.calc_range(1.0, 1.0, np.pi) == np.float32(15.0)                      # leaves the map
    assert np.isfinite(m.calc_range(np.nan, 1.0, 0.0))
    # start inside the occupied cell: distance to its own corner, not 0 (A.6)
    r = m.calc_range(40.5 * 0.05, 25.5 * 0.05, 0.0)
    assert 0 < r < 1.5 * 0.05
    # truncation toward zero: x in (-1, 0) px is in bounds
    empty = np.zeros((20, 20), np.uint8)
    empty[:, 10] = 1
    m2 = orc.Marcher(orc.edt_float(empty), 100, 1.0)
    assert m2.calc_range(-0.5, 5.5, 0.0) < 12


def test_empty_map_is_all_max_range(orc):
    m = orc.Marcher(orc.edt_float(np.zeros((30, 30), np.uint8)), 300, 0.05)
    out = m.calc_range_fan(np.array([[0.7, 0.7, 0.1]], np.float32), 64, 6.28)
    assert np.all(out == np.float32(15.0))


def test_synth_map_is_deterministic():
    import hashlib
    a = maps.synth_map(257, 7)
    assert a.shape == (257, 257) and a.dtype == np.uint8
    assert set(np.unique(a)) <= {0, 205, 254}
    assert hashlib.sha256(a.tobytes()).hexdigest() == hashlib.sha256(maps.synth_map(257, 7).tobytes()).hexdigest()
    assert (maps.synth_map(2049, 1234) == 0).mean() == pytest.approx(0.029, abs=0.003)  # Appendix D
