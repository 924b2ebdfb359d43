"""Helpers shared by the -m gpu parity tests (all calls go through the C ABI via the shim)."""
import numpy as np

from pyracecarsimulator_b200 import maps, range_libc


def build_synth(orc, n, seed, tmp_path_factory=None):
    """Synthetic map ingested on both sides.
    Returns (PyOMap, MapYaml, oracle occupancy, oracle dist)."""
    img = maps.synth_map(n, seed)
    y = maps.synth_yaml(n)
    grid = orc.mapserver_occupancy(img, y.negate, y.occupied_thresh, y.free_thresh)
    occ = orc.omap_from_grid(grid, True)
    dist = orc.sqrt_dist2(orc.edt_exact(occ))
    msg = maps.OccupancyGrid.make(np.where(grid > 0, 255, 0).ravel(), n, n, y.resolution, y.origin)
    omap = range_libc.PyOMap(msg)
    return omap, y, occ, dist


def assert_ranges_match(got, want, resolution, min_identical=1.0):
    """north_star bar: all within max(1e-4 rel, 0.5 cell) and >= 99.9 % of beams bit-identical.
    The device evaluates glibc's sinf/cosf (csrc/glibc_trig.cuh) and the march is IEEE-exact
    fp32 in the oracle's order, so the tests ask for more: every beam bit-identical."""
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape
    tol = np.maximum(1e-4 * np.abs(want), 0.5 * resolution)
    bad = np.abs(got.astype(np.float64) - want.astype(np.float64)) > tol
    assert not bad.any(), f"{bad.sum()} of {bad.size} ranges outside tolerance; first at {np.flatnonzero(bad)[:5]}"
    same = np.mean(got == want) if got.size else 1.0
    assert same >= min_identical, f"only {same:.6f} bit-identical"
    return same
