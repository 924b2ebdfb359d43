"""The CPU oracle against every pin available for it: the survey's known answers on
maps/colombia and the committed golden vectors."""
import numpy as np


def test_colombia_ingest_counts(orc, colombia):
    # colombia -> 109 212 occupied, 43 024 free, 14 unknown
    grid = orc.mapserver_occupancy(colombia["img"], colombia["negate"], colombia["occupied_thresh"],
                                   colombia["free_thresh"])
    assert (grid == 100).sum() == 109212
    assert (grid == 0).sum() == 43024
    assert (grid == -1).sum() == 14
    occ = orc.omap_from_grid(grid, True)
    assert occ.sum() == 109212  # unknown and free are both free after binarisation


def test_threshold_boundaries(orc):
    # occupied <=> p <= 89, free <=> p >= 206, unknown 90..205 (thresholds .65/.196)
    img = np.arange(256, dtype=np.uint8).reshape(1, 256)
    g = orc.mapserver_occupancy(img)[0]
    assert np.all(g[:90] == 100) and np.all(g[90:206] == -1) and np.all(g[206:] == 0)
    gn = orc.mapserver_occupancy(img, negate=1)[0]
    assert np.array_equal(gn, g[::-1])


def test_mapserver_flips_rows(orc):
    img = np.full((3, 2), 254, np.uint8)
    img[0, 1] = 0  # top image row
    g = orc.mapserver_occupancy(img)
    assert g[2, 1] == 100 and (g == 100).sum() == 1


def test_golden_ingest(orc, colombia, colombia_scan):
    grid = orc.mapserver_occupancy(colombia["img"], colombia["negate"], colombia["occupied_thresh"],
                                   colombia["free_thresh"])
    assert np.array_equal(grid, colombia_scan["grid"])
    occ = orc.omap_from_grid(grid, True)
    assert np.array_equal(occ, colombia_scan["occ"])
    assert np.array_equal(orc.edt_exact(occ), colombia_scan["d2"])
    dist, d2f = orc.edt_float(occ, want_dist2=True)
    # float Felzenszwalb == exact integer EDT on a map this size
    assert np.array_equal(d2f.astype(np.int64), colombia_scan["d2"].astype(np.int64))
    assert np.array_equal(dist, orc.sqrt_dist2(colombia_scan["d2"]))
    assert colombia_scan["d2"].max() == 2113  # Appendix C


def test_survey_known_answer_scan(orc, colombia, colombia_dist):
    # colombia, world pose (0.275, 0, 0), 1080 beams, fov 4.71, max 300 px
    _, dist = colombia_dist
    m = orc.Marcher(dist, 300, colombia["resolution"], colombia["origin"])
    out, steps = m.calc_range_fan(np.array([[0.275, 0, 0]], np.float32), 1080, 4.71, steps=True)
    want = {0: 1.2391100, 270: 1.1963634, 540: 3.7185178, 810: 1.8073667, 1079: 3.5077786}
    for j, v in want.items():
        assert abs(out[j] - v) < 2e-7 * max(1, v), (j, out[j])
    assert abs(out.min() - 1.1067855) < 2e-7
    assert out.max() == np.float32(15.0)
    assert abs(out.sum(dtype=np.float64) - 3282.2590) < 1e-3
    assert steps.sum() == 7541


def test_golden_scans(orc, colombia, colombia_dist, colombia_scan):
    _, dist = colombia_dist
    g = colombia_scan
    m = orc.Marcher(dist, 300, colombia["resolution"], colombia["origin"])
    fan, fs = m.calc_range_fan(g["poses"], 1080, 4.71, steps=True)
    assert np.array_equal(fan, g["fan"]) and np.array_equal(fs, g["fan_steps"])
    many, ms = m.calc_range_many(g["rays"], steps=True)
    assert np.array_equal(many, g["many"]) and np.array_equal(ms, g["many_steps"])
    rep, rs = m.calc_range_repeat_angles(g["poses"], g["angles"], steps=True)
    assert np.array_equal(rep, g["rep"]) and np.array_equal(rs, g["rep_steps"])


def test_threads_do_not_change_results(orc, colombia, colombia_dist, colombia_scan):
    _, dist = colombia_dist
    g = colombia_scan
    m = orc.Marcher(dist, 300, colombia["resolution"], colombia["origin"])
    assert np.array_equal(m.calc_range_fan(g["poses"], 1080, 4.71, threads=0), g["fan"])
    assert np.array_equal(m.calc_range_many(g["rays"], threads=0), g["many"])
    assert np.array_equal(m.calc_range_repeat_angles(g["poses"], g["angles"], threads=3), g["rep"])


def test_reference_layout_equals_compact(orc, colombia, colombia_dist, colombia_scan):
    # the fork's (B*num_rays, 3) layout with the pose in row k*num_rays (scripts/scan_simulator.py:119-127)
    _, dist = colombia_dist
    m = orc.Marcher(dist, 300, colombia["resolution"], colombia["origin"])
    poses = colombia_scan["poses"][:5]
    wide = np.zeros((5 * 1080, 3), np.float32)
    wide[::1080] = poses
    wide[1::1080] = 123.0  # dead rows must be ignored
    a = m.calc_range_fan(wide, 1080, 4.71, pose_stride_rows=1080)
    assert np.array_equal(a, colombia_scan["fan"][:5 * 1080])
