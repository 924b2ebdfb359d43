"""Map ingest on the GPU (csrc/ingest.cu) against the oracle: occupancy and d^2 bit-exact,
fp32 distance field bit-exact (north_star)."""
import numpy as np
import pytest

from pyracecarsimulator_b200 import maps, range_libc

pytestmark = pytest.mark.gpu


def oracle_ingest(orc, img, y):
    grid = orc.mapserver_occupancy(img, y["negate"], y["occupied_thresh"], y["free_thresh"])
    occ = orc.omap_from_grid(grid, True)
    return grid, occ, orc.edt_exact(occ)


def test_colombia_from_image_matches_golden(orc, colombia, colombia_scan, tmp_path):
    path = str(tmp_path / "map.pgm")
    maps.write_pgm(path, colombia["img"])
    y = maps.MapYaml(path, colombia["resolution"], colombia["origin"], colombia["negate"],
                     colombia["occupied_thresh"], colombia["free_thresh"])
    omap = range_libc.PyOMap(y)
    assert (omap.width(), omap.height()) == (350, 435)   # OMap.width = msg rows, height = msg cols
    assert np.array_equal(omap.occupancy(), colombia_scan["occ"])
    assert np.array_equal(omap.dist2(), colombia_scan["d2"])
    assert np.array_equal(omap.dist(), orc.edt_float(colombia_scan["occ"]))
    assert omap.occupancy().sum() == 109212
    assert omap.ingest_ms > 0


def test_colombia_from_message_both_binarisations(orc, colombia, colombia_scan):
    grid = colombia_scan["grid"]                                  # map_server values 100 / 0 / -1
    raw = maps.OccupancyGrid.make(grid.ravel(), 435, 350, 0.05, colombia["origin"])
    assert np.array_equal(range_libc.PyOMap(raw).dist2(), colombia_scan["d2"])
    binar = tuple(255 if v > 0 else 0 for v in grid.ravel().tolist())   # scripts/ros_interface.py:80-86
    msg = maps.OccupancyGrid.make(binar, 435, 350, 0.05, colombia["origin"])
    om = range_libc.PyOMap(msg)
    assert np.array_equal(om.occupancy(), colombia_scan["occ"])
    assert np.array_equal(om.dist2(), colombia_scan["d2"])
    assert om.isOccupied(0, 0) and not om.isOccupied(-1, 0) and not om.isOccupied(0, 435)


@pytest.mark.parametrize("n,seed", [(257, 7), (1025, 11), (2049, 1234)])
def test_synthetic_maps_bit_exact(orc, n, seed):
    img = maps.synth_map(n, seed)
    y = maps.synth_yaml(n)
    grid = orc.mapserver_occupancy(img)
    occ = orc.omap_from_grid(grid, True)
    d2 = orc.edt_exact(occ)
    omap = range_libc.PyOMap(maps.OccupancyGrid.make(grid.ravel(), n, n, y.resolution, y.origin))
    assert np.array_equal(omap.occupancy(), occ)
    assert np.array_equal(omap.dist2(), d2)
    # the reference's float transform is exact at these sizes (<= 2896 px/side)
    assert np.array_equal(omap.dist(), orc.edt_float(occ))


@pytest.mark.parametrize("shape,density,seed", [((1, 1), 1.0, 0), ((1, 1), 0.0, 0), ((1, 300), 0.05, 1),
                                                ((300, 1), 0.05, 2), ((33, 65), 0.5, 3),
                                                ((64, 64), 0.0, 4), ((64, 64), 1.0, 5),
                                                ((257, 129), 0.0005, 6), ((70, 1030), 0.002, 7),
                                                # widths that are multiples of 4 (rows start on 4-byte boundaries)
                                                ((1, 4), 0.5, 8), ((37, 4), 0.2, 9), ((95, 8), 0.1, 10), ((33, 516), 0.01, 11),
                                                ((130, 1024), 0.001, 12), ((64, 128), 0.0, 13), ((31, 2052), 0.3, 14)])
def test_edge_case_grids(orc, shape, density, seed):
    rng = np.random.default_rng(seed)
    occ = (rng.random(shape) < density).astype(np.uint8)
    omap = range_libc.PyOMap(occ.astype(bool))
    assert np.array_equal(omap.occupancy(), occ)
    want = orc.edt_exact(occ)
    assert np.array_equal(omap.dist2(), want)
    assert np.array_equal(omap.dist(), orc.sqrt_dist2(want))
    if not occ.any():
        assert np.all(omap.dist2() == 0x3FFFFFFF) and np.all(omap.dist() == np.float32(1e10))


def test_single_far_obstacle_long_reach(orc):
    occ = np.zeros((40, 3000), np.uint8)
    occ[20, 2999] = 1
    omap = range_libc.PyOMap(occ.astype(bool))
    assert np.array_equal(omap.dist2(), orc.edt_exact(occ))


def test_empty_large_map_is_cheap_and_all_infinite():
    """A map without any occupied cell: every distance is sqrt(1e20) and the row
    pass must not walk the whole row for every cell (rows without a reachable column exit early)."""
    n = 4096
    omap = range_libc.PyOMap(np.zeros((n, n), dtype=bool))
    assert np.all(omap.dist2() == 0x3FFFFFFF)
    assert np.all(omap.dist() == np.float32(1e10))
    assert omap.ingest_ms < 20.0, omap.ingest_ms
    rm = range_libc.PyRayMarchingGPU(omap, 300)
    ins = np.array([[10.0, 10.0, 0.3], [100.0, 50.0, -2.0]], dtype=np.float32)
    outs = np.zeros(2, dtype=np.float32)
    rm.calc_range_many(ins, outs)
    assert np.all(outs == np.float32(300.0))     # resolution 1: max_range px == metres


def test_large_map_exact_integer_edt(orc):
    # above 2896 px/side the float transform is not provably exact: the target is the exact
    # integer EDT; report how many cells the float restatement disagrees on
    n = 4096
    img = maps.synth_map(n, 5678)
    grid = orc.mapserver_occupancy(img)
    occ = orc.omap_from_grid(grid, True)
    d2 = orc.edt_exact(occ)
    omap = range_libc.PyOMap(maps.OccupancyGrid.make(grid.ravel(), n, n, 0.05, (0.0, 0.0, 0.0)))
    assert np.array_equal(omap.dist2(), d2)
    assert np.array_equal(omap.dist(), orc.sqrt_dist2(d2))
    differs = int((orc.edt_float(occ) != omap.dist()).sum())
    print(f"float-Felzenszwalb vs exact EDT on {n}^2: {differs} cells differ")


def test_bad_arguments_raise():
    with pytest.raises(ValueError):
        range_libc.PyOMap(np.zeros((4, 4, 4), bool))
    with pytest.raises(ValueError):
        range_libc.PyOMap(maps.OccupancyGrid.make([0, 0, 0], 2, 2, 0.05, (0, 0, 0)))
    with pytest.raises(ValueError):
        range_libc.PyOMap(np.zeros((2, 20000), bool))   # wider than the supported 16384
    with pytest.raises(ValueError):
        range_libc.PyOMap(42)


@pytest.mark.parametrize("mode,negate", [("trinary", 1), ("scale", 0), ("scale", 1), ("raw", 0)])
def test_map_server_modes(orc, tmp_path, mode, negate):
    """map_server `negate` and `mode: scale / raw`: still one byte LUT on the GPU."""
    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, (90, 130), dtype=np.uint8)
    img[20:30, 40:90] = 3
    path = str(tmp_path / "m.pgm")
    maps.write_pgm(path, img)
    y = maps.MapYaml(path, 0.05, (0.0, 0.0, 0.0), negate, 0.65, 0.196, mode)
    for binarise in (True, False):
        omap = range_libc.PyOMap(y, binarise=binarise)
        grid = orc.mapserver_occupancy(img, negate, 0.65, 0.196, mode)
        occ = orc.omap_from_grid(grid, binarise)
        assert np.array_equal(omap.occupancy(), occ)
        assert np.array_equal(omap.dist2(), orc.edt_exact(occ))
