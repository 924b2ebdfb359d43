"""Host-side file formats either side of the path: PGM (P2 with comments, P5), map_server yaml,
the OccupancyGrid-shaped message."""
import math
import os

import numpy as np
import pytest

from pyracecarsimulator_b200 import maps


def test_pgm_round_trip_binary_and_ascii(tmp_path):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (7, 11), dtype=np.uint8)
    p5 = tmp_path / "a.pgm"
    maps.write_pgm(str(p5), img)
    assert np.array_equal(maps.read_pgm(str(p5)), img)
    p2 = tmp_path / "b.pgm"
    body = "\n".join(" ".join(f"{v:3d}" for v in row) for row in img)
    p2.write_text("P2\n# 8-bit pgm gray\n11 7\n255\n" + body + "\n")   # colombia's header style
    assert np.array_equal(maps.read_pgm(str(p2)), img)
    bad = tmp_path / "c.pgm"
    bad.write_bytes(b"P6\n1 1\n255\n\0\0\0")
    with pytest.raises(ValueError):
        maps.read_pgm(str(bad))


def test_reference_colombia_pgm_matches_fixture(colombia):
    path = "/root/reference/maps/colombia/map.yaml"
    if not os.path.exists(path):
        pytest.skip("reference checkout not on this box")
    y = maps.load_map_yaml(path)
    assert (y.resolution, y.negate, y.occupied_thresh, y.free_thresh) == (0.05, 0, 0.65, 0.196)
    assert y.origin == pytest.approx((-5.70654, -2.020793, 0.0))
    assert np.array_equal(maps.read_pgm(y.image), colombia["img"])


def test_occupancy_grid_message_and_yaw():
    msg = maps.OccupancyGrid.make([0] * 6, 3, 2, 0.05, (1.0, -2.0, 0.7))
    assert (msg.info.width, msg.info.height) == (3, 2)
    assert maps.quaternion_to_yaw(msg.info.origin.orientation) == pytest.approx(0.7)
    assert maps.quaternion_to_yaw(maps.OccupancyGrid.make([], 0, 0, 1, (0, 0, 0)).info.origin.orientation) == 0.0
    assert math.isclose(msg.info.origin.position.y, -2.0)


def test_sample_free_poses_are_free_and_seeded(orc):
    img = maps.synth_map(257, 7)
    occ = orc.omap_from_grid(orc.mapserver_occupancy(img), True)
    dist = orc.edt_float(occ)
    y = maps.synth_yaml(257)
    a = maps.sample_free_poses(dist, 100, 5, y.resolution, y.origin)
    b = maps.sample_free_poses(dist, 100, 5, y.resolution, y.origin)
    assert a.dtype == np.float32 and a.shape == (100, 3) and np.array_equal(a, b)
    col = np.floor((a[:, 0].astype(np.float64) - y.origin[0]) / y.resolution).astype(int)
    row = np.floor((a[:, 1].astype(np.float64) - y.origin[1]) / y.resolution).astype(int)
    assert np.all(dist[np.clip(row, 0, 256), np.clip(col, 0, 256)] > 1.5)
    assert np.all(np.abs(a[:, 2]) <= np.float32(math.pi))
    assert maps.synth_yaml(2049).origin[0] == -51.224998   # maps/map.yaml:3
