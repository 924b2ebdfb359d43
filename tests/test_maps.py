"""Host-side file formats either side of the path: PGM (P2 with comments, P5), map_server yaml,
the OccupancyGrid-shaped message."""
import math
import os

import numpy as np
import pytest

from img_util import write_png as _write_png
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


# ------------------------------------------------------------------ colour / alpha / deep images
@pytest.mark.parametrize("ctype,ch", [(0, 1), (2, 3), (4, 2), (6, 4)])
def test_png_round_trip_all_filters(tmp_path, ctype, ch):
    rng = np.random.default_rng(ctype)
    img = rng.integers(0, 256, (13, 9) if ch == 1 else (13, 9, ch), dtype=np.uint8)
    path = str(tmp_path / "m.png")
    _write_png(path, img, ctype)
    got, has_alpha = maps.read_png(path)
    assert np.array_equal(got, img) and has_alpha == (ctype in (4, 6))
    assert np.array_equal(maps.read_image(path)[0], img)


def test_png_16_bit_and_packed_grey(tmp_path):
    rng = np.random.default_rng(5)
    deep = rng.integers(0, 65536, (6, 7), dtype=np.uint16)
    path = str(tmp_path / "d.png")
    _write_png(path, deep, 0, depth=16)
    got, _ = maps.read_png(path)
    assert np.array_equal(got, ((deep.astype(np.int64) * 255 + 32767) // 65535).astype(np.uint8))
    for depth in (1, 2, 4):
        img = rng.integers(0, 1 << depth, (5, 11), dtype=np.uint8)
        _write_png(path, img, 0, depth=depth, filters=(0,))
        got, _ = maps.read_png(path)
        assert np.array_equal(got, (img.astype(np.int64) * 255 // ((1 << depth) - 1)).astype(np.uint8))


def test_ppm_and_deep_pgm(tmp_path):
    rng = np.random.default_rng(1)
    rgb = rng.integers(0, 256, (4, 5, 3), dtype=np.uint8)
    p6 = tmp_path / "a.ppm"
    p6.write_bytes(b"P6\n5 4\n255\n" + rgb.tobytes())
    assert np.array_equal(maps.read_pnm(str(p6)), rgb)
    p3 = tmp_path / "b.ppm"
    p3.write_text("P3\n# c\n5 4\n255\n" + " ".join(str(v) for v in rgb.ravel()) + "\n")
    assert np.array_equal(maps.read_image(str(p3))[0], rgb)
    deep = rng.integers(0, 1024, (3, 4), dtype=np.uint16)              # maxval 1023, big-endian samples
    p5 = tmp_path / "c.pgm"
    p5.write_bytes(b"P5\n4 3\n1023\n" + deep.astype(">u2").tobytes())
    assert np.array_equal(maps.read_pgm(str(p5)), ((deep.astype(np.int64) * 255 + 511) // 1023).astype(np.uint8))
    low = rng.integers(0, 16, (3, 4), dtype=np.uint8)                  # maxval 15: SDL's v*255/maxval
    p5.write_bytes(b"P5\n4 3\n15\n" + low.tobytes())
    assert np.array_equal(maps.read_pgm(str(p5)), (low.astype(np.int64) * 255 // 15).astype(np.uint8))
    with pytest.raises(ValueError):
        maps.read_pgm(str(p6))                                          # colour where grey is required


def test_oracle_channel_averaging_rules(orc):
    """map_server's per-pixel rules on colour / alpha images, against a straight numpy restatement."""
    rng = np.random.default_rng(3)
    for ch, has_alpha in ((2, True), (3, False), (4, True), (4, False)):
        img = rng.integers(0, 256, (17, 19, ch), dtype=np.uint8)
        img[rng.random((17, 19)) < 0.3, ch - 1] = 0                    # last byte == 0: the alpha rule of scale mode
        for mode in ("trinary", "scale", "raw"):
            for negate in (0, 1):
                got = orc.mapserver_occupancy_channels(img, has_alpha, negate, 0.65, 0.196, mode)
                avg = ch if (mode == "trinary" or not has_alpha) else ch - 1
                ca = img[:, :, :avg].astype(np.int64).sum(axis=2) / float(avg)
                if negate:
                    ca = 255 - ca
                if mode == "raw":
                    want = ca.astype(np.uint8).astype(np.int8)
                else:
                    occ = (255 - ca) / 255.0
                    ratio = (occ - 0.196) / (0.65 - 0.196)
                    scale = (1 + 98 * ratio).astype(np.int64).astype(np.uint8).astype(np.int8)
                    mid = np.where((mode == "trinary") | (img[:, :, ch - 1] < 1), -1, scale)
                    want = np.where(occ > 0.65, 100, np.where(occ < 0.196, 0, mid)).astype(np.int8)
                assert np.array_equal(got, want[::-1]), (ch, has_alpha, mode, negate)
    # a grey image through the channel path == the grey path
    g = rng.integers(0, 256, (9, 8), dtype=np.uint8)
    for mode in ("trinary", "scale", "raw"):
        for negate in (0, 1):
            assert np.array_equal(orc.mapserver_occupancy_channels(g[:, :, None], False, negate, 0.65, 0.196, mode),
                                  orc.mapserver_occupancy(g, negate, 0.65, 0.196, mode))
