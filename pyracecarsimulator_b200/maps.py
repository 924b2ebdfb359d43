"""Map files either side of the scan path: ROS ``map_server`` yaml + PGM parsing, the
OccupancyGrid-shaped message the reference hands to ``range_libc.PyOMap``
(scripts/ros_interface.py:77-87, :202-223), and the deterministic synthetic maps that stand
in for the image blobs missing from the reference checkout.

Host-side file parsing only -- thresholding, y-flip, binarisation and the distance
transform run on the GPU (csrc/ingest.cu).
"""
from __future__ import annotations

import math
import os
import re
from dataclasses import dataclass, field

import numpy as np
import yaml


# --------------------------------------------------------------------------- messages
@dataclass
class _Position:
    x: float = 0.0
    y: float = 0.0
    z: float = 0.0


@dataclass
class _Quaternion:
    x: float = 0.0
    y: float = 0.0
    z: float = 0.0
    w: float = 1.0


@dataclass
class _Pose:
    position: _Position = field(default_factory=_Position)
    orientation: _Quaternion = field(default_factory=_Quaternion)


@dataclass
class MapMetaData:
    width: int = 0
    height: int = 0
    resolution: float = 0.05
    origin: _Pose = field(default_factory=_Pose)


@dataclass
class OccupancyGrid:
    """Duck-type of ``nav_msgs/OccupancyGrid`` as used at scripts/ros_interface.py:210-220:
    ``info.width/height/resolution/origin.position.{x,y}/origin.orientation.{x,y,z,w}`` and
    ``data`` row-major from the bottom-left cell."""
    info: MapMetaData = field(default_factory=MapMetaData)
    data: object = None

    @staticmethod
    def make(data, width, height, resolution, origin_xyyaw):
        ox, oy, yaw = origin_xyyaw
        q = _Quaternion(0.0, 0.0, math.sin(yaw / 2.0), math.cos(yaw / 2.0))
        info = MapMetaData(int(width), int(height), float(resolution),
                           _Pose(_Position(float(ox), float(oy), 0.0), q))
        return OccupancyGrid(info, data)


def quaternion_to_yaw(q) -> float:
    """Yaw of a quaternion (the ``euler_from_quaternion(...)[2]`` of scripts/ros_interface.py:216)."""
    return math.atan2(2.0 * (q.w * q.z + q.x * q.y), 1.0 - 2.0 * (q.y * q.y + q.z * q.z))


# --------------------------------------------------------------------------- files
def read_pgm(path: str) -> np.ndarray:
    """Read a binary (P5) or ASCII (P2) 8-bit PGM -> (H, W) uint8, rows top to bottom.
    maps/colombia/map.pgm is P2 with a ``#`` comment line."""
    with open(path, "rb") as f:
        raw = f.read()
    magic = raw[:2]
    if magic not in (b"P2", b"P5"):
        raise ValueError(f"{path}: not a PGM (magic {magic!r})")
    # header: magic, width, height, maxval separated by whitespace, '#' comments to end of line
    pos, vals = 2, []
    while len(vals) < 3:
        m = re.compile(rb"\s*(#[^\n]*\n|\d+)").match(raw, pos)
        if m is None:
            raise ValueError(f"{path}: malformed PGM header")
        pos = m.end()
        if not m.group(1).startswith(b"#"):
            vals.append(int(m.group(1)))
    w, h, maxval = vals
    if maxval > 255:
        raise ValueError(f"{path}: 16-bit PGM not supported")
    if magic == b"P5":
        pos += 1  # single whitespace byte after maxval
        img = np.frombuffer(raw, dtype=np.uint8, count=w * h, offset=pos)
    else:
        body = re.sub(rb"#[^\n]*", b"", raw[pos:])
        img = np.array(body.split(), dtype=np.int64)
        if img.size != w * h:
            raise ValueError(f"{path}: expected {w*h} samples, found {img.size}")
        img = img.astype(np.uint8)
    return np.ascontiguousarray(img.reshape(h, w))


def write_pgm(path: str, img: np.ndarray) -> None:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(img.tobytes())


@dataclass
class MapYaml:
    image: str
    resolution: float
    origin: tuple
    negate: int = 0
    occupied_thresh: float = 0.65
    free_thresh: float = 0.196
    mode: str = "trinary"


def load_map_yaml(path: str) -> MapYaml:
    """Parse a map_server yaml (maps/map.yaml:1-6)."""
    with open(path) as f:
        y = yaml.safe_load(f)
    image = y["image"]
    if not os.path.isabs(image):
        image = os.path.join(os.path.dirname(os.path.abspath(path)), image)
    return MapYaml(image, float(y["resolution"]), tuple(float(v) for v in y["origin"]),
                   int(y.get("negate", 0)), float(y.get("occupied_thresh", 0.65)),
                   float(y.get("free_thresh", 0.196)), str(y.get("mode", "trinary")))


# --------------------------------------------------------------------------- synthetic maps
def synth_map(n: int, seed: int) -> np.ndarray:
    """Deterministic synthetic map image: 2-px border wall, random
    axis-aligned wall segments 1-3 px thick, unknown (205) patches.  Image-row order
    (row 0 = top), uint8 with 0 = occupied, 254 = free, 205 = unknown."""
    rng = np.random.default_rng(seed)
    img = np.full((n, n), 254, dtype=np.uint8)
    img[:2, :] = 0
    img[-2:, :] = 0
    img[:, :2] = 0
    img[:, -2:] = 0
    k = n * n // 20000
    for _ in range(k):
        horiz = rng.integers(2)
        length = rng.integers(n // 32, n // 4)
        th = rng.integers(1, 4)
        r = rng.integers(0, n)
        c = rng.integers(0, n)
        if horiz:
            img[r:r + th, c:c + length] = 0
        else:
            img[r:r + length, c:c + th] = 0
    for _ in range(k // 4):
        s = rng.integers(4, 32)
        r = rng.integers(0, n - s)
        c = rng.integers(0, n - s)
        patch = img[r:r + s, c:c + s]
        patch[patch == 254] = 205
    return img


def synth_yaml(n: int) -> MapYaml:
    """Metadata of the stand-in maps: resolution 0.05; for n = 2049 the shipped origin of
    maps/map.yaml:3 (-51.224998), otherwise -(n/2)*0.05."""
    o = -51.224998 if n == 2049 else -(n / 2.0) * 0.05
    return MapYaml(f"synth_{n}.pgm", 0.05, (o, o, 0.0))


def sample_free_poses(dist: np.ndarray, n: int, seed: int, resolution: float, origin,
                      min_clear_px: float = 3.0) -> np.ndarray:
    """Seeded (n, 3) fp32 world poses over cells with DT > min_clear_px:
    sub-cell jitter U[0,1), theta U[-pi, pi); x = (col+u)*res + ox, y = (row+v)*res + oy."""
    rng = np.random.default_rng(seed)
    free = np.flatnonzero(np.asarray(dist).ravel() > min_clear_px)
    if free.size == 0:
        raise ValueError("map has no cell clear of obstacles")
    pick = free[rng.integers(0, free.size, size=n)]
    cols = dist.shape[1]
    row, col = pick // cols, pick % cols
    u, v = rng.random(n), rng.random(n)
    th = rng.uniform(-math.pi, math.pi, n)
    out = np.empty((n, 3), dtype=np.float32)
    out[:, 0] = (col + u) * resolution + origin[0]
    out[:, 1] = (row + v) * resolution + origin[1]
    out[:, 2] = th
    return out
