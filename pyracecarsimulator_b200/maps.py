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
def _pnm_header(raw: bytes, path: str, count: int):
    """Values after the magic number of a PNM file: whitespace separated, '#' comments to end of line."""
    pos, vals = 2, []
    while len(vals) < count:
        m = re.compile(rb"\s*(#[^\n]*\n|\d+)").match(raw, pos)
        if m is None:
            raise ValueError(f"{path}: malformed PNM header")
        pos = m.end()
        if not m.group(1).startswith(b"#"):
            vals.append(int(m.group(1)))
    return vals, pos


def _narrow(samples: np.ndarray, maxval: int) -> np.ndarray:
    """Samples with an arbitrary maxval -> 8 bits.  SDL_image (ROS1 map_server) rescales maxval < 255 as
    v*255/maxval and refuses maxval > 255; deeper files are narrowed the same way here (rounded), which is
    what an 8-bit export of the same picture would hold."""
    if maxval == 255:
        return samples.astype(np.uint8)
    if maxval < 255:
        return (samples.astype(np.int64) * 255 // maxval).astype(np.uint8)
    return ((samples.astype(np.int64) * 255 + maxval // 2) // maxval).astype(np.uint8)


def read_pnm(path: str) -> np.ndarray:
    """PGM (P2 ASCII / P5 binary) or PPM (P3 / P6), any maxval up to 65535 -> (H, W) or (H, W, 3) uint8,
    rows top to bottom.  maps/colombia/map.pgm is P2 with a ``#`` comment line."""
    with open(path, "rb") as f:
        raw = f.read()
    magic = raw[:2]
    if magic not in (b"P2", b"P5", b"P3", b"P6"):
        raise ValueError(f"{path}: not a PGM/PPM (magic {magic!r})")
    (w, h, maxval), pos = _pnm_header(raw, path, 3)
    if not 0 < maxval < 65536:
        raise ValueError(f"{path}: bad maxval {maxval}")
    ch = 3 if magic in (b"P3", b"P6") else 1
    n = w * h * ch
    if magic in (b"P5", b"P6"):
        pos += 1  # single whitespace byte after maxval
        dt = np.dtype(">u2") if maxval > 255 else np.dtype(np.uint8)
        if len(raw) - pos < n * dt.itemsize:
            raise ValueError(f"{path}: expected {n} samples, file is too short")
        img = np.frombuffer(raw, dtype=dt, count=n, offset=pos)
    else:
        body = re.sub(rb"#[^\n]*", b"", raw[pos:])
        img = np.array(body.split(), dtype=np.int64)
        if img.size != n:
            raise ValueError(f"{path}: expected {n} samples, found {img.size}")
    img = _narrow(img, maxval)
    return np.ascontiguousarray(img.reshape((h, w, 3) if ch == 3 else (h, w)))


def read_pgm(path: str) -> np.ndarray:
    """Grey PNM -> (H, W) uint8 (see :func:`read_pnm`)."""
    img = read_pnm(path)
    if img.ndim != 2:
        raise ValueError(f"{path}: colour image where a grey one was expected")
    return img


def read_png(path: str):
    """Minimal PNG reader (zlib is in the standard library; map_server loads PNGs through SDL_image):
    non-interlaced, bit depth 8 or 16 (and 1/2/4 for grey / palette), colour types 0 grey, 2 RGB, 3 palette
    (returned as palette INDICES, which is what map_server reads from SDL's 8-bit indexed surface),
    4 grey+alpha, 6 RGBA.  Returns ((H, W) or (H, W, C) uint8, has_alpha)."""
    import struct
    import zlib
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError(f"{path}: not a PNG")
    pos, idat, hdr = 8, [], None
    while pos + 8 <= len(raw):
        length, kind = struct.unpack(">I4s", raw[pos:pos + 8])
        body = raw[pos + 8:pos + 8 + length]
        pos += 12 + length
        if kind == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif kind == b"IDAT":
            idat.append(body)
        elif kind == b"IEND":
            break
    if hdr is None:
        raise ValueError(f"{path}: no IHDR")
    w, h, depth, ctype, _, _, interlace = hdr
    if interlace:
        raise ValueError(f"{path}: interlaced PNG not supported")
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}.get(ctype)
    if ch is None or depth not in (1, 2, 4, 8, 16) or (depth < 8 and ctype not in (0, 3)):
        raise ValueError(f"{path}: unsupported PNG colour type {ctype} / depth {depth}")
    data = zlib.decompress(b"".join(idat))
    bpp = max(1, ch * depth // 8)                 # bytes per complete pixel, for the filters
    stride = (w * ch * depth + 7) // 8
    rows = np.zeros((h, stride), dtype=np.uint8)
    prev = np.zeros(stride, dtype=np.int64)
    for y in range(h):
        ft = data[y * (stride + 1)]
        line = np.frombuffer(data, dtype=np.uint8, count=stride, offset=y * (stride + 1) + 1).astype(np.int64)
        if ft == 0:
            cur = line
        elif ft == 2:
            cur = (line + prev) & 255
        else:                                     # Sub / Average / Paeth depend on the pixel to the left
            cur = np.zeros(stride, dtype=np.int64)
            for x in range(stride):
                a = cur[x - bpp] if x >= bpp else 0
                b = prev[x]
                if ft == 1:
                    pred = a
                elif ft == 3:
                    pred = (a + b) >> 1
                elif ft == 4:
                    c = prev[x - bpp] if x >= bpp else 0
                    pa, pb, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                else:
                    raise ValueError(f"{path}: bad PNG filter {ft}")
                cur[x] = (line[x] + pred) & 255
        rows[y] = cur
        prev = cur
    if depth == 16:
        img = _narrow(rows.reshape(h, w * ch, 2).astype(np.int64) @ np.array([256, 1]), 65535)
    elif depth == 8:
        img = rows
    else:                                         # packed 1/2/4-bit samples, most significant first
        bits = np.unpackbits(rows, axis=1)[:, :w * depth].reshape(h, w, depth)
        vals = bits.astype(np.int64) @ (1 << np.arange(depth - 1, -1, -1))
        img = vals.astype(np.uint8) if ctype == 3 else _narrow(vals, (1 << depth) - 1)
    img = np.ascontiguousarray(img.reshape((h, w, ch) if ch > 1 else (h, w)).astype(np.uint8))
    return img, ctype in (4, 6)


def read_image(path: str):
    """Any map image the loader understands -> ((H, W) or (H, W, C) uint8 rows top to bottom, has_alpha)."""
    with open(path, "rb") as f:
        magic = f.read(8)
    if magic[:8] == b"\x89PNG\r\n\x1a\n":
        return read_png(path)
    if magic[:2] in (b"P2", b"P5", b"P3", b"P6"):
        return read_pnm(path), False
    raise ValueError(f"{path}: unsupported image format (PGM, PPM and PNG are read)")


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
