"""Drop-in for the names the reference imports from the external ``range_libc`` extension:
``PyOMap`` (scripts/ros_interface.py:210, scripts/mcts_driver.py:278), ``PyRayMarching`` and
``PyRayMarchingGPU`` (scripts/scan_simulator.py:72-76) with ``calc_range_many`` in both the
upstream 2-arg form (scripts/two_player/scan.py:69-70) and the fork's 4-arg form
(scripts/scan_simulator.py:103-106, :130-133), ``calc_range_repeat_angles`` and ``calc_range``.

Everything runs on the GPU through the C ABI (``_native``): both marcher classes are the same
CUDA implementation (north_star: no CPU fallback).  Arrays follow upstream's Cython
signatures -- caller-allocated, C-contiguous ``float32``, written in place, ``None`` returned --
and may be numpy arrays (host pointers, staged through pinned memory, synchronous) or torch
CUDA tensors (device pointers, enqueued on the current torch stream, no host copy).
"""
from __future__ import annotations

import atexit
import collections
import ctypes as C
import os
import threading

import numpy as np

from . import _native
from .maps import MapYaml, load_map_yaml, quaternion_to_yaw, read_image


MAP_MODES = {"trinary": 0, "scale": 1, "raw": 2}


def _is_torch(a) -> bool:
    return type(a).__module__.split(".")[0] == "torch"


def _current_stream_ptr(device_index: int) -> int:
    import torch
    return int(torch.cuda.current_stream(device_index).cuda_stream)


class _HostRegistry:
    """Page-locks the numpy buffers a caller passes again and again.

    The reference allocates its scan buffers once and reuses them for every call
    (scripts/scan_simulator.py:32-40, scripts/two_player/scan.py:51-53).  A pageable ``outs`` costs a
    staging copy of every range on every call; a page-locked one is written in place by the kernel.
    So the SECOND time the same (address, size) shows up it is registered with ``rl_host_register``
    and stays registered; the registry keeps a reference to the array, so its memory cannot be freed
    while it is page-locked.  Only arrays whose memory numpy itself owns are taken, at most
    ``MAX_ENTRIES`` of them and ``MAX_BYTES`` in total, nothing is ever evicted behind a caller's back;
    :func:`release_host_buffers` (or interpreter exit) unregisters everything.  ``RL_HOST_REGISTER=0``
    disables it.
    """
    MIN_BYTES = 1 << 20
    MAX_ENTRIES = 8
    MAX_BYTES = 1 << 30

    def __init__(self):
        self._lock = threading.Lock()
        self._seen = collections.OrderedDict()    # (ptr, nbytes) -> sightings
        self._reg = collections.OrderedDict()     # (ptr, nbytes) -> (array, device)
        self._skip = set()
        self.enabled = os.environ.get("RL_HOST_REGISTER", "1") != "0"

    @staticmethod
    def _numpy_owns(a) -> bool:
        while isinstance(a, np.ndarray) and a.base is not None:
            a = a.base
        return isinstance(a, np.ndarray) and bool(a.flags["OWNDATA"])

    def note(self, a: np.ndarray, device: int) -> None:
        if not self.enabled or a.nbytes < self.MIN_BYTES:
            return
        key = (int(a.ctypes.data), int(a.nbytes))
        with self._lock:
            if key in self._reg or key in self._skip:
                return
            n = self._seen.pop(key, 0) + 1
            if n < 2:
                self._seen[key] = n
                while len(self._seen) > 64:
                    self._seen.popitem(last=False)
                return
            total = sum(k[1] for k in self._reg)
            lo, hi = key[0], key[0] + key[1]
            overlaps = any(k[0] < hi and lo < k[0] + k[1] for k in self._reg)
            if (len(self._reg) >= self.MAX_ENTRIES or total + key[1] > self.MAX_BYTES or overlaps
                    or not self._numpy_owns(a)):
                self._remember_skip(key)
                return
            was = C.c_int32(0)
            rc = _native.lib().rl_host_register(int(device), key[0], key[1], C.byref(was))
            if rc != _native.RL_OK or was.value:
                self._remember_skip(key)      # already page-locked by someone else, or the driver refused
                return
            self._reg[key] = (a, int(device))

    def _remember_skip(self, key):
        if len(self._skip) > 256:
            self._skip.clear()
        self._skip.add(key)

    def registered_bytes(self) -> int:
        with self._lock:
            return sum(k[1] for k in self._reg)

    def release(self) -> None:
        with self._lock:
            if _native._lib is not None:
                for (ptr, _), (_, device) in self._reg.items():
                    _native._lib.rl_host_unregister(device, ptr)
            self._reg.clear()
            self._seen.clear()
            self._skip.clear()


_HOST_REGISTRY = _HostRegistry()
atexit.register(_HOST_REGISTRY.release)


def release_host_buffers() -> None:
    """Unregister every numpy buffer the library has page-locked (see :class:`_HostRegistry`).  Call
    it only when no scan call is in flight on another thread."""
    _HOST_REGISTRY.release()


class _Buf:
    """Pointer + placement of a caller-owned float32 array."""
    __slots__ = ("ptr", "on_device", "shape", "keep")

    def __init__(self, a, name: str, ndim: int, device: int):
        if _is_torch(a):
            import torch
            if a.dtype != torch.float32 or not a.is_contiguous() or a.dim() != ndim:
                raise ValueError(f"{name}: expected a C-contiguous float32 tensor with ndim={ndim}")
            if a.is_cuda:
                if a.device.index != device:
                    raise ValueError(f"{name}: tensor is on cuda:{a.device.index}, map is on cuda:{device}")
                self.on_device = True
            else:
                self.on_device = False
            self.ptr = int(a.data_ptr())
            self.shape = tuple(a.shape)
        elif isinstance(a, np.ndarray):
            if a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"] or a.ndim != ndim:
                raise ValueError(f"{name}: Buffer dtype mismatch / not C-contiguous: expected float32, ndim={ndim}")
            self.on_device = False
            self.ptr = int(a.ctypes.data)
            self.shape = a.shape
            _HOST_REGISTRY.note(a, device)
        else:
            raise ValueError(f"{name}: expected a numpy array or torch tensor, got {type(a).__name__}")
        self.keep = a


def _same_side(*bufs) -> bool:
    side = bufs[0].on_device
    if any(b.on_device != side for b in bufs):
        raise ValueError("inputs and outputs must all be host arrays or all be CUDA tensors")
    return side


class PyOMap:
    """Occupancy map + its distance field, resident on one GPU.

    ``PyOMap(map_msg)`` with an OccupancyGrid-shaped object follows upstream: cell occupied iff
    ``data > 10``; ``world_scale = resolution``, ``world_angle = -yaw``, origin from
    ``info.origin.position``.  Also accepted: a 2-D boolean numpy array (resolution 1, origin 0),
    a ``map.yaml`` path or :class:`MapYaml` (map_server thresholds + y-flip + the reference's
    binarisation, all on the GPU).
    """

    def __init__(self, arg, device: int | None = None, binarise: bool = True):
        L = _native.lib()
        if device is None:
            device = 0
            try:
                import torch
                if torch.cuda.is_available():
                    device = torch.cuda.current_device()
            except ImportError:
                pass
        self.device = int(device)
        h = C.c_void_p()
        if isinstance(arg, (str, MapYaml)):
            y = load_map_yaml(arg) if isinstance(arg, str) else arg
            img, has_alpha = read_image(y.image)
            self._meta = (y.resolution, y.origin[0], y.origin[1], y.origin[2])
            if img.ndim == 2:
                rc = L.rl_map_from_image(img.ctypes.data, img.shape[1], img.shape[0], y.negate,
                                         y.occupied_thresh, y.free_thresh, MAP_MODES[y.mode], int(bool(binarise)),
                                         y.resolution, y.origin[0], y.origin[1], y.origin[2],
                                         self.device, C.byref(h))
            else:   # colour / alpha: map_server averages the channels of every pixel first
                rc = L.rl_map_from_image_channels(img.ctypes.data, img.shape[1], img.shape[0], img.shape[2],
                                                  int(has_alpha), y.negate, y.occupied_thresh, y.free_thresh,
                                                  MAP_MODES[y.mode], int(bool(binarise)), y.resolution,
                                                  y.origin[0], y.origin[1], y.origin[2], self.device, C.byref(h))
        elif isinstance(arg, np.ndarray):
            if arg.ndim != 2:
                raise ValueError("PyOMap: numpy occupancy grid must be 2-D")
            cells = np.ascontiguousarray(arg != 0, dtype=np.uint8)
            self._meta = (1.0, 0.0, 0.0, 0.0)
            rc = L.rl_map_from_cells(cells.ctypes.data, cells.shape[1], cells.shape[0], 1.0, 0.0,
                                     0.0, 0.0, self.device, C.byref(h))
        elif hasattr(arg, "info") and hasattr(arg, "data"):
            info = arg.info
            data = np.asarray(arg.data)
            if data.size != info.width * info.height:
                raise ValueError("PyOMap: len(data) != info.width * info.height")
            if data.dtype != np.int8:  # e.g. the reference's tuple of 0/255 python ints
                data = np.clip(data, -128, 127).astype(np.int8)  # keeps `> 10` unchanged
            data = np.ascontiguousarray(data)
            yaw = quaternion_to_yaw(info.origin.orientation)
            self._meta = (float(info.resolution), float(info.origin.position.x),
                          float(info.origin.position.y), yaw)
            rc = L.rl_map_from_occupancy(data.ctypes.data, int(info.width), int(info.height), 0,
                                         *self._meta, self.device, C.byref(h))
        else:
            raise ValueError(f"PyOMap: unsupported argument {type(arg).__name__}")
        _native.check(rc, "PyOMap")
        self._h = h
        w, hh = C.c_int32(), C.c_int32()
        _native.check(L.rl_map_shape(self._h, C.byref(w), C.byref(hh), None))
        # upstream naming: OMap.width = msg rows, OMap.height = msg columns
        self._rows, self._cols = int(hh.value), int(w.value)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _native._lib is not None:
            _native._lib.rl_map_destroy(h)

    # ---- upstream accessors ----
    def width(self) -> int:
        return self._rows

    def height(self) -> int:
        return self._cols

    def isOccupied(self, x: int, y: int) -> bool:
        if x < 0 or y < 0 or x >= self._rows or y >= self._cols:
            return False
        occ = getattr(self, "_occ_host", None)
        if occ is None:   # the map is immutable: one device->host copy serves every later query
            occ = self._occ_host = self.occupancy()
        return bool(occ[x, y])

    def error(self) -> bool:
        return False

    # ---- bit-exact views (host copies), shape (msg rows, msg columns) ----
    def _get(self, fn, dtype):
        out = np.empty((self._rows, self._cols), dtype=dtype)
        _native.check(fn(self._h, out.ctypes.data))
        return out

    def occupancy(self) -> np.ndarray:
        return self._get(_native.lib().rl_map_get_occupancy, np.uint8)

    def dist2(self) -> np.ndarray:
        return self._get(_native.lib().rl_map_get_dist2, np.int32)

    def dist(self) -> np.ndarray:
        return self._get(_native.lib().rl_map_get_dist, np.float32)

    @property
    def ingest_ms(self) -> float:
        ms = C.c_float()
        _native.check(_native.lib().rl_map_ingest_ms(self._h, C.byref(ms)))
        return float(ms.value)

    @property
    def resolution(self) -> float:
        return self._meta[0]

    @property
    def origin(self):
        return self._meta[1:]


class PyRayMarchingGPU:
    """Ray marcher over a :class:`PyOMap`; ``max_range`` is in pixels
    (scripts/racecar_simulator_v2.py:196)."""

    def __init__(self, omap: PyOMap, max_range: float, flags: int = _native.RL_FLAG_DEFAULT):
        if not isinstance(omap, PyOMap):
            raise ValueError("expected a PyOMap")
        self._omap = omap  # the C side retains the map as well
        self.device = omap.device
        h = C.c_void_p()
        _native.check(_native.lib().rl_marcher_create(omap._h, float(max_range), int(flags), C.byref(h)),
                      type(self).__name__)
        self._h = h
        self._tls = threading.local()   # calc_range scratch, one pair per calling thread

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _native._lib is not None:
            _native._lib.rl_marcher_destroy(h)

    # ---- upstream API ----
    def calc_range(self, x, y, heading) -> float:
        t = self._tls
        if not hasattr(t, "one_in"):
            t.one_in, t.one_out = np.zeros((1, 3), dtype=np.float32), np.zeros(1, dtype=np.float32)
        t.one_in[0] = (x, y, heading)
        self.calc_range_many(t.one_in, t.one_out)
        return float(t.one_out[0])

    def calc_range_many(self, ins, outs, fov=None, num_rays=None):
        """2-arg: ``outs[i] = range(ins[i])``.  4-arg (fork): pose ``k`` is row ``k*num_rays`` of
        ``ins``; ``outs[k*num_rays + j]`` is beam ``j`` heading ``theta - fov/2 + j*fov/num_rays``."""
        L = _native.lib()
        i, o = _Buf(ins, "ins", 2, self.device), _Buf(outs, "outs", 1, self.device)
        if i.shape[1] != 3:
            raise ValueError("ins must have shape (N, 3)")
        dev = _same_side(i, o)
        if fov is None and num_rays is None:
            n = o.shape[0]
            if i.shape[0] < n:
                raise ValueError("ins has fewer rows than outs")
            if dev:
                rc = L.rl_calc_range_many(self._h, i.ptr, o.ptr, n, _current_stream_ptr(self.device))
            else:
                rc = L.rl_calc_range_many_host(self._h, i.ptr, o.ptr, n)
        else:
            if fov is None or num_rays is None:
                raise ValueError("calc_range_many takes (ins, outs) or (ins, outs, fov, num_rays)")
            num_rays = int(num_rays)
            if num_rays <= 0:
                raise ValueError("num_rays must be positive")
            b = o.shape[0] // num_rays
            if b > 0 and i.shape[0] < (b - 1) * num_rays + 1:
                raise ValueError("ins is too short for outs.shape[0] // num_rays poses")
            if dev:
                rc = L.rl_calc_range_fan(self._h, i.ptr, num_rays, o.ptr, b, num_rays, float(fov),
                                         _current_stream_ptr(self.device))
            else:
                rc = L.rl_calc_range_fan_host(self._h, i.ptr, num_rays, o.ptr, b, num_rays, float(fov))
        _native.check(rc, "calc_range_many")

    def calc_range_repeat_angles(self, ins, angles, outs):
        """``outs[i*A + a] = range(x_i, y_i, theta_i + angles[a])``."""
        L = _native.lib()
        i, a, o = (_Buf(ins, "ins", 2, self.device), _Buf(angles, "angles", 1, self.device),
                   _Buf(outs, "outs", 1, self.device))
        if i.shape[1] != 3:
            raise ValueError("ins must have shape (N, 3)")
        n, na = i.shape[0], a.shape[0]
        if o.shape[0] < n * na:
            raise ValueError("outs must hold ins.shape[0] * angles.shape[0] ranges")
        dev = _same_side(i, a, o)
        if dev:
            rc = L.rl_calc_range_repeat_angles(self._h, i.ptr, a.ptr, o.ptr, n, na,
                                               _current_stream_ptr(self.device))
        else:
            rc = L.rl_calc_range_repeat_angles_host(self._h, i.ptr, a.ptr, o.ptr, n, na)
        _native.check(rc, "calc_range_repeat_angles")

    # ---- additive fast path: compact (B, 3) poses, no dead rows ----
    def calc_range_fan(self, poses, outs, fov, num_rays):
        L = _native.lib()
        p, o = _Buf(poses, "poses", 2, self.device), _Buf(outs, "outs", 1, self.device)
        if p.shape[1] != 3:
            raise ValueError("poses must have shape (B, 3)")
        b, num_rays = p.shape[0], int(num_rays)
        if o.shape[0] < b * num_rays:
            raise ValueError("outs must hold B * num_rays ranges")
        if _same_side(p, o):
            rc = L.rl_calc_range_fan(self._h, p.ptr, 1, o.ptr, b, num_rays, float(fov),
                                     _current_stream_ptr(self.device))
        else:
            rc = L.rl_calc_range_fan_host(self._h, p.ptr, 1, o.ptr, b, num_rays, float(fov))
        _native.check(rc, "calc_range_fan")

    # ---- pipelined launches (include/rangelib_b200.h: rl_marcher_set_pipelined) ----
    def set_pipelined(self, mode=_native.RL_PIPELINE_STREAMS):
        """Let consecutive device-tensor calls of this marcher overlap (``mode``: ``"streams"``/1, ``"pdl"``/2,
        ``False``/0).  The ranges of a call are then valid on the current stream after the NEXT call on
        this marcher or after :meth:`join`; a call must not consume what the call before it produces."""
        mode = {"off": 0, "streams": 1, "pdl": 2, False: 0, True: 1, None: 0}.get(mode, mode)
        _native.check(_native.lib().rl_marcher_set_pipelined(self._h, int(mode)), "set_pipelined")

    def join(self):
        """Make the current torch stream wait for every pipelined march issued so far."""
        _native.check(_native.lib().rl_marcher_join(self._h, _current_stream_ptr(self.device)), "join")

    # ---- roofline support: count distance-field loads ----
    def count_steps(self, enable: bool = True):
        _native.check(_native.lib().rl_marcher_count_steps(self._h, int(enable)))

    def last_steps(self) -> int:
        v = C.c_uint64()
        _native.check(_native.lib().rl_marcher_last_steps(self._h, C.byref(v)))
        return int(v.value)


class PyRayMarching(PyRayMarchingGPU):
    """Upstream's CPU class name; here the same GPU implementation (no CPU path exists)."""
