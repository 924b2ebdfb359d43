"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; ``pyracecarsimulator_b200`` never
does.  ``liboracle.so`` holds the restatements (``rangelib_oracle.c``: the external,
un-vendored range_libc scan path, PARITY UNPINNED; ``car_oracle.c``: the vendored
vehicle model) and ``_ref/libracecar_ref.so`` is the unmodified reference ``Car``
(``/root/reference/racecar/src/racecar.cpp``) behind ``ref_shim.cpp``.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

D2_INF = 0x3FFFFFFF

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile liboracle.so (and _ref/ when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("rangelib_oracle.c", "car_oracle.c", "trig_twin.c", "followgap_oracle.c",
                                             "philox_oracle.c")]
    stale = force or not os.path.exists(so) or any(
        os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    ref_so = os.path.join(_HERE, "_ref", "libracecar_ref.so")
    fg_so = os.path.join(_HERE, "_ref", "libfollowgap_ref.so")
    want_ref = os.path.exists("/root/reference/racecar/src/racecar.cpp") and not (
        os.path.exists(ref_so) and os.path.exists(fg_so))
    if stale or want_ref:
        env = dict(os.environ)
        env.pop("CC", None)
        env.pop("CXX", None)
        subprocess.run(["make", "-C", _HERE, "CC=gcc", "CXX=g++"] + (["-B"] if force else []),
                       check=True, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return so


class CarParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "wb", "fc", "h_cg", "l_f", "l_r", "cs_f", "cs_r", "mass", "i_z", "crash_thresh",
        "width", "length", "max_steer_vel", "max_steer_ang", "max_speed", "max_accel",
        "max_decel")]

    def as_array(self):
        return np.array([getattr(self, n) for n, _ in self._fields_], dtype=np.float64)


# params.yaml:1-22,31,36 -- the reference's car (ctor order racecar/src/racecar.cpp:10-13)
DEFAULT_CAR = dict(wb=0.3302, fc=1.0, h_cg=0.08255, l_f=0.15875, l_r=0.17145, cs_f=2.3, cs_r=2.3,
                   mass=3.17, i_z=0.0398378, crash_thresh=0.001, width=0.2032, length=0.4064,
                   max_steer_vel=5.0, max_steer_ang=0.4189, max_speed=7.0, max_accel=3.0,
                   max_decel=20.0)


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    L.orc_mapserver_occupancy.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, i8p]
    L.orc_mapserver_occupancy_mode.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, i8p]
    L.orc_mapserver_occupancy_channels.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                                   C.c_double, C.c_int, i8p]
    L.orc_omap_from_grid.argtypes = [i8p, C.c_int64, C.c_int, u8p]
    L.orc_edt_float.argtypes = [u8p, C.c_int, C.c_int, f32p, C.c_void_p]
    L.orc_edt_exact.argtypes = [u8p, C.c_int, C.c_int, i32p]
    L.orc_sqrt_dist2.argtypes = [i32p, C.c_int64, f32p]
    L.orc_marcher_create.restype = C.c_void_p
    L.orc_marcher_create.argtypes = [f32p, C.c_int, C.c_int, C.c_float, C.c_double, C.c_double,
                                     C.c_double, C.c_double]
    L.orc_marcher_destroy.argtypes = [C.c_void_p]
    L.orc_calc_range.restype = C.c_float
    L.orc_calc_range.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
    L.orc_calc_range_many.argtypes = [C.c_void_p, f32p, f32p, C.c_int64, C.c_void_p, C.c_int]
    L.orc_variant_fan.argtypes = [C.c_void_p, f32p, f32p, C.c_int64, C.c_int, C.c_float, C.c_uint, C.c_int]
    L.orc_calc_range_fan.argtypes = [C.c_void_p, f32p, f32p, C.c_int64, C.c_int, C.c_float,
                                     C.c_int64, C.c_void_p, C.c_int]
    L.orc_calc_range_repeat_angles.argtypes = [C.c_void_p, f32p, f32p, f32p, C.c_int64, C.c_int,
                                               C.c_void_p, C.c_int]
    L.orc_max_threads.restype = C.c_int
    L.orc_twin_sincosf.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.orc_twin_mismatches.restype = C.c_uint64
    L.orc_twin_mismatches.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
    L.orc_followgap_eval.restype = C.c_float
    L.orc_followgap_eval.argtypes = [f32p, C.c_int, C.c_float, C.c_float, C.c_float]
    L.orc_car_step.argtypes = [C.POINTER(CarParams), f64p, C.c_double, C.c_double, C.c_double]
    L.orc_car_scan_pose.argtypes = [f64p, C.c_double, f64p]
    L.orc_car_edge_distances.argtypes = [C.POINTER(CarParams), C.c_int, C.c_double, C.c_double,
                                         C.c_double, f64p]
    L.orc_philox4x32_10.argtypes = [u32p, u32p, u32p]
    L.orc_rollout_actions.argtypes = [f64p, C.c_int64, C.c_int32, C.c_uint64, C.c_uint32, C.c_int64,
                                      C.c_double, C.c_double, C.c_double, C.c_double]
    L.orc_car_is_crashed.restype = C.c_int
    L.orc_car_is_crashed.argtypes = [f32p, f64p, C.c_int, C.c_int, C.c_double]
    _LIB = L
    return L


# --------------------------------------------------------------------------- map ingest
MODES = {"trinary": 0, "scale": 1, "raw": 2}


def mapserver_occupancy(img, negate=0, occupied_thresh=0.65, free_thresh=0.196, mode="trinary"):
    """(H, W) uint8 image rows top-to-bottom -> (H, W) int8 OccupancyGrid, row 0 = bottom."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    out = np.empty((h, w), dtype=np.int8)
    if mode == "trinary":
        lib().orc_mapserver_occupancy(img, w, h, int(negate), float(occupied_thresh), float(free_thresh), out)
    else:
        lib().orc_mapserver_occupancy_mode(img, w, h, int(negate), float(occupied_thresh), float(free_thresh),
                                           MODES[mode], out)
    return out


def mapserver_occupancy_channels(img, has_alpha, negate=0, occupied_thresh=0.65, free_thresh=0.196, mode="trinary"):
    """(H, W, C) uint8 colour / alpha image -> (H, W) int8 OccupancyGrid (map_server's channel averaging)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w, c = img.shape
    out = np.empty((h, w), dtype=np.int8)
    lib().orc_mapserver_occupancy_channels(img.ctypes.data, w, h, c, int(bool(has_alpha)), int(negate),
                                           float(occupied_thresh), float(free_thresh), MODES[mode], out)
    return out


def omap_from_grid(grid, binarise=True):
    grid = np.ascontiguousarray(grid, dtype=np.int8)
    out = np.empty(grid.shape, dtype=np.uint8)
    lib().orc_omap_from_grid(grid, grid.size, int(binarise), out)
    return out


def edt_float(occupied, want_dist2=False):
    occupied = np.ascontiguousarray(occupied, dtype=np.uint8)
    rows, cols = occupied.shape
    dist = np.empty((rows, cols), dtype=np.float32)
    d2 = np.empty((rows, cols), dtype=np.float32) if want_dist2 else None
    lib().orc_edt_float(occupied, rows, cols, dist, d2.ctypes.data if want_dist2 else None)
    return (dist, d2) if want_dist2 else dist


def edt_exact(occupied):
    occupied = np.ascontiguousarray(occupied, dtype=np.uint8)
    rows, cols = occupied.shape
    d2 = np.empty((rows, cols), dtype=np.int32)
    lib().orc_edt_exact(occupied, rows, cols, d2)
    return d2


def sqrt_dist2(d2):
    d2 = np.ascontiguousarray(d2, dtype=np.int32)
    out = np.empty(d2.shape, dtype=np.float32)
    lib().orc_sqrt_dist2(d2, d2.size, out)
    return out


# --------------------------------------------------------------------------- marcher
class Marcher:
    """Restated range_libc ``RayMarching`` over a given fp32 distance field."""

    def __init__(self, dist, max_range_px, resolution, origin=(0.0, 0.0, 0.0)):
        self.dist = np.ascontiguousarray(dist, dtype=np.float32)  # keep alive: borrowed by C
        rows, cols = self.dist.shape
        self.rows, self.cols = rows, cols
        self.resolution = float(resolution)
        self._h = lib().orc_marcher_create(self.dist, rows, cols, float(max_range_px),
                                           float(resolution), float(origin[0]), float(origin[1]),
                                           float(origin[2]))

    def __del__(self):
        if getattr(self, "_h", None) and _LIB is not None:
            _LIB.orc_marcher_destroy(self._h)
            self._h = None

    @staticmethod
    def _steps(n, want):
        s = np.zeros(n, dtype=np.int32) if want else None
        return s, (s.ctypes.data if want else None)

    def calc_range(self, x, y, theta):
        return float(lib().orc_calc_range(self._h, x, y, theta))

    def calc_range_many(self, ins, outs=None, steps=False, threads=1):
        ins = np.ascontiguousarray(ins, dtype=np.float32)
        n = ins.shape[0]
        outs = np.empty(n, dtype=np.float32) if outs is None else outs
        s, sp = self._steps(n, steps)
        lib().orc_calc_range_many(self._h, ins, outs, n, sp, threads)
        return (outs, s) if steps else outs

    def calc_range_fan(self, poses, num_rays, fov, outs=None, steps=False, threads=1,
                       pose_stride_rows=1):
        poses = np.ascontiguousarray(poses, dtype=np.float32)
        b = poses.shape[0] // pose_stride_rows if pose_stride_rows > 1 else poses.shape[0]
        outs = np.empty(b * num_rays, dtype=np.float32) if outs is None else outs
        s, sp = self._steps(b * num_rays, steps)
        lib().orc_calc_range_fan(self._h, poses, outs, b, num_rays, fov, pose_stride_rows, sp,
                                 threads)
        return (outs, s) if steps else outs

    def variant_fan(self, poses, num_rays, fov, mask, threads=1):
        """NOT the oracle: the fan with a subset of the restatement's fp32 choices flipped
        (rangelib_oracle.c, section A.7; ``mask`` bits 1 .. 32), for the sensitivity study."""
        poses = np.ascontiguousarray(poses, dtype=np.float32)
        outs = np.empty(poses.shape[0] * num_rays, dtype=np.float32)
        lib().orc_variant_fan(self._h, poses, outs, poses.shape[0], num_rays, fov, int(mask), threads)
        return outs

    def calc_range_repeat_angles(self, ins, angles, outs=None, steps=False, threads=1):
        ins = np.ascontiguousarray(ins, dtype=np.float32)
        angles = np.ascontiguousarray(angles, dtype=np.float32)
        n, a = ins.shape[0], angles.shape[0]
        outs = np.empty(n * a, dtype=np.float32) if outs is None else outs
        s, sp = self._steps(n * a, steps)
        lib().orc_calc_range_repeat_angles(self._h, ins, angles, outs, n, a, sp, threads)
        return (outs, s) if steps else outs


def twin_sincosf(y):
    """(sin, cos) from the C twin of the device's glibc-compatible sincosf."""
    s, c = C.c_float(), C.c_float()
    lib().orc_twin_sincosf(float(y), C.byref(s), C.byref(c))
    return s.value, c.value


def twin_mismatches(start, stride, count):
    return int(lib().orc_twin_mismatches(start, stride, count))


def max_threads():
    return int(lib().orc_max_threads())


# --------------------------------------------------------------------------- car
def car_params(**kw):
    d = dict(DEFAULT_CAR)
    d.update(kw)
    return CarParams(**d)


def car_step(params, state, speed, steer, dt=0.01):
    lib().orc_car_step(C.byref(params), state, speed, steer, dt)
    return state


def car_rollout_poses(params, states, actions, steps, every=10, dt=0.01):
    """Vehicle half of MCTS.rollout for many cars on the CPU: returns the (steps, n, 3) float32 base-link poses
    (step-major, like the device rollout); ``states`` (n, 11) float64 is updated in place."""
    states = np.ascontiguousarray(states, dtype=np.float64)
    actions = np.ascontiguousarray(actions, dtype=np.float64)
    n = states.shape[0]
    poses = np.empty((steps, n, 3), dtype=np.float32)
    f = lib().orc_car_rollout_poses
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_void_p]
    f(C.byref(params), states.ctypes.data, actions.ctypes.data, n, int(steps), int(every), float(dt), poses.ctypes.data)
    return poses, states


def car_scan_pose(state, scan_dist_to_base):
    pose = np.empty(3, dtype=np.float64)
    lib().orc_car_scan_pose(state, scan_dist_to_base, pose)
    return pose


def car_edge_distances(params, num_rays, min_ang, inc, scan_dist_to_base):
    edge = np.empty(num_rays, dtype=np.float64)
    lib().orc_car_edge_distances(C.byref(params), num_rays, min_ang, inc, scan_dist_to_base, edge)
    return edge


def car_is_crashed(rays, edge, num_rays, poses, crash_thresh):
    rays = np.ascontiguousarray(rays, dtype=np.float32)
    return int(lib().orc_car_is_crashed(rays, edge, num_rays, poses, crash_thresh))


def philox4x32_10(ctr, key):
    """One Philox4x32-10 block: 4 counter words, 2 key words -> 4 output words (uint32)."""
    ctr = np.ascontiguousarray(ctr, dtype=np.uint32)
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.empty(4, dtype=np.uint32)
    lib().orc_philox4x32_10(ctr, key, out)
    return out


def rollout_actions(n_cars, n_actions, seed=42, stream_id=0, car_offset=0, speed_range=(0.0, 7.0),
                    steer_range=(-0.4189, 0.4189)):
    """(n_cars, n_actions, 2) float64 (speed, steer): the schedule rl_rollout_actions draws on the device."""
    out = np.empty((n_cars, n_actions, 2), dtype=np.float64)
    lib().orc_rollout_actions(out.reshape(-1), n_cars, n_actions, seed, stream_id, car_offset,
                              speed_range[0], speed_range[1], steer_range[0], steer_range[1])
    return out


def followgap_eval(lidar, max_distance=15.0, max_angle=0.4189, angle_inc=0.004):
    """FollowGap(ws, md, ma, inc).eval(lidar, len(lidar)) restated (window size is unused upstream)."""
    lidar = np.ascontiguousarray(lidar, dtype=np.float32)
    return float(lib().orc_followgap_eval(lidar, lidar.size, max_distance, max_angle, angle_inc))


_FG = None


def ref_followgap_eval(lidar, ws=10, max_distance=15.0, max_angle=0.4189, angle_inc=0.004):
    """The unmodified reference FollowGap (followgap/followgap.hpp) via ref_followgap_shim.cpp."""
    global _FG
    if _FG is None:
        build()
        _FG = C.CDLL(os.path.join(_HERE, "_ref", "libfollowgap_ref.so"))
        _FG.ref_followgap_eval.restype = C.c_float
        _FG.ref_followgap_eval.argtypes = [f32p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
    lidar = np.ascontiguousarray(lidar, dtype=np.float32).copy()
    return float(_FG.ref_followgap_eval(lidar, lidar.size, ws, max_distance, max_angle, angle_inc))


def followgap_ref_available():
    build()
    return os.path.exists(os.path.join(_HERE, "_ref", "libfollowgap_ref.so"))


# --------------------------------------------------------------------------- the real reference
def ref_available():
    build()
    return os.path.exists(os.path.join(_HERE, "_ref", "libracecar_ref.so"))


class RefCar:
    """The unmodified reference ``Car`` (racecar/src/racecar.cpp) via ref_shim.cpp."""

    def __init__(self, params):
        global _REF
        if _REF is None:
            build()
            R = C.CDLL(os.path.join(_HERE, "_ref", "libracecar_ref.so"))
            R.ref_car_create.restype = C.c_void_p
            R.ref_car_create.argtypes = [f64p]
            R.ref_car_destroy.argtypes = [C.c_void_p]
            R.ref_car_control.argtypes = [C.c_void_p, C.c_double, C.c_double]
            R.ref_car_update.argtypes = [C.c_void_p, C.c_double]
            R.ref_car_get_state.argtypes = [C.c_void_p, f64p]
            R.ref_car_set_state.argtypes = [C.c_void_p, f64p]
            R.ref_car_scan_pose.argtypes = [C.c_void_p, C.c_double, f64p]
            R.ref_car_set_edges.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]
            R.ref_car_is_crashed.restype = C.c_int
            R.ref_car_is_crashed.argtypes = [C.c_void_p, f32p, C.c_int, C.c_int]
            _REF = R
        self._h = _REF.ref_car_create(params.as_array())

    def __del__(self):
        if getattr(self, "_h", None) and _REF is not None:
            _REF.ref_car_destroy(self._h)
            self._h = None

    def control(self, speed, steer):
        _REF.ref_car_control(self._h, speed, steer)

    def update(self, dt=0.01):
        _REF.ref_car_update(self._h, dt)

    def get_state(self):
        s = np.zeros(11, dtype=np.float64)
        _REF.ref_car_get_state(self._h, s)
        return s

    def set_state(self, s):
        _REF.ref_car_set_state(self._h, np.ascontiguousarray(s, dtype=np.float64))

    def scan_pose(self, d):
        p = np.zeros(3, dtype=np.float64)
        _REF.ref_car_scan_pose(self._h, d, p)
        return p

    def set_edges(self, num_rays, min_ang, inc, d):
        _REF.ref_car_set_edges(self._h, num_rays, min_ang, inc, d)

    def is_crashed(self, rays, num_rays, poses):
        rays = np.ascontiguousarray(rays, dtype=np.float32)
        return int(_REF.ref_car_is_crashed(self._h, rays, num_rays, poses))
