"""ctypes binding of ``librangelib_b200.so`` (include/rangelib_b200.h).

There is no CPU implementation behind this module: if the CUDA library has not been built
(``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C pyracecarsimulator_b200/csrc``)
loading fails loudly, and every call that needs a device returns ``RL_ERR_NO_DEVICE`` on a
box without one.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librangelib_b200.so")

RL_OK = 0
RL_ERR_BAD_ARG = -1
RL_ERR_CUDA = -2
RL_ERR_NO_DEVICE = -3
RL_ERR_OOM = -4
RL_FLAG_DEFAULT = 0
RL_FLAG_NO_L2_WINDOW = 1
RL_FLAG_NO_PADDED_FIELD = 2
RL_FLAG_NO_POSE_SORT = 4
RL_PIPELINE_OFF, RL_PIPELINE_STREAMS, RL_PIPELINE_PDL = 0, 1, 2
RL_DIST2_INF = 0x3FFFFFFF


class NativeLibraryMissing(ImportError):
    pass


_lib = None

_vp = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64
_f = C.c_float
_d = C.c_double

# name -> (restype, argtypes); mirrors include/rangelib_b200.h one to one
SIGNATURES = {
    "rl_abi_version": (_i32, []),
    "rl_last_error": (C.c_char_p, []),
    "rl_device_count": (_i32, [C.POINTER(_i32)]),
    "rl_map_from_image": (_i32, [_vp, _i32, _i32, _i32, _d, _d, _i32, _i32, _d, _d, _d, _d, _i32, C.POINTER(_vp)]),
    "rl_map_from_image_channels": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _d, _d, _i32, _i32, _d, _d, _d, _d, _i32, C.POINTER(_vp)]),
    "rl_map_from_occupancy": (_i32, [_vp, _i32, _i32, _i32, _d, _d, _d, _d, _i32, C.POINTER(_vp)]),
    "rl_map_from_cells": (_i32, [_vp, _i32, _i32, _d, _d, _d, _d, _i32, C.POINTER(_vp)]),
    "rl_map_shape": (_i32, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32)]),
    "rl_map_get_occupancy": (_i32, [_vp, _vp]),
    "rl_map_get_dist2": (_i32, [_vp, _vp]),
    "rl_map_get_dist": (_i32, [_vp, _vp]),
    "rl_map_dist_device": (_i32, [_vp, C.POINTER(_vp)]),
    "rl_map_ingest_ms": (_i32, [_vp, C.POINTER(_f)]),
    "rl_map_destroy": (_i32, [_vp]),
    "rl_marcher_create": (_i32, [_vp, _f, C.c_uint32, C.POINTER(_vp)]),
    "rl_marcher_destroy": (_i32, [_vp]),
    "rl_calc_range_many": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "rl_calc_range_many_host": (_i32, [_vp, _vp, _vp, _i64]),
    "rl_calc_range_fan": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _f, _vp]),
    "rl_calc_range_fan_host": (_i32, [_vp, _vp, _i64, _vp, _i64, _i32, _f]),
    "rl_calc_range_repeat_angles": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "rl_calc_range_repeat_angles_host": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32]),
    "rl_peer_alloc": (_i32, [_i32, _i64, C.POINTER(_vp), _vp]),
    "rl_peer_open": (_i32, [_i32, _vp, C.POINTER(_vp)]),
    "rl_peer_close": (_i32, [_i32, _vp]),
    "rl_peer_free": (_i32, [_i32, _vp]),
    "rl_calc_range_fan_allgather": (_i32, [_vp, _vp, _i64, _vp, _i32, _i32, _i64, _i64, _i32, _f, C.c_uint32, _vp]),
    "rl_calc_range_repeat_angles_allgather": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i64, _i64, _i32, C.c_uint32, _vp]),
    "rl_allgather_ranges": (_i32, [_i32, _vp, _vp, _i32, _i32, _i64, _i64, C.c_uint32, _vp]),
    "rl_marcher_set_pipelined": (_i32, [_vp, _i32]),
    "rl_marcher_join": (_i32, [_vp, _vp]),
    "rl_marcher_count_steps": (_i32, [_vp, _i32]),
    "rl_marcher_last_steps": (_i32, [_vp, C.POINTER(C.c_uint64)]),
    "rl_car_create": (_i32, [_vp, _i32, C.POINTER(_vp)]),
    "rl_car_destroy": (_i32, [_vp]),
    "rl_car_set_edge_distances": (_i32, [_vp, _i32, _d, _d, _d]),
    "rl_car_get_edge_distances": (_i32, [_vp, _vp, _i32]),
    "rl_car_step": (_i32, [_vp, _vp, _vp, _vp, _i64, _d, _vp]),
    "rl_is_crashed": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "rl_scan_crash": (_i32, [_vp, _vp, _vp, _i64, _i32, _f, _vp, _vp, _vp]),
    "rl_rollout": (_i32, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _d, _i32, _d, _f, _vp, _vp, _vp, _vp, _vp]),
    "rl_rollout_actions": (_i32, [_vp, _i64, _i32, C.c_uint64, C.c_uint32, _i64, _d, _d, _d, _d, _i32, _vp]),
    "rl_rollout_value": (_i32, [_vp, _vp, _i64, _vp, _i32, _vp]),
    "rl_follow_gap": (_i32, [_vp, _i64, _i32, _f, _f, _f, _vp, _vp]),
    "rl_host_register": (_i32, [_i32, _vp, _i64, C.POINTER(_i32)]),
    "rl_host_unregister": (_i32, [_i32, _vp]),
    "rl_probe_sincosf": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "rl_l2_reset_persisting": (_i32, [_i32]),
    "rl_gather_bandwidth": (_i32, [_i32, _i64, _i32, _i32, C.POINTER(_f)]),
    "rl_l2_sector_bandwidth": (_i32, [_i32, _i64, _i32, _i32, C.POINTER(_f)]),
}


def lib():
    """Load the CUDA library once; raise NativeLibraryMissing if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: build the CUDA extension first "
            "(make -C pyracecarsimulator_b200/csrc). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    if L.rl_abi_version() != 1:
        raise NativeLibraryMissing(f"{LIB_PATH}: ABI version {L.rl_abi_version()} != 1")
    _lib = L
    return L


def last_error() -> str:
    return lib().rl_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    """Map a negative rl_status to the Python exception the shim raises (never sys.exit)."""
    if rc == RL_OK:
        return
    msg = f"{what}: {last_error()}" if what else last_error()
    if rc == RL_ERR_BAD_ARG:
        raise ValueError(msg)
    if rc == RL_ERR_OOM:
        raise MemoryError(msg)
    raise RuntimeError(msg)


def gather_bandwidth(device: int, buffer_bytes: int, rounds: int = 64, iters: int = 10) -> float:
    """GB/s (4 B per gather) of random gathers from an L2-resident buffer of this size."""
    v = _f()
    check(lib().rl_gather_bandwidth(device, buffer_bytes, rounds, iters, C.byref(v)), "gather_bandwidth")
    return float(v.value)


def l2_sector_bandwidth(device: int, buffer_bytes: int, rounds: int = 64, iters: int = 10) -> float:
    """10^9 sectors/s of random full-sector (32 B) reads from an L2-resident buffer of this size."""
    v = _f()
    check(lib().rl_l2_sector_bandwidth(device, buffer_bytes, rounds, iters, C.byref(v)), "l2_sector_bandwidth")
    return float(v.value)


def device_count() -> int:
    n = _i32(0)
    rc = lib().rl_device_count(C.byref(n))
    return int(n.value) if rc == RL_OK else 0
