"""The drop-in boundary without a GPU: the C-ABI library loads, exports every symbol the header
declares, the ctypes table matches the header, the product never routes through the oracle, and
a box without a device gets a loud error instead of a CPU fallback."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rangelib_b200.h")


def header_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"^RL_API\s+[\w\s\*]+?\b(rl_\w+)\s*\(", text, flags=re.M)))


def test_header_declares_the_boundary():
    names = header_symbols()
    for must in ("rl_map_from_image", "rl_map_from_occupancy", "rl_marcher_create", "rl_calc_range_many",
                 "rl_calc_range_fan", "rl_calc_range_repeat_angles", "rl_calc_range_many_host",
                 "rl_calc_range_fan_host", "rl_calc_range_repeat_angles_host", "rl_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from pyracecarsimulator_b200 import _native
    assert os.path.exists(_native.LIB_PATH), "build the CUDA extension first (__graft_entry__.build())"
    L = ctypes.CDLL(_native.LIB_PATH)
    for name in header_symbols():
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert L.rl_abi_version() == 1


def test_ctypes_table_matches_header():
    from pyracecarsimulator_b200 import _native
    assert sorted(_native.SIGNATURES) == header_symbols()
    # argument counts agree with the prototypes
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, args) in _native.SIGNATURES.items():
        proto = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, text, flags=re.S).group(1).strip()
        n = 0 if proto == "void" else proto.count(",") + 1
        assert n == len(args), name


def test_header_cites_the_reference_interfaces():
    text = open(HEADER).read()
    for cite in ("scripts/scan_simulator.py:103-106", "scripts/scan_simulator.py:72-73",
                 "scripts/two_player/scan.py:69-70", "scripts/ros_interface.py:210"):
        assert cite in text


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "pyracecarsimulator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src, f
                assert "#include \"../../oracle" not in src, f


def test_missing_library_fails_loudly(monkeypatch):
    from pyracecarsimulator_b200 import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", "/nonexistent/librangelib_b200.so")
    with pytest.raises(ImportError):
        _native.lib()


def test_no_device_is_an_error_not_a_fallback():
    from pyracecarsimulator_b200 import _native, range_libc
    if _native.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        range_libc.PyOMap(np.zeros((8, 8), dtype=bool))


def test_bad_arguments_return_status_codes():
    from pyracecarsimulator_b200 import _native
    L = _native.lib()
    out = ctypes.c_void_p()
    assert L.rl_map_from_cells(None, 4, 4, 0.05, 0.0, 0.0, 0.0, 0, ctypes.byref(out)) == _native.RL_ERR_BAD_ARG
    buf = np.zeros(16, np.uint8)
    assert L.rl_map_from_cells(buf.ctypes.data, 0, 4, 0.05, 0.0, 0.0, 0.0, 0, ctypes.byref(out)) == _native.RL_ERR_BAD_ARG
    assert L.rl_map_from_cells(buf.ctypes.data, 4, 4, -1.0, 0.0, 0.0, 0.0, 0, ctypes.byref(out)) == _native.RL_ERR_BAD_ARG
    assert b"resolution" in L.rl_last_error()
    assert L.rl_marcher_create(None, 300.0, 0, ctypes.byref(out)) == _native.RL_ERR_BAD_ARG
    assert L.rl_calc_range_many(None, None, None, 4, None) == _native.RL_ERR_BAD_ARG
    with pytest.raises(ValueError):
        _native.check(_native.RL_ERR_BAD_ARG, "x")


def test_buffer_validation_follows_upstream_signatures():
    from pyracecarsimulator_b200.range_libc import _Buf
    ok = _Buf(np.zeros((4, 3), np.float32), "ins", 2, 0)
    assert not ok.on_device and ok.shape == (4, 3)
    with pytest.raises(ValueError):
        _Buf(np.zeros((4, 3), np.float64), "ins", 2, 0)          # dtype
    with pytest.raises(ValueError):
        _Buf(np.zeros((3, 4), np.float32).T, "ins", 2, 0)        # not C-contiguous
    with pytest.raises(ValueError):
        _Buf(np.zeros(12, np.float32), "ins", 2, 0)              # ndim
    with pytest.raises(ValueError):
        _Buf([[0.0, 0.0, 0.0]], "ins", 2, 0)                     # not an array
