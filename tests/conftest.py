import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        from pyracecarsimulator_b200 import _native
        return _native.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a device must fail loudly rather than pass by skipping
    # (the driver records which .so files were loaded); CPU-only runs use -m "not gpu".
    pass


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def colombia():
    """The reference's only shipped map (maps/colombia) from the committed fixture."""
    z = np.load(os.path.join(GOLDEN, "colombia_map.npz"))
    return dict(img=z["img"], resolution=float(z["resolution"]), origin=tuple(float(v) for v in z["origin"]),
                negate=int(z["negate"]), occupied_thresh=float(z["occupied_thresh"]),
                free_thresh=float(z["free_thresh"]))


@pytest.fixture(scope="session")
def colombia_scan():
    return dict(np.load(os.path.join(GOLDEN, "colombia_scan.npz")))


@pytest.fixture(scope="session")
def car_golden():
    return dict(np.load(os.path.join(GOLDEN, "car_golden.npz")))


@pytest.fixture(scope="session")
def colombia_dist(orc, colombia):
    grid = orc.mapserver_occupancy(colombia["img"], colombia["negate"], colombia["occupied_thresh"],
                                   colombia["free_thresh"])
    occ = orc.omap_from_grid(grid, True)
    return occ, orc.edt_float(occ)


def range_tolerance(want, resolution):
    """north_star: ranges within max(1e-4 relative, 0.5 map cell)."""
    return np.maximum(1e-4 * np.abs(want), 0.5 * resolution)
