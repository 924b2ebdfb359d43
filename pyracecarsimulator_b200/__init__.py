"""B200-native batched 2-D lidar scan path, drop-in for felrock/PyRacecarSimulator's
``scan_simulator.ScanSimulator2D`` and the ``range_libc`` names it binds
(``PyOMap`` / ``PyRayMarching`` / ``PyRayMarchingGPU``).

    from pyracecarsimulator_b200 import range_libc            # instead of `import range_libc`
    from pyracecarsimulator_b200.scan_simulator import ScanSimulator2D

All arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI declared in
``include/rangelib_b200.h`` (``librangelib_b200.so``, loaded with ctypes on first use).
There is no CPU fallback: without the built library, or without a GPU, calls raise.
"""
from . import maps  # noqa: F401  (host-side file formats only)
from . import _native  # noqa: F401
from . import range_libc  # noqa: F401
from .range_libc import PyOMap, PyRayMarching, PyRayMarchingGPU  # noqa: F401
from .scan_simulator import ScanSimulator2D  # noqa: F401

__all__ = ["maps", "range_libc", "PyOMap", "PyRayMarching", "PyRayMarchingGPU", "ScanSimulator2D"]
