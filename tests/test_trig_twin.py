"""The march kernels evaluate glibc's sinf/cosf algorithm on the device (csrc/glibc_trig.cuh) so
that ray directions carry the same bits as the host libm the reference calls.  The C twin of
that code (oracle/trig_twin.c) is compared with this host's libm here; the device code is
compared with the twin in tests/test_gpu_march.py::test_device_trig_matches_twin."""
import math

import numpy as np


def test_twin_matches_host_libm_bit_for_bit(orc):
    # every 4099th float bit pattern over the whole finite range, both signs (incl. |y| >= 120)
    assert orc.twin_mismatches(0, 4099, 600_000) == 0
    # dense over the headings the scan path produces: |theta_g| up to ~20 rad
    hi = int(np.float32(20.0).view(np.uint32))
    assert orc.twin_mismatches(0, 211, hi // 211 + 1) == 0
    # every float in [1, 1 + 2^-4) and around the pi/4 and 120 branch points
    one = int(np.float32(1.0).view(np.uint32))
    assert orc.twin_mismatches(one, 1, 1 << 19) == 0
    for edge in (float.fromhex("0x1.921FB6p-1"), 120.0, 2.0 ** -12):
        e = int(np.float32(edge).view(np.uint32))
        assert orc.twin_mismatches(e - 5000, 1, 10000) == 0


def test_twin_special_values(orc):
    assert orc.twin_sincosf(0.0) == (0.0, 1.0)
    s, c = orc.twin_sincosf(float("inf"))
    assert math.isnan(s) and math.isnan(c)
    s, c = orc.twin_sincosf(float("nan"))
    assert math.isnan(s) and math.isnan(c)
    s, c = orc.twin_sincosf(-4.712389)          # rotation_const of every shipped map
    assert abs(s - 1.0) < 1e-6 and abs(c) < 1e-6
