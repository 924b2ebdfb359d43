"""Host logic of the numpy-buffer registry in pyracecarsimulator_b200.range_libc (no GPU: the C ABI is
replaced by a recording stub).  The registry page-locks a caller's array the SECOND time it sees it, keeps
the array alive, never evicts, and only takes memory numpy itself owns."""
import ctypes as C

import numpy as np
import pytest

from pyracecarsimulator_b200 import _native, range_libc


class _StubLib:
    def __init__(self, fail=False, already=False):
        self.calls, self.fail, self.already = [], fail, already

    def rl_host_register(self, device, ptr, nbytes, was):
        self.calls.append(("reg", device, ptr, nbytes))
        if self.already:
            was._obj.value = 1
        return _native.RL_ERR_CUDA if self.fail else _native.RL_OK

    def rl_host_unregister(self, device, ptr):
        self.calls.append(("unreg", device, ptr))
        return _native.RL_OK


@pytest.fixture
def reg(monkeypatch):
    stub = _StubLib()
    monkeypatch.setattr(_native, "lib", lambda: stub)
    monkeypatch.setattr(_native, "_lib", stub)
    r = range_libc._HostRegistry()
    r.enabled = True
    return r, stub


def big(n=1 << 19):
    return np.zeros(n, dtype=np.float32)     # 2 MiB


def test_registers_on_the_second_sighting_and_keeps_the_array(reg):
    r, stub = reg
    a = big()
    r.note(a, 0)
    assert stub.calls == [] and r.registered_bytes() == 0
    r.note(a, 0)
    assert stub.calls == [("reg", 0, a.ctypes.data, a.nbytes)] and r.registered_bytes() == a.nbytes
    r.note(a, 0)
    r.note(a[:], 0)                                      # another view of the same memory: same key, no new call
    assert len(stub.calls) == 1
    import sys
    assert sys.getrefcount(a) >= 3                       # the registry holds a reference: memory cannot be freed
    r.release()
    assert stub.calls[-1] == ("unreg", 0, a.ctypes.data) and r.registered_bytes() == 0


def test_small_foreign_overlapping_and_refused_buffers_are_left_alone(reg):
    r, stub = reg
    small = np.zeros(1000, dtype=np.float32)
    for _ in range(3):
        r.note(small, 0)
    backing = C.create_string_buffer(4 << 20)            # memory numpy does not own
    foreign = np.frombuffer(backing, dtype=np.float32)
    for _ in range(3):
        r.note(foreign, 0)
    assert stub.calls == []
    a = big()
    r.note(a, 0); r.note(a, 0)
    part = a[: a.size // 2]                              # overlaps a registered range with a different size
    r.note(part, 0); r.note(part, 0); r.note(part, 0)
    assert len(stub.calls) == 1
    stub.fail = True
    b = big()
    r.note(b, 0); r.note(b, 0); r.note(b, 0)             # the driver refuses: tried once, then left alone
    assert [c[0] for c in stub.calls].count("reg") == 2 and r.registered_bytes() == a.nbytes
    stub.fail, stub.already = False, True
    c = big()
    r.note(c, 0); r.note(c, 0); r.note(c, 0)             # page-locked by someone else: not ours to unregister
    assert r.registered_bytes() == a.nbytes
    r.release()
    assert [c[0] for c in stub.calls].count("unreg") == 1


def test_caps_hold_and_nothing_is_evicted(reg):
    r, stub = reg
    keep = [big() for _ in range(r.MAX_ENTRIES + 3)]
    for a in keep:
        r.note(a, 0); r.note(a, 0)
    assert r.registered_bytes() == r.MAX_ENTRIES * keep[0].nbytes
    assert [c[0] for c in stub.calls].count("unreg") == 0
    r.MAX_BYTES = r.registered_bytes()                   # byte cap
    r.release()
    r.MAX_BYTES = 3 << 20
    for a in keep[:3]:
        r.note(a, 0); r.note(a, 0)
    assert r.registered_bytes() == keep[0].nbytes        # 2 MiB fits, 4 MiB would not


def test_disabled_by_environment(monkeypatch):
    monkeypatch.setenv("RL_HOST_REGISTER", "0")
    r = range_libc._HostRegistry()
    a = big()
    r.note(a, 0); r.note(a, 0)
    assert r.registered_bytes() == 0
