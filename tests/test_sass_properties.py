"""Properties of the compiled march kernels that the measured numbers depend on, read off the
SASS of the built library with cuobjdump -- no GPU needed.  They guard against a compiler flag, a refactoring or a
toolkit change silently undoing them:
  * the unrolled march loop of the plain kernel is 10 instructions per step, with the NEXT sample's float->int
    conversions issued BEFORE the exit branch of the current sample (the early-address form);
  * the tail's look-ahead touch is a cp.async (LDGSTS), not a load into a register;
  * every march kernel fits 32 registers per thread (16 CTAs of 128 threads per SM).
"""
import os
import re
import shutil
import subprocess

import pytest

from pyracecarsimulator_b200 import _native

PLAIN = "march_pose_kernelILb1ELb0ELb1ELi0ELb1E"   # FAN, no step counter, 32-bit index, local output, padded field

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(_native.LIB_PATH),
                                reason="needs cuobjdump and the built library")


@pytest.fixture(scope="module")
def sass():
    txt = subprocess.run(["cuobjdump", "-sass", _native.LIB_PATH], capture_output=True, text=True, check=True).stdout
    funcs, name = {}, None
    for line in txt.split("\n"):
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
            funcs[name] = []
            continue
        m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(.*?);", line)
        if name and m:
            funcs[name].append(m.group(1).strip())
    return funcs


def _mnemonic(ins):
    ins = re.sub(r"^@!?U?P\d\s+", "", ins)
    return ins.split()[0].split(".")[0]


def test_plain_march_loop_is_ten_instructions_with_the_address_before_the_branch(sass):
    body = next(v for k, v in sass.items() if PLAIN in k)
    ops = [_mnemonic(i) for i in body]
    loads = [i for i, o in enumerate(ops) if o == "LDG"]
    # a step = the instructions from one field load up to (not including) the next one
    steps = [ops[a:b] for a, b in zip(loads[:-1], loads[1:])]
    regular = [s for s in steps if len(s) == 10 and s.count("F2I") == 2 and s.count("BRA") == 1]
    assert len(regular) >= 25, f"only {len(regular)} of {len(steps)} unrolled steps have the 10-instruction shape"
    for s in regular:
        assert sorted(s) == sorted(["LDG", "FADD", "FFMA", "FFMA", "FSETP", "F2I", "F2I", "IMAD", "BRA", "IMAD"]), s
        last_cvt = max(i for i, o in enumerate(s) if o == "F2I")
        assert last_cvt < s.index("BRA"), f"conversion after the exit branch: {s}"


def test_tail_touch_is_an_async_copy(sass):
    for name, body in sass.items():
        if "march_pose_kernel" in name or "march_many_kernel" in name or "march_territory_kernel" in name or "march_crash_kernel" in name:
            ops = [_mnemonic(i) for i in body]
            assert "LDGSTS" in ops, f"{name[:80]}: no cp.async touch in the tail loop"


def test_march_kernels_fit_32_registers():
    txt = subprocess.run(["cuobjdump", "--dump-resource-usage", _native.LIB_PATH], capture_output=True, text=True, check=True).stdout
    seen = 0
    for fn, usage in re.findall(r"Function ([^:\n]+):\n\s*(REG:\d+[^\n]*)", txt):
        if any(k in fn for k in ("march_pose_kernel", "march_many_kernel", "march_territory_kernel", "march_crash_kernel")):
            regs = int(re.search(r"REG:(\d+)", usage).group(1))
            assert regs <= 32, f"{fn[:80]} uses {regs} registers"
            seen += 1
    assert seen >= 20
