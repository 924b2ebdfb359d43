"""Multi-GPU parity under pytest: spawns tests/multigpu_check.py under torch.distributed.run for world
sizes 2, 4 and 8 when the box has that many GPUs (skipped otherwise).  The script compares, bit for bit
with the single-GPU result, ShardedScanner's NCCL gathers (whole and chunk-pipelined), the fused
peer-memory gather for the fan and repeat_angles forms, back-to-back fused calls, a batch gathered in
pieces, and ShardedRollout."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_paths_match_single_gpu(world):
    from pyracecarsimulator_b200 import _native
    if _native.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {_native.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert f"MULTIGPU CHECK PASSED (world {world})" in r.stdout, tail
