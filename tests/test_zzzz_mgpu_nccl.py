"""NCCL interface exchange under the driver-run suite: spawns tests/mgpu_parity.py with torch.distributed.run on
min(2, visible) GPUs -- every rank evaluates one rhs! and three CK2N54 steps on its own partition and compares with the
oracle's all-ranks restatement (deterministic DSS: bit-exact on every rank; atomics DSS and the interface-first split:
<= 1e-12 per node, <= 1e-10 relative L2).  Skipped on a box with fewer than two GPUs, where the single-rank periodic
self-exchange tests (tests/test_gpu_parity.py) cover the same assemble() code path without NCCL."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_nccl_interface_exchange_parity_two_ranks():
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one box")
    env = dict(os.environ)
    env.setdefault("MASTER_ADDR", "127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mgpu_parity.py")]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("-> OK") >= 10 and "FAIL" not in r.stdout      # 5 cases x 2 ranks
