"""The bench.py contract that can be checked without a GPU: the reference arm (`--impl reference`, the oracle port on
the host cores) prints one JSON line with the agreed keys, and the algorithmic-byte model matches SURVEY.md 8(d)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                                   "--warmup", "0", "--ref-nel", "3"], text=True, cwd=ROOT, timeout=600)
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GDOF/s" and line["higher_is_better"] is True
    assert line["metric"] == "3D Euler nop=4 RHS GDOF/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"]


def test_algorithmic_bytes_model():
    sys.path.insert(0, ROOT)
    import bench
    assert abs(bench.algorithmic_bytes_per_node(4, False) - 259.875) < 1e-9          # 259.9 B/node
    assert abs(bench.algorithmic_bytes_per_node(4, True) - 307.875) < 1e-9
    assert abs(bench.algorithmic_bytes_per_node(7, False) - (88 + (8 / 7) ** 3 * 88)) < 1e-9
    assert abs(bench.elem_kernel_bytes_per_node(4, False) - (40 + 1.953125 * 88)) < 1e-9


def test_overlap_default_and_nccl_channel_cap():
    """bench.py turns the interface-first split on only where there is an exchange to hide (N > 1) and caps NCCL's
    channel count to the SMs the interior launch leaves free -- without overriding a cap the user has set."""
    sys.path.insert(0, ROOT)
    import bench
    env = {}
    assert bench.resolve_overlap(-1, 1, env) == 0 and env == {}
    assert bench.resolve_overlap(-1, 8, env) == 4
    assert env == {"NCCL_MAX_CTAS": "4", "NCCL_MAX_NCHANNELS": "4", "NCCL_MAX_P2P_NCHANNELS": "4"}
    env = {"NCCL_MAX_CTAS": "8"}
    assert bench.resolve_overlap(2, 2, env) == 2 and env["NCCL_MAX_CTAS"] == "8" and env["NCCL_MAX_P2P_NCHANNELS"] == "2"
    assert bench.resolve_overlap(0, 8, {}) == 0
