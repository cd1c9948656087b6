"""bench.py's reference arm on the host cores (the smallest workload, so the CPU suite stays short): one JSON line with
the keys the measurement contract names, timed on the unmodified reference (oracle/_ref/abm_ref) when it is built."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "pp8-4", "--steps", "1",
                          "--warmup", "3"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "chain-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 1e4 and d["steps"] == 1 and d["warmup"] == 3
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if (ROOT / "oracle" / "_ref" / "abm_ref").exists():
        assert cb["kind"] == "reference"


def test_ours_arm_fails_loudly_without_a_device():
    """No CPU fallback: without a CUDA device the product arm exits with an error instead of printing a number."""
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode != 0
    assert not [ln for ln in out.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_runs_on_rank_zero_only():
    """Under torchrun (N > 1) rank 0 alone runs and prints the reference line; the other ranks exit 0 without work."""
    import os
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "pp8-4"], env=env,
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
