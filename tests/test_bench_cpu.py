"""CPU tests of bench.py's contract: the reference arm (the CPU oracle port timed on the host cores)
prints one JSON line with the agreed keys, non-zero ranks of a multi-rank launch print nothing, and
the product arm refuses to run without a CUDA device (there is no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          cwd=ROOT, env=e, timeout=300)


def test_reference_arm_prints_the_contract_line():
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--matches", "48", "--beams", "360", "--base", "3"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "scan matches/sec" and d["unit"] == "matches/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert "sample" in d["cpu_baseline"] and "workload" in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": "matches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_runs_on_rank_zero_only():
    p = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--matches", "16"],
             env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = _run(["--steps", "1", "--warmup", "0", "--matches", "8"])
    assert p.returncode != 0 and "no CPU fallback" in p.stderr
