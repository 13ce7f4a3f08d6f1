"""bench.py's JSON-line contract, as far as it can be checked without a GPU: the reference arm
(`--impl reference`: the unmodified reference sources built by oracle/Makefile, timed on the
host cores) and the refusal of the product arm when there is no CUDA device."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def run_bench(*args, env=None):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True,
                          cwd=str(ROOT), env=env, timeout=600)


def test_reference_arm_line_has_the_contract_keys():
    from oracle import ref_binding as rb
    if not rb.available():
        pytest.skip("oracle/_ref not built (run __graft_entry__.build())")
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample-seconds", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["higher_is_better"] is True and line["unit"] == "x realtime"
    assert line["value"] > 0 and line["steps"] == 1
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == line["value"]
    e = line["e2e"]
    assert e["value"] == line["value"] and e["unit"] == line["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]
    assert line.get("gpu_launches", 0) == 0          # nothing of ours runs on this arm


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run_bench("--steps", "1", "--warmup", "0", "--no-cpu-baseline")
    assert r.returncode != 0                           # no CPU fallback: it must fail loudly
    assert "CUDA" in (r.stderr + r.stdout) or "device" in (r.stderr + r.stdout)
