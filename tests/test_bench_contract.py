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


def test_workloads_follow_the_baseline_configs():
    """bench.py --config N: geometry, stream counts and audio accounting of BASELINE configs 2-5;
    config 3 shards its 4096 streams over the ranks without gaps or overlaps (no data-path
    collective: every rank owns a contiguous block)."""
    import sys
    sys.path.insert(0, str(ROOT))
    import bench
    from libspecbleach_b200.sharding import shard_bounds
    w2 = bench.Workload(2, 8, 3)
    assert (w2.sr, w2.frame, w2.fft, w2.hop, w2.bins) == (48000, 2400, 3200, 600, 1601)
    assert w2.n_streams == 2 and w2.job_audio_seconds() == 8 * 600.0 and w2.scaling == "weak"
    for world in (1, 2, 4, 8, 3):
        owned = []
        for rank in range(world):
            w = bench.Workload(3, world, rank)
            lo, hi = shard_bounds(4096, world, rank)
            assert w.n_streams == hi - lo and w.wave == 128
            owned += list(range(lo, hi))
            assert w.job_audio_seconds() == 4096 * 60.0 and w.scaling == "strong"
        assert owned == list(range(4096))
    w4 = bench.Workload(4, 1, 0)
    assert (w4.sr, w4.frame, w4.fft, w4.hop, w4.bins) == (44100, 2205, 3006, 551, 1504)
    assert w4.n_learn == 0 and [p["noise_estimation_method"] for p in w4.params] == [0, 2]
    assert w4.n // w4.hop == 288130 or w4.n // w4.hop == 288131
    w5 = bench.Workload(5, 1, 0)
    assert (w5.frame, w5.fft, w5.hop, w5.bins) == (7392, 8192, 1848, 4097) and w5.n_streams == 8
    assert w5.job_audio_seconds() == 120.0
    # both arms describe the workload identically (the driver compares their `config`)
    assert bench.Workload(3, 4, 0).config_dict() == bench.Workload(3, 4, 2).config_dict()
