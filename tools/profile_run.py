"""Small, fixed job for ncu: 2D NLM denoiser, 48 kHz mono, learn 1 s + denoise
`seconds` (default 60 s -> 4800 frames, one sub-call: three NLM chunks of 2176 / 2176 / 448
targets).  Same kernels as bench.py, ~40 launches.

    python tools/profile_run.py [seconds]
"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

import libspecbleach_b200 as sb  # noqa: E402
from libspecbleach_b200 import api  # noqa: E402
from libspecbleach_b200.synth import speech_plus_noise  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
sr = 48000
x = speech_plus_noise(seconds + 1.0, sr, 3)
d = sb.Denoiser2D(sr, 50.0)
PARAMS = dict(learn_noise=0, reduction_amount=20.0, smoothing_factor=1.0, whitening_factor=50.0,
              nlm_masking_protection=0.5, suppression_strength=50.0, aggressiveness=0.0,
              tonal_reduction=0.0)
for rep in range(2):
    d.reset_profile()
    d.load_parameters(learn_noise=1)
    d.process(x[:sr])
    d.load_parameters(**PARAMS)
    api.lib().specbleach_b200_profile(1)
    api.lib().specbleach_b200_stats_reset()
    t0 = time.perf_counter()
    y = d.process(x[sr:])
    dt = time.perf_counter() - t0
    st = api.stats()
    print(f"rep {rep}: {seconds:.0f} s mono in {dt*1e3:.2f} ms -> {seconds/dt:.0f}x realtime (mono); "
          f"nlm {st['nlm_ms']:.3f} ms over {st['nlm_launches']} launches; "
          f"{st['kernel_launches']} kernel launches; rms {np.sqrt(np.mean(y**2)):.4f}")
cs = api.class_stats()
tot = sum(v[0] for v in cs.values())
for k, (ms, n) in sorted(cs.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:14s} {ms:8.3f} ms  {n:4d} launches  {100*ms/tot:5.1f}%")
print(f"  sum of kernel time {tot:.3f} ms")
print("fp32 peak TFLOP/s:", api.lib().specbleach_b200_debug_fp32_peak_tflops())
