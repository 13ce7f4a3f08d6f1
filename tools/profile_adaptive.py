"""Per-kernel-class device times of the adaptive 2D path (BASELINE configs[3]) and of the
large-frame path with tonal reduction / whitening (configs[4]) on 60 s of audio.

    python tools/profile_adaptive.py"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

import libspecbleach_b200 as sb  # noqa: E402
from libspecbleach_b200 import api  # noqa: E402
from libspecbleach_b200.synth import nonstationary_noise, speech_plus_noise  # noqa: E402

BASE = dict(learn_noise=0, reduction_amount=20.0, smoothing_factor=1.0, whitening_factor=50.0,
            nlm_masking_protection=0.5, suppression_strength=50.0, aggressiveness=0.0,
            tonal_reduction=6.0)
CASES = [
    ("config4 spp 44.1k/50ms", 44100, 50.0, 0, dict(BASE, adaptive_noise=1, noise_estimation_method=0), "noise"),
    ("config4 martin", 44100, 50.0, 0, dict(BASE, adaptive_noise=1, noise_estimation_method=2), "noise"),
    ("brandt", 44100, 50.0, 0, dict(BASE, adaptive_noise=1, noise_estimation_method=1), "noise"),
    ("config5 96k/77ms tonal12 whit100", 96000, 77.0, 96000,
     dict(BASE, tonal_reduction=12.0, nlm_masking_protection=1.0, whitening_factor=100.0), "speech"),
]
for name, sr, ms, learn, prm, kind in CASES:
    x = nonstationary_noise(61.0, sr, 4) if kind == "noise" else speech_plus_noise(61.0, sr, 5, hum_hz=60.0)
    d = sb.Denoiser2D(sr, ms)
    for rep in range(2):
        d.reset_profile()
        if learn:
            d.load_parameters(learn_noise=1)
            d.process(x[:learn])
        d.load_parameters(**prm)
        api.lib().specbleach_b200_profile(1)
        api.lib().specbleach_b200_stats_reset()
        t0 = time.perf_counter()
        d.process(x[sr:])
        dt = time.perf_counter() - t0
        cs = api.class_stats()
        api.lib().specbleach_b200_profile(0)
    tot = sum(v[0] for v in cs.values())
    print(f"{name}: 60 s mono in {dt*1e3:.1f} ms -> {60/dt:.0f}x realtime; kernel time {tot:.2f} ms")
    for k, (msv, n) in sorted(cs.items(), key=lambda kv: -kv[1][0])[:8]:
        print(f"    {k:14s} {msv:8.3f} ms {n:4d} launches {100*msv/tot:5.1f}%")
    d.close()
