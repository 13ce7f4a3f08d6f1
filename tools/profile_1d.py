"""Per-kernel-class device times of the 1D denoiser (BASELINE configs[0] geometry: 44.1 kHz mono,
50 ms frames, learned profile) on 60 s of audio.   python tools/profile_1d.py"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import libspecbleach_b200 as sb  # noqa: E402
from libspecbleach_b200 import api  # noqa: E402
from libspecbleach_b200.synth import speech_plus_noise  # noqa: E402

for sr in (44100, 48000):
    x = speech_plus_noise(61.0, sr, 3)
    d = sb.Denoiser1D(sr, 50.0)
    for rep in range(2):
        d.reset_profile()
        d.load_parameters(learn_noise=1)
        d.process(x[:sr])
        d.load_parameters(learn_noise=0, reduction_amount=20.0, smoothing_factor=0.0, whitening_factor=50.0,
                          masking_depth=0.5, suppression_strength=50.0, aggressiveness=1.0)
        api.lib().specbleach_b200_profile(1)
        api.lib().specbleach_b200_stats_reset()
        t0 = time.perf_counter()
        d.process(x[sr:])
        dt = time.perf_counter() - t0
        cs = api.class_stats()
        api.lib().specbleach_b200_profile(0)
    tot = sum(v[0] for v in cs.values())
    print(f"1D {sr} Hz: 60 s mono in {dt*1e3:.2f} ms -> {60/dt:.0f}x realtime; kernel time {tot:.2f} ms")
    for k, (msv, n) in sorted(cs.items(), key=lambda kv: -kv[1][0])[:7]:
        print(f"    {k:14s} {msv:8.3f} ms {n:4d} launches {100*msv/tot:5.1f}%")
