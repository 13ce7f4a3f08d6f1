"""60 s of 48 kHz mono through the 1D denoiser with the Brandt estimator (for ncu)."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import libspecbleach_b200 as sb  # noqa: E402
from libspecbleach_b200 import api  # noqa: E402
from libspecbleach_b200.synth import nonstationary_noise  # noqa: E402

sr = 48000
x = nonstationary_noise(60.0, sr, 5)
d = sb.Denoiser(sr, 50.0, False)
d.load_parameters(learn_noise=0, reduction_amount=20.0, smoothing_factor=0.0, whitening_factor=50.0,
                  masking_depth=0.5, suppression_strength=50.0, aggressiveness=1.0,
                  adaptive_noise=1, noise_estimation_method=1)
d.process(x[:5 * sr])
for rep in range(2):
    api.lib().specbleach_b200_profile(1)
    api.lib().specbleach_b200_stats_reset()
    t0 = time.perf_counter()
    d.process(x)
    dt = time.perf_counter() - t0
    cs = api.class_stats()
    api.lib().specbleach_b200_profile(0)
    print(f"rep {rep}: 60 s in {dt*1e3:.1f} ms; adaptive class {cs['adaptive'][0]:.2f} ms")
