"""Deterministic synthetic workloads: speech-like harmonic
segments plus stationary coloured noise (+ optional hum), first second noise only.
Pure numpy; used by bench.py and the parity tests so that both arms see identical
input."""
from __future__ import annotations

import numpy as np


def _one_pole(x: np.ndarray, a: float) -> np.ndarray:
    # y[n] = x[n] + a*y[n-1]  (vectorised in blocks through scipy if available)
    try:
        from scipy.signal import lfilter
        return lfilter([1.0], [1.0, -a], x)
    except Exception:  # pragma: no cover
        y = np.empty_like(x)
        acc = 0.0
        for i, v in enumerate(x):
            acc = v + a * acc
            y[i] = acc
        return y


def speech_plus_noise(seconds: float, sr: int = 48000, seed: int = 20260002,
                      hum_hz: float | None = 50.0, noise_dbfs: float = -32.0,
                      hum_dbfs: float = -45.0) -> np.ndarray:
    """One channel of the config-2/3 workload, float32 in [-1, 1]."""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * sr))
    t = np.arange(n, dtype=np.float64) / sr
    sig = np.zeros(n, dtype=np.float64)
    pos = int(1.0 * sr)                      # first second: noise only
    weights = np.array([1.0, 0.8, 0.6, 0.9, 0.7, 0.4, 0.3, 0.35, 0.2, 0.15, 0.1, 0.08])
    f0 = rng.uniform(90.0, 250.0)
    while pos < n:
        on = int(rng.uniform(0.3, 1.2) * sr)
        off = int(rng.uniform(0.1, 0.6) * sr)
        end = min(n, pos + on)
        m = end - pos
        if m > 16:
            # slow f0 random walk, 12 harmonics, 4 Hz syllabic envelope
            steps = rng.normal(0.0, 0.6, size=m // 480 + 2).cumsum()
            f0 = float(np.clip(f0 + steps[-1] * 0.1, 90.0, 250.0))
            f_inst = np.clip(f0 + np.interp(np.arange(m), np.arange(steps.size) * 480.0, steps),
                             90.0, 250.0)
            phase = 2.0 * np.pi * np.cumsum(f_inst) / sr
            seg = np.zeros(m)
            for hi, w in enumerate(weights, start=1):
                seg += w * np.sin(hi * phase + rng.uniform(0, 2 * np.pi))
            env = 0.5 * (1.0 - np.cos(2.0 * np.pi * 4.0 * np.arange(m) / sr))
            ramp = np.minimum(1.0, np.minimum(np.arange(m), np.arange(m)[::-1]) / (0.01 * sr))
            seg *= env * ramp
            peak = np.max(np.abs(seg)) + 1e-12
            sig[pos:end] += 0.5 * seg / peak
        pos = end + off
    noise = _one_pole(rng.standard_normal(n), 0.7)
    noise *= 10.0 ** (noise_dbfs / 20.0) / (np.sqrt(np.mean(noise ** 2)) + 1e-20)
    out = sig + noise
    if hum_hz:
        out += np.sqrt(2.0) * 10.0 ** (hum_dbfs / 20.0) * np.sin(2.0 * np.pi * hum_hz * t)
    return np.clip(out, -1.0, 1.0).astype(np.float32)


def nonstationary_noise(seconds: float, sr: int = 44100, seed: int = 20260004) -> np.ndarray:
    """Config-4 workload: white noise with a 0.05 Hz +-9 dB level sweep, level steps
    every 120 s and intermittent tone bursts."""
    rng = np.random.default_rng(seed)
    n = int(round(seconds * sr))
    t = np.arange(n, dtype=np.float64) / sr
    level_db = -30.0 + 9.0 * np.sin(2.0 * np.pi * 0.05 * t)
    nsteps = int(seconds // 120) + 1
    steps = rng.uniform(-6.0, 6.0, size=nsteps)
    level_db += steps[np.minimum((t // 120).astype(int), nsteps - 1)]
    x = rng.standard_normal(n) * 10.0 ** (level_db / 20.0)
    pos = int(2.0 * sr)
    while pos < n:
        m = min(n - pos, int(rng.uniform(0.2, 0.8) * sr))
        f = rng.uniform(300.0, 4000.0)
        x[pos:pos + m] += 0.1 * np.sin(2.0 * np.pi * f * t[:m]) * np.hanning(m)
        pos += m + int(rng.uniform(1.0, 5.0) * sr)
    return np.clip(x, -1.0, 1.0).astype(np.float32)
