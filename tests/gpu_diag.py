"""GPU-side diagnostic sweep (not a pytest): runs every stage of the CUDA path
against the reference build (oracle/_ref) and prints one JSON line per check.

    gpurun -- 'python tests/gpu_diag.py > gpurun_out/diag.jsonl'

Used while developing kernels; the assertions proper live in tests/test_gpu_*.py.
"""
from __future__ import annotations

import json
import sys
import time
import traceback
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import libspecbleach_b200 as sb  # noqa: E402
from libspecbleach_b200 import api  # noqa: E402
from libspecbleach_b200.synth import nonstationary_noise, speech_plus_noise  # noqa: E402
from oracle import ref_binding as rb  # noqa: E402


def emit(**kw):
    print(json.dumps(kw), flush=True)


def cmp(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = a - b
    num = float(np.sum(b * b))
    den = float(np.sum(d * d))
    snr = 10.0 * np.log10(num / den) if den > 0 else float("inf")
    return {"max_abs": float(np.max(np.abs(d))) if d.size else 0.0,
            "snr_db": snr if np.isfinite(snr) else 999.0,
            "n_gt_1e-5": int(np.sum(np.abs(d) > 1e-5)),
            "first_bad": int(np.argmax(np.abs(d) > 1e-5)) if np.any(np.abs(d) > 1e-5) else -1}


def section(fn):
    try:
        fn()
    except Exception as e:  # noqa: BLE001
        emit(check=fn.__name__, error=repr(e), tb=traceback.format_exc()[-1500:])


def check_fft():
    rng = np.random.default_rng(1)
    for n in (3200, 3006, 8192, 1682, 2564, 1760):
        x = rng.standard_normal((3, n)).astype(np.float32)
        hc, rt = api.debug_rfft(x)
        X = np.fft.rfft(x.astype(np.float64), axis=1)
        ref = np.concatenate([X.real, X.imag[:, 1:n // 2][:, ::-1]], axis=1)
        emit(check="rfft", n=n, fwd_rel=float(np.abs(hc - ref).max() / np.abs(ref).max()),
             roundtrip=float(np.abs(rt / n - x).max()))


def snr_map(F, K, seed):
    rng = np.random.default_rng(seed)
    s = rng.exponential(1.0, size=(F, K)).astype(np.float32)
    # a smooth "tonal" ridge, a quiet region and an all-zero region
    s[:, K // 3:K // 3 + 40] = 20.0 + 0.05 * rng.standard_normal((F, 40)).astype(np.float32)
    s[:, K // 2:K // 2 + 100] *= 0.01
    s[:, 2 * K // 3:2 * K // 3 + 24] = 0.0
    s[F // 2:, :16] = 0.3
    return s


def check_nlm():
    for (F, K, h) in ((48, 1601, 1.0), (40, 203, 1.0), (40, 1504, 2.0), (30, 4097, 0.7)):
        s = snr_map(F, K, 7)
        ref = rb.ref_nlm(s, h)
        t0 = time.perf_counter()
        out = api.debug_nlm(s, h)
        dt = time.perf_counter() - t0
        r = cmp(out[20:], ref[20:])
        rel = float(np.max(np.abs(out[20:] - ref[20:]) / (np.abs(ref[20:]) + 1e-6)))
        emit(check="nlm", F=F, K=K, h=h, rel=rel, secs=dt, **r)


PARAMS_2D = dict(learn_noise=0, reduction_amount=20.0, smoothing_factor=1.0,
                 whitening_factor=50.0, nlm_masking_protection=0.5, suppression_strength=50.0,
                 aggressiveness=0.0, tonal_reduction=0.0)
PARAMS_1D = dict(learn_noise=0, reduction_amount=20.0, smoothing_factor=0.0,
                 whitening_factor=50.0, masking_depth=0.5, suppression_strength=50.0,
                 aggressiveness=1.0, tonal_reduction=0.0)


def run_pair(two_d, sr, ms, x, learn_samples, params, chunk_ref=None, chunk_gpu=None):
    ref = rb.RefDenoiser(sr, ms, two_d)
    gpu = sb.Denoiser(sr, ms, two_d)
    outs = []
    for d, chunk in ((ref, chunk_ref), (gpu, chunk_gpu)):
        ys = []
        if learn_samples:
            d.load_parameters(learn_noise=1)
            ys.append(d.process(x[:learn_samples], chunk))
        d.load_parameters(**params)
        ys.append(d.process(x[learn_samples:], chunk))
        outs.append(np.concatenate(ys))
    profs = [(ref.get_profile(m), gpu.get_profile(m)) for m in (1, 2, 3, 4)]
    meta = dict(lat_ref=ref.latency(), lat_gpu=gpu.latency(), size_ref=ref.profile_size(),
                size_gpu=gpu.profile_size(),
                blocks=[(ref.block_count(m), gpu.block_count(m)) for m in (1, 2, 3, 4)])
    ref.close()
    gpu.close()
    return outs[0], outs[1], profs, meta


def check_learn():
    x = speech_plus_noise(3.0, 48000, 11)
    for two_d in (True, False):
        ref = rb.RefDenoiser(48000, 50.0, two_d)
        gpu = sb.Denoiser(48000, 50.0, two_d)
        ref.load_parameters(learn_noise=1)
        gpu.load_parameters(learn_noise=1)
        yr = ref.process(x, 4096)
        yg = gpu.process(x, 5000)
        emit(check="learn_passthrough", two_d=two_d, **cmp(yg, yr))
        for m in (1, 2, 3, 4):
            pr, pg = ref.get_profile(m), gpu.get_profile(m)
            rel = float(np.max(np.abs(pg - pr) / (np.abs(pr) + 1e-12)))
            emit(check="learn_profile", two_d=two_d, mode=m, rel=rel,
                 blocks=(ref.block_count(m), gpu.block_count(m)),
                 avail=(ref.profile_available(m), gpu.profile_available(m)), **cmp(pg, pr))
        ref.close()
        gpu.close()


def check_denoise():
    x = speech_plus_noise(8.0, 48000, 20260002)
    cases = [
        ("2d_manual", True, 48000, 50.0, PARAMS_2D, None, None),
        ("2d_manual_chunked", True, 48000, 50.0, PARAMS_2D, 4096, 777),
        ("2d_tonal12", True, 48000, 50.0, dict(PARAMS_2D, tonal_reduction=12.0), None, None),
        ("2d_aggr-0.5_resid", True, 48000, 50.0,
         dict(PARAMS_2D, aggressiveness=-0.5, residual_listen=True), None, None),
        ("1d_manual", False, 48000, 50.0, PARAMS_1D, None, None),
        ("1d_smooth50_chunk512", False, 48000, 50.0, dict(PARAMS_1D, smoothing_factor=50.0), 512, 512),
        ("2d_44k1_20ms", True, 44100, 20.0, PARAMS_2D, None, 3001),
        ("1d_44k1_50ms", False, 44100, 50.0, PARAMS_1D, None, None),
    ]
    for name, two_d, sr, ms, prm, cr, cg in cases:
        try:
            xs = x if sr == 48000 else speech_plus_noise(6.0, sr, 5)
            yr, yg, profs, meta = run_pair(two_d, sr, ms, xs, sr, prm, cr, cg)
            emit(check="denoise", case=name, rms_ref=float(np.sqrt(np.mean(yr ** 2))), **cmp(yg, yr), **meta)
            for m, (pr, pg) in zip((1, 2, 3, 4), profs):
                emit(check="denoise_profile", case=name, mode=m, **cmp(pg, pr))
        except Exception as e:  # noqa: BLE001
            emit(check="denoise", case=name, error=repr(e), tb=traceback.format_exc()[-800:])


def check_adaptive():
    x = nonstationary_noise(8.0, 44100, 20260004)
    for name, two_d, method, learn in (("2d_spp", True, 0, 0), ("2d_martin", True, 2, 0),
                                       ("1d_spp", False, 0, 0), ("1d_martin", False, 2, 0),
                                       ("2d_spp_learned", True, 0, 44100)):
        try:
            base = PARAMS_2D if two_d else PARAMS_1D
            prm = dict(base, adaptive_noise=1, noise_estimation_method=method)
            yr, yg, _, meta = run_pair(two_d, 44100, 50.0, x, learn, prm, None, 10000)
            emit(check="adaptive", case=name, rms_ref=float(np.sqrt(np.mean(yr ** 2))), **cmp(yg, yr))
        except Exception as e:  # noqa: BLE001
            emit(check="adaptive", case=name, error=repr(e), tb=traceback.format_exc()[-800:])


def check_brandt():
    x = nonstationary_noise(6.0, 44100, 20260004)
    for name, two_d, ms, learn in (("2d_brandt", True, 50.0, 0), ("1d_brandt_20ms", False, 20.0, 0),
                                   ("2d_brandt_learned", True, 50.0, 44100)):
        try:
            base = PARAMS_2D if two_d else PARAMS_1D
            prm = dict(base, adaptive_noise=1, noise_estimation_method=1)
            t0 = time.perf_counter()
            yr, yg, _, meta = run_pair(two_d, 44100, ms, x, learn, prm, None, 10000)
            emit(check="brandt", case=name, secs=time.perf_counter() - t0,
                 rms_ref=float(np.sqrt(np.mean(yr ** 2))), **cmp(yg, yr))
        except Exception as e:  # noqa: BLE001
            emit(check="brandt", case=name, error=repr(e), tb=traceback.format_exc()[-800:])
    # throughput of the estimator on its own: 60 s mono, 1D (no NLM)
    sr = 48000
    xs = nonstationary_noise(60.0, sr, 5)
    d = sb.Denoiser(sr, 50.0, False)
    d.load_parameters(**dict(PARAMS_1D, adaptive_noise=1, noise_estimation_method=1))
    d.process(xs[:sr * 5])
    api.lib().specbleach_b200_profile(1)
    api.lib().specbleach_b200_stats_reset()
    t0 = time.perf_counter()
    d.process(xs)
    dt = time.perf_counter() - t0
    emit(check="brandt_speed", audio_s=60.0, secs=dt, x_realtime=60.0 / dt,
         classes={k: v[0] for k, v in api.class_stats().items()})
    api.lib().specbleach_b200_profile(0)
    d.close()


def check_speed():
    sr = 48000
    x = speech_plus_noise(61.0, sr, 3)
    gpu = sb.Denoiser2D(sr, 50.0)
    gpu.load_parameters(learn_noise=1)
    gpu.process(x[:sr])
    gpu.load_parameters(**PARAMS_2D)
    body = x[sr:]
    gpu.process(body[:sr * 5])       # warm-up (allocations, first launches)
    api.lib().specbleach_b200_profile(1)
    api.lib().specbleach_b200_stats_reset()
    t0 = time.perf_counter()
    gpu.process(body)
    dt = time.perf_counter() - t0
    st = api.stats()
    emit(check="speed_host_api", audio_s=len(body) / sr, secs=dt, x_realtime=len(body) / sr / dt, **st)
    api.lib().specbleach_b200_profile(0)
    gpu.close()


if __name__ == "__main__":
    emit(check="env", devices=sb.device_count(), ref=rb.available())
    only = set(sys.argv[1:])
    for fn in (check_fft, check_nlm, check_learn, check_denoise, check_adaptive, check_brandt,
               check_speed):
        if not only or fn.__name__ in only:
            section(fn)
    emit(check="done")
