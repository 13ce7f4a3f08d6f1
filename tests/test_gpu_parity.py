"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C ABI, against
the oracle on identical inputs -- golden vectors, the reference build (oracle/_ref,
prebuilt, travels with the snapshot) and size-independent properties at the
BASELINE sizes.  Tolerance (BASELINE.json north_star): max abs error <= 1e-4 and
residual SNR >= 90 dB; byte/index results (block counts, sizes, latencies) exact."""
import ctypes as C

import numpy as np
import pytest

from common import (GOLDEN_CASES, P1D, P2D, assert_parity, assert_parity_discontinuous, load_golden,
                    parity, run_flow)

pytestmark = pytest.mark.gpu

# bench.py's parameters for configs[1] (tonal_reduction 0, unlike P2D)
P2D_BENCH = dict(P2D, tonal_reduction=0.0)


def synth(seconds, sr, seed):
    from libspecbleach_b200.synth import speech_plus_noise
    return speech_plus_noise(seconds, sr, seed)


# ----------------------------- single kernels --------------------------------
@pytest.mark.parametrize("n", [3200, 3006, 8192, 1682, 2564, 1760, 3008, 6010, 5114, 9620, 1814])
def test_rfft_kernels(gpu, n):
    """forward_kernel / inverse_kernel vs numpy pocketfft in double (fft_transform.c:98-103,
    tests/test_fft.c:79-105: forward o backward = N*x within 1e-4)."""
    from libspecbleach_b200 import api
    x = np.random.default_rng(n).standard_normal((4, n)).astype(np.float32)
    hc, rt = api.debug_rfft(x)
    X = np.fft.rfft(x.astype(np.float64), axis=1)
    ref = np.concatenate([X.real, X.imag[:, 1:n // 2][:, ::-1]], axis=1)
    assert np.abs(hc - ref).max() <= 3e-6 * np.abs(ref).max()
    assert np.abs(rt / n - x).max() <= 1e-5


def test_nlm_kernel_vs_golden(gpu):
    from libspecbleach_b200 import api
    g = load_golden("nlm_module")
    out = api.debug_nlm(g["snr"], float(g["h"]))
    assert np.isnan(out[:20]).all()
    assert np.abs(out[20:] - g["out"]).max() <= 1e-5 * max(1.0, float(np.abs(g["out"]).max()))


@pytest.mark.parametrize("F,K,h", [(48, 1601, 1.0), (40, 203, 1.0), (40, 1504, 2.0), (30, 4097, 0.7),
                                    (24, 16, 1.0), (22, 9, 3.0)])
def test_nlm_kernel_vs_reference_module(gpu, ref, F, K, h):
    """nlm_kernel (+ tail kernel when K % 8 != 0) vs nlm_filter_process on random maps with a
    smooth ridge, a quiet region, an all-zero region (block skip) and the clamped edges."""
    from libspecbleach_b200 import api
    rng = np.random.default_rng(K)
    s = rng.exponential(1.0, size=(F, K)).astype(np.float32)
    w = s[:, K // 3:K // 3 + 40].shape[1]
    s[:, K // 3:K // 3 + 40] = 20.0 + 0.05 * rng.standard_normal((F, w)).astype(np.float32)
    s[:, K // 2:K // 2 + 100] *= 0.01
    s[:, 2 * K // 3:2 * K // 3 + 24] = 0.0
    s[F // 2:, :16] = 0.3
    want = ref.ref_nlm(s, h)
    got = api.debug_nlm(s, h)
    err = np.abs(got[20:] - want[20:]) / (np.abs(want[20:]) + 1.0)
    assert err.max() <= 2e-5


# ----------------------------- whole path ------------------------------------
@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_cuda_matches_golden(gpu, name):
    two_d, sr, ms, learn, prm, chunk = GOLDEN_CASES[name]
    g = load_golden(name)
    d = gpu.Denoiser(sr, ms, two_d)
    y = run_flow(d, g["x"], learn, prm, chunk)
    assert_parity(y, g["y"], name)
    assert d.latency() == int(g["latency"]) and d.profile_size() == int(g["profile_size"])


def test_learned_profiles_match_golden(gpu):
    two_d, sr, ms, learn, prm, chunk = GOLDEN_CASES["g2d_48k_manual"]
    g = load_golden("g2d_48k_manual")
    d = gpu.Denoiser(sr, ms, two_d)
    d.load_parameters(learn_noise=1)
    d.process(g["x"][:learn], chunk)
    for m in (1, 2, 3, 4):
        want = g["profiles_after_learn"][m - 1]
        got = d.get_profile(m)
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 2e-5 * max(1e-12, float(np.abs(want).max())), m
        assert d.block_count(m) == int(g["blocks"][m - 1])
        assert d.profile_available(m)
    assert np.all(d.get_profile(4) == 0.0)            # quirk Q7: MINIMUM stays 0


CASES = [
    ("2d_manual", True, 48000, 50.0, P2D, None),
    ("2d_tonal12_chunk777", True, 48000, 50.0, dict(P2D, tonal_reduction=12.0), 777),
    ("2d_median_residual", True, 48000, 50.0, dict(P2D, aggressiveness=-0.5, residual_listen=True), 4096),
    ("2d_transparent", True, 48000, 50.0, dict(P2D, reduction_amount=0.0, tonal_reduction=0.0), None),
    ("2d_h2_whit100", True, 48000, 50.0, dict(P2D, smoothing_factor=2.0, whitening_factor=100.0), None),
    ("1d_manual", False, 48000, 50.0, P1D, None),
    ("1d_smooth0_chunk512", False, 48000, 50.0, dict(P1D, smoothing_factor=0.0), 512),
    ("1d_tonal12", False, 48000, 50.0, dict(P1D, tonal_reduction=12.0), 4800),
    ("1d_tonal6_residual", False, 48000, 50.0, dict(P1D, tonal_reduction=6.0, residual_listen=True,
                                                     aggressiveness=-0.3), 999),
    ("1d_residual_44k1_20ms", False, 44100, 20.0, dict(P1D, residual_listen=True), None),
    ("2d_44k1_20ms", True, 44100, 20.0, P2D, 3001),
    ("1d_44k1_50ms", False, 44100, 50.0, P1D, None),
    ("2d_96k_77ms", True, 96000, 77.0, dict(P2D, tonal_reduction=12.0, nlm_masking_protection=1.0,
                                           whitening_factor=100.0), None),
]


@pytest.mark.parametrize("name,two_d,sr,ms,prm,chunk", CASES, ids=[c[0] for c in CASES])
def test_cuda_matches_reference_build(gpu, ref, name, two_d, sr, ms, prm, chunk):
    x = synth(5.0 if sr < 96000 else 3.0, sr, 20260002)
    learn = sr
    want = run_flow(ref.RefDenoiser(sr, ms, two_d), x, learn, prm, 4096)
    d = gpu.Denoiser(sr, ms, two_d)
    got = run_flow(d, x, learn, prm, chunk)
    assert_parity(got, want, name)


GEOMETRIES = [(8000, 32.0), (22050, 46.0), (44100, 23.2), (48000, 5.0), (48000, 100.0), (88200, 50.0),
              (192000, 20.0), (192000, 100.0), (96000, 170.0)]


@pytest.mark.parametrize("sr,ms", GEOMETRIES, ids=[f"{a}Hz_{b}ms" for a, b in GEOMETRIES])
def test_unusual_geometries_match_reference_build(gpu, ref, sr, ms):
    """frame sizes from 240 to 19 200 samples, fft sizes with factors 2, 3, 5, 7, 11, 19, 107, 521
    ... (fft_transform.c:44-63: N = frame + 800 rounded to even): latency, profile size and PCM
    of both denoisers (tools: tests/gpu_geometry_sweep.py)."""
    x = synth(2.5, sr, 7)
    for two_d in (True, False):
        r = ref.RefDenoiser(sr, ms, two_d)
        d = gpu.Denoiser(sr, ms, two_d)
        assert (r.latency(), r.profile_size()) == (d.latency(), d.profile_size())
        prm = P2D if two_d else P1D
        want = run_flow(r, x, sr, prm, 8192)
        got = run_flow(d, x, sr, prm, 7001)
        assert_parity(got, want, f"{sr} Hz / {ms} ms two_d={two_d}")
        d.close()
        r.close()


ADAPTIVE = [("2d_spp", True, 0, 0), ("2d_martin", True, 2, 0), ("1d_spp", False, 0, 0),
            ("1d_martin", False, 2, 0), ("2d_spp_learned", True, 0, 44100),
            ("2d_brandt", True, 1, 0), ("1d_brandt", False, 1, 0), ("2d_brandt_learned", True, 1, 44100)]


@pytest.mark.parametrize("name,two_d,method,learn", ADAPTIVE, ids=[c[0] for c in ADAPTIVE])
def test_adaptive_estimators_match_reference_build(gpu, ref, name, two_d, method, learn):
    from libspecbleach_b200.synth import nonstationary_noise
    x = nonstationary_noise(8.0, 44100, 20260004)
    prm = dict(P2D if two_d else P1D, adaptive_noise=1, noise_estimation_method=method)
    got = run_flow(gpu.Denoiser(44100, 50.0, two_d), x, learn, prm, 10000)
    if method == 1:   # Brandt: discontinuous estimator, see assert_parity_discontinuous
        how = assert_parity_discontinuous(
            got, lambda v: run_flow(ref.RefDenoiser(44100, 50.0, two_d), v, learn, prm, None), x, name)
        print(name, how)
        return
    want = run_flow(ref.RefDenoiser(44100, 50.0, two_d), x, learn, prm, None)
    assert_parity(got, want, name)


def test_brandt_long_history_and_silence(gpu, ref):
    """Brandt with 20 ms frames (history 262 frames, ring wraps twice), a silent gap
    (frames skipped by the silence gate keep the last estimate, brandt:169-173) and
    odd chunking; then streaming invariance of the estimator state, bit-exact."""
    from libspecbleach_b200.synth import nonstationary_noise
    sr = 44100
    x = nonstationary_noise(12.0, sr, 77)
    x[3 * sr:3 * sr + 30000] = 0.0
    prm = dict(P1D, adaptive_noise=1, noise_estimation_method=1)
    got = run_flow(gpu.Denoiser(sr, 20.0, False), x, 0, prm, 7777)
    assert_parity_discontinuous(
        got, lambda v: run_flow(ref.RefDenoiser(sr, 20.0, False), v, 0, prm, None), x, "brandt 20 ms")
    for chunk in (None, 220, 50000):
        assert np.array_equal(got, run_flow(gpu.Denoiser(sr, 20.0, False), x, 0, prm, chunk))


def test_silence_and_edge_inputs(gpu, ref):
    """digital silence (estimator silence gate, NLM block skip), a single sample per call,
    and a chunk shorter than one hop (tests/test_nlm_filter.c:340-389)."""
    sr = 48000
    x = synth(3.0, sr, 5)
    x[sr + 20000:sr + 60000] = 0.0
    for two_d, prm in ((True, dict(P2D, adaptive_noise=1, noise_estimation_method=0)), (False, P1D)):
        want = run_flow(ref.RefDenoiser(sr, 50.0, two_d), x, sr, prm, 4096)
        got = run_flow(gpu.Denoiser(sr, 50.0, two_d), x, sr, prm, 599)
        assert_parity(got, want, f"silence two_d={two_d}")
    d = gpu.Denoiser2D(sr, 50.0)
    d.load_parameters(**P2D)
    y = d.process(x[:50], 1)
    assert np.all(y == 0.0)                      # latency: first frame of output is zero


def test_streaming_invariance_is_bit_exact(gpu):
    """any chunking of the input gives the identical output (stft_processor.c:134)."""
    sr = 48000
    x = synth(6.0, sr, 9)
    outs = []
    for chunk in (None, 777, 4096, 100000):
        outs.append(run_flow(gpu.Denoiser2D(sr, 50.0), x, sr, P2D, chunk))
    for o in outs[1:]:
        assert np.array_equal(outs[0], o)


def test_in_place_and_api_state(gpu):
    """in == out is allowed (test_specbleach_2d_denoiser.c:258-264); profile load/reset and
    mode validation (:79-164)."""
    sr = 48000
    x = synth(2.0, sr, 3)
    d = gpu.Denoiser2D(sr, 50.0)
    d.load_parameters(learn_noise=1)
    buf = x.copy()
    assert d.process_raw(len(buf), buf.ctypes.data, buf.ctypes.data)
    d2 = gpu.Denoiser2D(sr, 50.0)
    d2.load_parameters(learn_noise=1)
    assert np.array_equal(buf, d2.process(x))
    n = d.profile_size()
    assert n == 3200 and d.block_count(1) == 160 and d.profile_available(1)
    assert not d.load_profile(np.ones(n - 1, np.float32), 3, 1)         # size mismatch
    assert not d.load_profile(np.ones(n, np.float32), 3, 0) and not d.load_profile(np.ones(n, np.float32), 3, 5)
    assert d.load_profile(np.full(n, 0.25, np.float32), 7, 3)
    assert d.block_count(3) == 7 and np.all(d.get_profile(3) == 0.25)
    assert d.reset_profile() and d.block_count(1) == 0 and not d.profile_available(1)
    assert np.all(d.get_profile(1) == 0)
    assert d.block_count(9) == 0 and not d.profile_available(0)


def test_loaded_profile_is_used_as_is(gpu, ref):
    """a profile restored with load_noise_profile_for_mode replaces learning."""
    sr = 48000
    x = synth(3.0, sr, 12)
    src = ref.RefDenoiser(sr, 50.0, True)
    src.load_parameters(learn_noise=1)
    src.process(x[:sr])
    src.load_parameters(**P2D)
    src.process(x[sr:sr + 4800])                      # finalises the profiles
    profs = [src.get_profile(m) for m in (1, 2, 3, 4)]
    outs = []
    for mk in (lambda: ref.RefDenoiser(sr, 50.0, True), lambda: gpu.Denoiser2D(sr, 50.0)):
        d = mk()
        for m, p in zip((1, 2, 3, 4), profs):
            assert d.load_profile(p, 80, m)
        d.load_parameters(**P2D)
        outs.append(d.process(x[sr:]))
    assert_parity(outs[1], outs[0], "loaded profile")


def test_batch_and_device_entry_points_equal_per_handle_calls(gpu):
    import torch
    sr = 48000
    xs = [synth(4.0, sr, 100 + i) for i in range(5)]
    solo = [run_flow(gpu.Denoiser2D(sr, 50.0), x, sr, P2D, None) for x in xs]
    # host batch
    ds = [gpu.Denoiser2D(sr, 50.0) for _ in xs]
    outs = [np.empty_like(x) for x in xs]
    for d in ds:
        d.load_parameters(learn_noise=1)
    assert gpu.process_batch(ds, [x[:sr] for x in xs], [o[:sr] for o in outs])
    for d in ds:
        d.load_parameters(**P2D)
    assert gpu.process_batch(ds, [x[sr:] for x in xs], [o[sr:] for o in outs])
    for a, b in zip(solo, outs):
        assert np.array_equal(a, b)
    # device-resident batch
    ds = [gpu.Denoiser2D(sr, 50.0) for _ in xs]
    xd = [torch.from_numpy(x).cuda() for x in xs]
    yd = [torch.empty_like(t) for t in xd]
    for d in ds:
        d.load_parameters(learn_noise=1)
    assert gpu.process_batch(ds, [(t.data_ptr(), sr) for t in xd], [(t.data_ptr(), sr) for t in yd], device=True)
    for d in ds:
        d.load_parameters(**P2D)
    n = len(xs[0]) - sr
    assert gpu.process_batch(ds, [(t.data_ptr() + 4 * sr, n) for t in xd],
                             [(t.data_ptr() + 4 * sr, n) for t in yd], device=True)
    torch.cuda.synchronize()
    for a, t in zip(solo, yd):
        assert np.array_equal(a, t.cpu().numpy())


def test_full_size_properties_config2(gpu, ref):
    """BASELINE configs[1] at full size (10 min, 48 kHz, 48 000 frames per channel):
    causality/prefix property anchors the long run to the oracle -- the first 12 s of
    the 600 s output must be bit-identical to a 12 s run, which is checked against the
    reference build; plus determinism, finiteness and noise reduction."""
    sr = 48000
    x = np.tile(synth(60.0, sr, 20260002), 10)
    d = gpu.Denoiser2D(sr, 50.0)
    y = run_flow(d, x, sr, P2D, None)
    assert np.all(np.isfinite(y))
    short = run_flow(gpu.Denoiser2D(sr, 50.0), x[:12 * sr], sr, P2D, None)
    assert np.array_equal(y[:12 * sr], short)
    want = run_flow(ref.RefDenoiser(sr, 50.0, True), x[:12 * sr], sr, P2D, 8192)
    assert_parity(short, want, "config2 prefix")
    y2 = run_flow(gpu.Denoiser2D(sr, 50.0), x, sr, P2D, 1 << 20)
    assert np.array_equal(y, y2)                                    # determinism + chunking
    assert float(np.mean(y[2 * sr:] ** 2)) < 0.96 * float(np.mean(x[2 * sr:] ** 2))  # test_audio_regression.c:237


def _streams(count, seconds, sr, seed):
    """`count` distinct mono streams built cheaply from a few synthesised ones (time shift + gain)."""
    base = [synth(seconds + 1.0, sr, seed + i) for i in range(4)]
    n = int(seconds * sr)
    out = []
    for i in range(count):
        b = base[i % 4]
        off = (37 * i) % sr
        out.append(np.ascontiguousarray(b[off:off + n] * np.float32(0.6 + 0.4 * ((i * 7) % 11) / 10.0)))
    return out


def test_full_size_properties_config3(gpu, ref):
    """BASELINE configs[2], one GPU's share at 8 GPUs: 512 independent 60 s mono streams,
    2D NLM, through specbleach_b200_process_batch in waves of 128.  Anchors: (a) stream
    independence -- every 64th stream equals the same stream processed on its own handle,
    bit-exact; (b) one stream's 8 s prefix against the reference build; (c) finiteness."""
    sr, secs, wave = 48000, 60.0, 128
    checked = 0
    for w0 in range(0, 512, wave):
        xs = _streams(wave, secs, sr, 20260003 + w0)
        ds = [gpu.Denoiser2D(sr, 50.0) for _ in xs]
        ys = [np.empty_like(x) for x in xs]
        for d in ds:
            d.load_parameters(learn_noise=1)
        assert gpu.process_batch(ds, [x[:sr] for x in xs], [y[:sr] for y in ys])
        for d in ds:
            d.load_parameters(**P2D)
        assert gpu.process_batch(ds, [x[sr:] for x in xs], [y[sr:] for y in ys])
        for d in ds:
            d.close()
        assert all(np.isfinite(y).all() for y in ys)
        for i in range(0, wave, 64):
            solo = run_flow(gpu.Denoiser2D(sr, 50.0), xs[i], sr, P2D, 1 << 19)
            assert np.array_equal(solo, ys[i])
            checked += 1
        if w0 == 0:
            n = 8 * sr
            want = run_flow(ref.RefDenoiser(sr, 50.0, True), xs[5][:n], sr, P2D, 8192)
            assert_parity(ys[5][:n], want, "config3 stream 5 prefix")
    assert checked == 8


def test_full_size_properties_config4(gpu, ref):
    """BASELINE configs[3]: 1 h of non-stationary noise at 44.1 kHz, adaptive SPP-MMSE and
    Martin (288 131 frames x 1504 bins through the per-bin sequential scans).  Anchors: the
    first 10 s of the 1 h output are bit-identical to a 10 s run, which is checked against
    the reference build; a second chunking gives the identical hour; finiteness."""
    from libspecbleach_b200.synth import nonstationary_noise
    sr = 44100
    x = nonstationary_noise(3600.0, sr, 20260004)
    for method in (0, 2):
        prm = dict(P2D, adaptive_noise=1, noise_estimation_method=method)
        y = run_flow(gpu.Denoiser2D(sr, 50.0), x, 0, prm, None)
        assert np.all(np.isfinite(y))
        n = 10 * sr
        short = run_flow(gpu.Denoiser2D(sr, 50.0), x[:n], 0, prm, None)
        assert np.array_equal(y[:n], short)
        want = run_flow(ref.RefDenoiser(sr, 50.0, True), x[:n], 0, prm, None)
        assert_parity(short, want, f"config4 method {method} prefix")
        y2 = run_flow(gpu.Denoiser2D(sr, 50.0), x, 0, prm, 3_000_000)
        assert np.array_equal(y, y2)


def test_full_size_properties_config5(gpu, ref):
    """BASELINE configs[4]: 96 kHz, 77 ms frames (fft 8192, 4097 bins), 8 channels x 120 s, 2D
    with tonal reduction 12 dB, full masking protection and 100 % whitening, learned profile
    with mains hum.  Anchors: channel independence (batch == solo, bit-exact), a 4 s prefix of
    one channel against the reference build, finiteness."""
    from libspecbleach_b200.synth import speech_plus_noise
    sr, secs = 96000, 120.0
    prm = dict(P2D, tonal_reduction=12.0, nlm_masking_protection=1.0, whitening_factor=100.0)
    xs = [speech_plus_noise(secs, sr, 20260005 + c, hum_hz=60.0) for c in range(8)]
    ds = [gpu.Denoiser2D(sr, 77.0) for _ in xs]
    assert ds[0].profile_size() == 8192
    ys = [np.empty_like(x) for x in xs]
    for d in ds:
        d.load_parameters(learn_noise=1)
    assert gpu.process_batch(ds, [x[:sr] for x in xs], [y[:sr] for y in ys])
    for d in ds:
        d.load_parameters(**prm)
    assert gpu.process_batch(ds, [x[sr:] for x in xs], [y[sr:] for y in ys])
    assert all(np.isfinite(y).all() for y in ys)
    solo = run_flow(gpu.Denoiser2D(sr, 77.0), xs[3], sr, prm, 1 << 20)
    assert np.array_equal(solo, ys[3])
    n = 4 * sr
    want = run_flow(ref.RefDenoiser(sr, 77.0, True), xs[6][:n], sr, prm, 8192)
    assert_parity(ys[6][:n], want, "config5 channel 6 prefix")


def test_smoke_entry_point(gpu):
    import __graft_entry__ as ge
    ge.smoke()


def _record(name, payload):
    """measurements made as a side effect of a test, kept for BASELINE.md / DESIGN.md"""
    import json
    from pathlib import Path
    out = Path(__file__).resolve().parent.parent / "gpurun_out"
    if out.is_dir():
        (out / f"{name}.json").write_text(json.dumps(payload, indent=1))


def test_full_length_config2_matches_reference_build(gpu, ref):
    """the full-size tests anchored only a prefix to the oracle.  Here ONE WHOLE channel
    of the configs[1] bench input (600 s, 48 000 frames: learn 1 s + denoise 599 s, >= 4 internal
    sub-calls, every NLM tile boundary) goes through the reference build (4 OpenMP threads, the
    library default) and through the CUDA path, and every output sample is compared."""
    import time
    sr = 48000
    x = synth(600.0, sr, 20260002)                     # channel 0 of bench.py's file on rank 0
    ref.set_threads(4)
    r = ref.RefDenoiser(sr, 50.0, True)
    t0 = time.perf_counter()
    want = run_flow(r, x, sr, P2D_BENCH, 1 << 20)
    t_ref = time.perf_counter() - t0
    ref.set_threads(1)
    d = gpu.Denoiser2D(sr, 50.0)
    t0 = time.perf_counter()
    got = run_flow(d, x, sr, P2D_BENCH, None)
    t_gpu = time.perf_counter() - t0
    err, snr = assert_parity(got, want, "config 2, whole 600 s channel")
    n_bad = int(np.sum(np.abs(got.astype(np.float64) - want) > 1e-5))
    print(f"config 2 full length: max_abs={err:.3e} snr={snr:.1f} dB, samples off by > 1e-5: {n_bad}; "
          f"reference {600.0 / t_ref:.1f}x realtime mono (4 OMP threads, {t_ref:.1f} s), "
          f"CUDA path through specbleach_2d_process with pageable buffers {600.0 / t_gpu:.0f}x")
    _record("full_length_config2", {"max_abs": err, "snr_db": snr, "samples_gt_1e-5": n_bad,
                                    "ref_seconds_4_omp_threads": t_ref, "gpu_seconds_dropin_pageable": t_gpu,
                                    "audio_seconds": 600.0})


def test_long_adaptive_config4_matches_reference_build(gpu, ref, tmp_path):
    """10 min of the configs[3] input (44.1 kHz, non-stationary noise) through the adaptive 2D path
    with SPP-MMSE: 48 000 frames of a per-bin recursion that feeds back its own estimate, long
    enough for a rounding-level difference to drift if it were going to.

    (a) with a profile learned on the first second the tonal detector works on the static MAX
        profile and the path is continuous: the WHOLE output must meet the BASELINE tolerance,
        and the last minute must be as good as the first (no drift).
    (b) without a learned profile (the bench flow) the tonal mask is re-detected on every frame's
        noise estimate with a hard 1.41 peak ratio (tonal_detector.c:24-162), and the REFERENCE
        ITSELF is not reproducible at the tolerance: a 1-ULP perturbation of its input moves it by
        2.5e-4 / 85 dB over these 10 minutes (measured).  No implementation
        with another FFT rounding can do better than that, so the CUDA path is held to the
        reference's own self-consistency: not further from the reference than the perturbed
        reference is (6 dB margin), and out of tolerance in at most two more minutes."""
    import subprocess
    import sys
    from pathlib import Path
    from libspecbleach_b200.synth import nonstationary_noise
    sr = 44100
    n = 60 * sr
    x = nonstationary_noise(600.0, sr, 20260004)
    prm = dict(P2D, adaptive_noise=1, noise_estimation_method=0)
    # three 10-minute reference runs side by side, each in its own process (tests/ref_long_run.py)
    helper = str(Path(__file__).resolve().parent / "ref_long_run.py")
    jobs = [(sr, 0, str(tmp_path / "learned.npy")), (0, 0, str(tmp_path / "exact.npy")),
            (0, 1, str(tmp_path / "perturbed.npy"))]
    procs = [subprocess.Popen([sys.executable, helper, str(a), str(b), c]) for a, b, c in jobs]
    for p in procs:
        assert p.wait(timeout=900) == 0
    want_l, want, pert = (np.load(c) for _, _, c in jobs)
    # (a) learned profile
    got_l = run_flow(gpu.Denoiser2D(sr, 50.0), x, sr, prm, None)
    err, snr = assert_parity(got_l, want_l, "config 4 input, learned profile, 10 min SPP-MMSE")
    e0, s0 = assert_parity(got_l[:n], want_l[:n], "first minute")
    e1, s1 = assert_parity(got_l[-n:], want_l[-n:], "last minute")
    print(f"10 min SPP-MMSE, learned profile: max_abs={err:.3e} snr={snr:.1f} dB; first minute {s0:.1f} dB, "
          f"last minute {s1:.1f} dB")
    # (b) unlearned: against the reference's own reproducibility
    got = run_flow(gpu.Denoiser2D(sr, 50.0), x, 0, prm, None)
    assert np.all(np.isfinite(got))
    e_c, s_c = parity(got, want)
    e_r, s_r = parity(pert, want)
    bad_c = bad_r = 0
    for m in range(10):
        ec, sc = parity(got[m * n:(m + 1) * n], want[m * n:(m + 1) * n])
        er, sr_ = parity(pert[m * n:(m + 1) * n], want[m * n:(m + 1) * n])
        bad_c += not (ec <= 1e-4 and sc >= 90.0)
        bad_r += not (er <= 1e-4 and sr_ >= 90.0)
    print(f"10 min SPP-MMSE, no learned profile: CUDA vs reference {e_c:.3e} / {s_c:.1f} dB "
          f"({bad_c} of 10 minutes out of tolerance); reference vs 1-ULP-perturbed reference "
          f"{e_r:.3e} / {s_r:.1f} dB ({bad_r} minutes)")
    assert (e_c <= 1e-4 and s_c >= 90.0) or s_c >= s_r - 6.0
    assert bad_c <= bad_r + 2
    _record("long_adaptive_config4", {
        "learned": {"max_abs": err, "snr_db": snr, "first_minute_snr_db": s0, "last_minute_snr_db": s1},
        "unlearned_cuda_vs_reference": {"max_abs": e_c, "snr_db": s_c, "minutes_out_of_tolerance": int(bad_c)},
        "unlearned_reference_vs_perturbed_reference": {"max_abs": e_r, "snr_db": s_r,
                                                       "minutes_out_of_tolerance": int(bad_r)}})
