"""CPU tests: the oracle itself.  The numpy restatement (oracle/specbleach_np.py) is
pinned against golden vectors generated from the reference build, and -- where
oracle/_ref is present -- against the reference build directly, including the
known-answer checks the reference's own unit tests hold for this path."""
import math

import numpy as np
import pytest

from common import GOLDEN_CASES, assert_parity, load_golden, run_flow
from oracle import specbleach_np as snp


class NpDenoiser:
    """adapter: OracleDenoiser with the (x, chunk) process signature of the bindings"""

    def __init__(self, sr, ms, two_d):
        self.d = snp.OracleDenoiser(sr, ms, two_d)

    def load_parameters(self, **kw):
        self.d.load_parameters(**kw)

    def process(self, x, chunk=None):
        chunk = chunk or len(x)
        return np.concatenate([self.d.process(x[i:i + chunk]) for i in range(0, len(x), chunk)])


@pytest.mark.parametrize("name", ["g1d_44k_demo", "g1d_48k_martin", "g2d_48k_manual", "g2d_48k_brandt",
                                  "g1d_44k_brandt_learned"])
def test_restatement_matches_golden(name):
    two_d, sr, ms, learn, prm, chunk = GOLDEN_CASES[name]
    g = load_golden(name)
    y = run_flow(NpDenoiser(sr, ms, two_d), g["x"], learn, prm, 7000)
    assert_parity(y, g["y"], name)


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_reference_build_matches_golden(ref, name):
    two_d, sr, ms, learn, prm, chunk = GOLDEN_CASES[name]
    g = load_golden(name)
    d = ref.RefDenoiser(sr, ms, two_d)
    y = run_flow(d, g["x"], learn, prm, chunk)
    assert_parity(y, g["y"], name, max_abs=1e-6, snr_db=120.0)
    assert d.latency() == int(g["latency"]) and d.profile_size() == int(g["profile_size"])


def test_reference_chunking_invariance(ref):
    """stft_processor.c:134: any chunking of the input gives the same output."""
    two_d, sr, ms, learn, prm, _ = GOLDEN_CASES["g1d_44k_demo"]
    g = load_golden("g1d_44k_demo")
    a = run_flow(ref.RefDenoiser(sr, ms, two_d), g["x"], learn, prm, 512)
    b = run_flow(ref.RefDenoiser(sr, ms, two_d), g["x"], learn, prm, 4096)
    assert np.array_equal(a, b)


def test_nlm_restatement_matches_golden():
    g = load_golden("nlm_module")
    s = g["snr"]
    for i, j in enumerate(range(20, s.shape[0])):
        out = snp.nlm_frame(s[j - 20:j + 1], float(g["h"]))
        assert np.max(np.abs(out - g["out"][i])) <= 2e-6 * max(1.0, float(np.max(np.abs(g["out"][i]))))


def test_nlm_reference_matches_golden(ref):
    g = load_golden("nlm_module")
    out = ref.ref_nlm(g["snr"], float(g["h"]))
    assert np.isnan(out[:20]).all()          # ring not full: nlm_filter.c:238-255
    assert np.array_equal(out[20:], g["out"])


def test_fast_exp_known_answers():
    """tests/test_shared_utils.c:252-269: within 0.05 of expf on [-5, 0], exactly 0 below -20."""
    x = np.linspace(-5.0, 0.0, 101, dtype=np.float32)
    assert np.max(np.abs(snp.fast_exp(x) - np.exp(x))) < 0.05
    assert np.all(snp.fast_exp(np.array([-20.5, -100.0, -1e9], dtype=np.float32)) == 0.0)


def test_berouti_alpha_known_answers():
    """tests/test_suppression_engine.c:50-89: alpha = 4 at very low SNR (strength 1),
    1 at high SNR, linear in between."""
    d = snp.OracleDenoiser(48000, 50.0, False)
    d.load_parameters(suppression_strength=100.0, tonal_reduction=0.0)
    K = d.g.K
    noise = np.full(K, 1.0, dtype=np.float32)
    lo = d._alpha(np.full(K, 1e-3, dtype=np.float32), noise)
    hi = d._alpha(np.full(K, 1e4, dtype=np.float32), noise)
    mid = d._alpha(np.full(K, 10.0 ** 0.75, dtype=np.float32), noise)   # 7.5 dB -> halfway
    assert np.allclose(lo, 4.0) and np.allclose(hi, 1.0) and np.allclose(mid, 2.5, atol=1e-3)


def test_snr_and_reconstruct_known_answers():
    """tests/test_nlm_filter.c:452-497: snr = P / max(noise, 1e-9); mag = snr * max(noise, 1e-9)."""
    P = np.array([4.0, 2.0, 1.0, 0.0], dtype=np.float32)
    n = np.array([2.0, 0.5, 0.0, 1.0], dtype=np.float32)
    snr = P / np.maximum(n, np.float32(1e-9))
    assert np.allclose(snr[:2], [2.0, 4.0]) and snr[2] == np.float32(1.0) / np.float32(1e-9)
    assert np.allclose(snr * np.maximum(n, np.float32(1e-9)), P)


def test_geometry_matches_survey_table():
    """SURVEY.md appendix C (computed with the reference's own float arithmetic)."""
    rows = [(48000, 50, 2400, 3200, 600, 1601), (44100, 50, 2205, 3006, 551, 1504),
            (44100, 20, 882, 1682, 220, 842), (96000, 77, 7392, 8192, 1848, 4097),
            (48000, 46, 2208, 3008, 552, 1505)]
    for sr, ms, frame, N, hop, K in rows:
        g = snp.make_geometry(sr, float(ms))
        assert (g.frame, g.N, g.hop, g.K) == (frame, N, hop, K)
    g = snp.make_geometry(48000, 50.0)
    assert list(g.band_start[:4]) == [0, 3, 7, 10] and int(g.band_end[-1]) == 1601
    assert int(g.band_start[-1]) == 517


def test_fftw_shim_matches_numpy(ref):
    import ctypes as C
    lib = ref.load()
    lib.fftwf_plan_r2r_1d.restype = C.c_void_p
    lib.fftwf_plan_r2r_1d.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_uint]
    lib.fftwf_execute.argtypes = [C.c_void_p]
    rng = np.random.default_rng(0)
    for n in (3200, 3006, 1682):
        x = rng.standard_normal(n).astype(np.float32)
        y = np.zeros(n, np.float32)
        z = np.zeros(n, np.float32)
        pf = lib.fftwf_plan_r2r_1d(n, x.ctypes.data, y.ctypes.data, 0, 64)
        pb = lib.fftwf_plan_r2r_1d(n, y.ctypes.data, z.ctypes.data, 1, 64)
        lib.fftwf_execute(pf)
        lib.fftwf_execute(pb)
        X = np.fft.rfft(x.astype(np.float64))
        hc = np.concatenate([X.real, X.imag[1:n // 2][::-1]])
        assert np.max(np.abs(y - hc)) <= 1e-6 * np.max(np.abs(hc))
        assert np.max(np.abs(z / n - x)) <= 1e-6     # tests/test_fft.c:79-105 asks 1e-4


def test_latency_quirk(ref):
    """Reported 2D latency = frame + 4*(fft_size/4) (specbleach_2d_denoiser.c:63,102-118)."""
    d = ref.RefDenoiser(48000, 50.0, True)
    assert d.latency() == 2400 + 4 * (3200 // 4) and d.profile_size() == 3200
    d1 = ref.RefDenoiser(48000, 50.0, False)
    assert d1.latency() == 2400 and d1.profile_size() == 1601
    assert math.isclose(float(snp.db_to_coeff(-20.0)), 0.1, rel_tol=1e-6)


def test_reference_variant_builds_load_and_differ(ref):
    """oracle/variants: rebuilds of the reference with one compile-time strategy selector pinned
    (N4).  Every variant must load, run both denoisers and differ from the shipped build (the
    transient-aware smoother only exists in the 1D processor)."""
    import numpy as np
    from libspecbleach_b200.synth import speech_plus_noise
    from common import P1D, P2D, run_flow
    if not all(ref.variant_path(v).exists() for v in ref.VARIANTS):
        pytest.skip("oracle/_ref/variants not built (make -C oracle variants)")
    sr = 48000
    x = speech_plus_noise(2.5, sr, 3)
    for two_d, prm in ((True, P2D), (False, dict(P1D, smoothing_factor=40.0))):
        base = run_flow(ref.RefDenoiser(sr, 50.0, two_d), x, sr, prm, 4096)
        for v in ref.VARIANTS:
            y = run_flow(ref.RefDenoiser(sr, 50.0, two_d, variant=v), x, sr, prm, 4096)
            assert np.all(np.isfinite(y))
            changed = float(np.max(np.abs(y - base))) > 1e-4
            assert changed == (not (v == "transient" and two_d)), (v, two_d)
