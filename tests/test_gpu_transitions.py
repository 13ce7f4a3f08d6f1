"""GPU parity across PARAMETER TRANSITIONS (pytest -m gpu).

SURVEY 8b: load_parameters may change anything between any two process() calls.  The
reference keeps its state machine in spectral_2d_denoiser.c:287-322 / spectral_denoiser.c:225-256
(estimator re-created when the method changes, NLM h only when smoothing_factor > 0) and in
denoiser_profile_core.c:28-102 (learn -> finalize, adaptive seed when the mode was off, manual
branch clears last_adaptive_state).  The CUDA engine mirrors it with version counters and cached
rows (sb_engine.cu: noise_version, noise_hist_version, aux_*_version), which is exactly where a
stale cache would hide -- so one stream is driven through every transition on both sides and the
WHOLE output is compared at the BASELINE tolerance (1e-4 / 90 dB).
"""
import numpy as np
import pytest

from common import (BRANDT_BRANCHES, P1D, P2D, assert_parity, assert_parity_discontinuous, run_script)

pytestmark = pytest.mark.gpu


def _speech(seconds, sr, seed):
    from libspecbleach_b200.synth import speech_plus_noise
    return speech_plus_noise(seconds, sr, seed)


def _script(two_d, sr, with_brandt):
    P = P2D if two_d else P1D
    depth = "nlm_masking_protection" if two_d else "masking_depth"
    A = dict(P, adaptive_noise=1)
    steps = [
        (0.30, "none", {}),                                            # process before any load_parameters
        (1.00, "params", dict(learn_noise=1)),                         # learn
        (1.50, "params", P),                                           # manual denoise
        (1.50, "params", dict(A, noise_estimation_method=0)),          # adaptive on: SPP-MMSE
        (1.50, "params", dict(A, noise_estimation_method=2)),          # method switch: Martin (estimator re-created)
    ]
    if with_brandt:
        steps.append((1.50, "params", dict(A, noise_estimation_method=1)))   # Brandt
    steps += [
        (1.00, "params", P),                                           # adaptive off
        (0.70, "params", dict(P, aggressiveness=-1.0)),                # median profile
        (0.70, "params", dict(P, aggressiveness=1.0)),                 # max profile
        (1.00, "params", dict(A, noise_estimation_method=0, aggressiveness=1.0)),   # adaptive on again: re-seed
        (1.00, "params", dict(A, noise_estimation_method=0, aggressiveness=-0.5)),  # aggressiveness while adaptive
        (0.50, "params", dict(learn_noise=1)),                         # re-enter learn after denoising
        (1.00, "params", dict(P, residual_listen=True)),               # finalize again; residual listen on
        (0.70, "params", P),                                           # residual listen off
        (1.00, "reset", dict(P, smoothing_factor=0.5 if two_d else 60.0)),   # reset_profile mid-stream, h 0.5
        (0.80, "params", dict(P, smoothing_factor=2.0 if two_d else 0.0)),   # h 2.0 (1D: smoothing off)
        (0.60, "params", dict(P, smoothing_factor=0.0)),               # 2D: h <= 0 keeps the previous h
        (0.80, "params", dict(P, tonal_reduction=12.0, whitening_factor=100.0, **{depth: 1.0})),
        (0.80, "params", dict(P, reduction_amount=0.0, tonal_reduction=0.0)),       # transparency guard
        (0.80, "params", dict(A, noise_estimation_method=2, tonal_reduction=12.0)),  # unlearned + adaptive + tonal
    ]
    return {"sr": sr, "steps": steps, "seconds": sum(s[0] for s in steps)}


TRANSITIONS = [("2d_48k", True, 48000, 50.0), ("1d_44k", False, 44100, 50.0)]


@pytest.mark.parametrize("name,two_d,sr,ms", TRANSITIONS, ids=[t[0] for t in TRANSITIONS])
def test_parameter_transitions_match_reference_build(gpu, ref, name, two_d, sr, ms):
    sc = _script(two_d, sr, with_brandt=False)
    x = _speech(sc["seconds"] + 0.01, sr, 20260021)
    want = run_script(ref.RefDenoiser(sr, ms, two_d), x, sc, 4096)
    got_a = run_script(gpu.Denoiser(sr, ms, two_d), x, sc, 599)
    got_b = run_script(gpu.Denoiser(sr, ms, two_d), x, sc, 100000)
    assert np.array_equal(got_a, got_b), "chunking changed the output across parameter transitions"
    err, snr = assert_parity(got_a, want, f"transitions {name}")
    print(f"transitions {name}: max_abs={err:.3e} snr={snr:.1f} dB over {sc['seconds']:.1f} s")
    # every segment on its own as well: a stale cache would hide in a short quiet segment
    pos = 0
    for seconds, action, kw in sc["steps"]:
        n = int(round(seconds * sr))
        a, b = got_a[pos:pos + n], want[pos:pos + n]
        if float(np.sum(b.astype(np.float64) ** 2)) > 1e-9:
            assert_parity(a, b, f"transitions {name}, segment at {pos / sr:.2f} s ({action})",
                          snr_db=80.0)
        pos += n


@pytest.mark.parametrize("name,two_d,sr,ms", TRANSITIONS, ids=[t[0] for t in TRANSITIONS])
def test_parameter_transitions_with_brandt(gpu, ref, name, two_d, sr, ms):
    """the same walk with a Brandt segment between Martin and manual (estimator re-created twice)."""
    sc = _script(two_d, sr, with_brandt=True)
    x = _speech(sc["seconds"] + 0.01, sr, 20260022)
    got = run_script(gpu.Denoiser(sr, ms, two_d), x, sc, 7001)
    assert_parity_discontinuous(
        got, lambda v: run_script(ref.RefDenoiser(sr, ms, two_d), v, sc, 4096), x,
        f"transitions+brandt {name}")


def test_profile_loaded_mid_stream(gpu, ref):
    """load_noise_profile_for_mode between two process() calls of a running stream (2D and 1D),
    then a second, different profile: the cached tonal mask / whitening rows must follow."""
    sr = 48000
    x = _speech(6.0, sr, 31)
    for two_d in (True, False):
        P = dict(P2D if two_d else P1D, tonal_reduction=6.0)
        src = ref.RefDenoiser(sr, 50.0, two_d)
        src.load_parameters(learn_noise=1)
        src.process(x[:sr])
        src.load_parameters(**P)
        src.process(x[sr:sr + 4800])
        K = 1601
        profs = {m: (src.get_profile(m)[:K].copy(), 80) for m in (1, 2, 3)}
        louder = {m: (2.5 * a, b) for m, (a, b) in profs.items()}
        sc = {"sr": sr, "steps": [
            (1.0, "params", P),                                     # unlearned, zero profile
            (1.5, "load", {"profiles": profs, "params": P}),
            (1.5, "load", {"profiles": louder, "params": dict(P, aggressiveness=0.7)}),
            (1.0, "reset", P),
            (1.0, "load", {"profiles": {3: profs[3]}, "params": dict(P, aggressiveness=1.0)}),
        ]}
        want = run_script(ref.RefDenoiser(sr, 50.0, two_d), x, sc, 4096)
        got = run_script(gpu.Denoiser(sr, 50.0, two_d), x, sc, 3001)
        assert_parity(got, want, f"profile loaded mid-stream two_d={two_d}")


def test_adaptive_header_race_stress(gpu):
    """64 adaptive handles at K = 4097 (17 CTAs per scan) in one process_batch over 32 lanes: the
    scans of different handles share the GPU, so a CTA of a scan may be scheduled after another
    CTA of the same scan has retired.  The first-frame flag and Martin's counters are read by
    every CTA at kernel start; they are committed by a follow-up kernel (sb_adaptive.cu), hence
    every handle must reproduce its solo run bit for bit, call after call."""
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import libspecbleach_b200 as sb
from libspecbleach_b200 import api
from libspecbleach_b200.synth import nonstationary_noise
from common import P2D
sb.set_device(0)
assert api.lib().specbleach_b200_configure(512, 32)   # 32 lanes, small sub-calls (scratch ~0.6 GB per lane)
sr, ms, n_handles, calls, seg = 96000, 77.0, 64, 6, 30000
base = [nonstationary_noise(calls * seg / sr + 0.1, sr, 900 + i) for i in range(4)]
xs = []
for i in range(n_handles):
    x = np.roll(base[i %% 4], 977 * i).copy()
    if i %% 3 == 0:
        x[:seg + 5000] = 0.0          # silent start: the first-frame flag stays set across calls
    xs.append(np.ascontiguousarray(x[:calls * seg]))
def prm(i):
    return dict(P2D, adaptive_noise=1, noise_estimation_method=(2 if i %% 2 else 0))
ds = [sb.Denoiser2D(sr, ms) for _ in xs]
for i, d in enumerate(ds):
    d.load_parameters(**prm(i))
ys = [np.empty_like(x) for x in xs]
for c in range(calls):
    sl = slice(c * seg, (c + 1) * seg)
    assert sb.process_batch(ds, [x[sl] for x in xs], [y[sl] for y in ys])
bad = 0
for i in (0, 1, 2, 3, 17, 30, 45, 63):
    d = sb.Denoiser2D(sr, ms)
    d.load_parameters(**prm(i))
    solo = d.process(xs[i], seg)
    if not np.array_equal(solo, ys[i]):
        bad += 1
        print("handle", i, "differs from its solo run:", float(np.abs(solo - ys[i]).max()))
print("BAD", bad)
sys.exit(1 if bad else 0)
''' % (str(root), str(root / "tests"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def test_state_snapshot_after_profile_load_and_corrupt_blobs(gpu):
    """(a) load_profile -> save_state -> load_state -> process must denoise with the
    NEW profile (state_save syncs the device copy first); (b) a truncated / corrupt / mismatched
    blob is rejected and leaves the live handle untouched."""
    sr = 48000
    x = _speech(4.0, sr, 77)
    src = gpu.Denoiser2D(sr, 50.0)
    src.load_parameters(learn_noise=1)
    src.process(x[:sr])
    src.load_parameters(**P2D)
    src.process(x[sr:sr + 4800])
    profs = [src.get_profile(m) for m in (1, 2, 3)]

    def fresh_with_profile():
        d = gpu.Denoiser2D(sr, 50.0)
        for m, p in zip((1, 2, 3), profs):
            assert d.load_profile(p, 80, m)
        d.load_parameters(**P2D)
        return d

    want = fresh_with_profile().process(x[sr:])
    a = fresh_with_profile()
    blob = a.save_state()                      # host mirror is newer than the device copy here
    b = gpu.Denoiser2D(sr, 50.0)
    assert b.load_state(blob)
    assert np.array_equal(b.process(x[sr:]), want)
    assert np.array_equal(a.process(x[sr:]), want)
    for m, p in zip((1, 2, 3), profs):
        assert np.array_equal(b.get_profile(m), p)

    # (b) an adaptive handle in mid-stream; bad blobs must not disturb it
    prm = dict(P2D, adaptive_noise=1, noise_estimation_method=2)
    live = gpu.Denoiser2D(sr, 50.0)
    twin = gpu.Denoiser2D(sr, 50.0)
    for d in (live, twin):
        d.load_parameters(**prm)
        d.process(x[:2 * sr])
    good = live.save_state()
    spp = gpu.Denoiser2D(sr, 50.0)
    spp.load_parameters(**dict(prm, noise_estimation_method=0))
    spp.process(x[:sr])
    other = spp.save_state()                   # another estimator layout: valid, but must be all-or-nothing
    bad = [good[:len(good) // 2], good[:-1], other[:len(other) - 64],
           good[:200] + bytes([good[200] ^ 0x5a]) + good[201:], b"SBST" + bytes(60), b""]
    hdr = bytearray(other)
    hdr[40:48] = (7).to_bytes(8, "little")     # StateHeader.est_floats: far too small for the estimator named
    bad.append(bytes(hdr))
    for blob in bad:
        assert not live.load_state(blob)
        assert gpu.last_error()
    assert np.array_equal(live.process(x[2 * sr:]), twin.process(x[2 * sr:]))
    assert live.load_state(other)              # a valid blob of another estimator still loads
    assert np.array_equal(live.process(x[sr:2 * sr]), spp.process(x[sr:2 * sr]))


def test_brandt_fallback_budget():
    """the Brandt escape hatch (tests/common.py::assert_parity_discontinuous) must
    say which branch it took and stay rare: at most one Brandt comparison of the whole run may
    need the 1-ULP-perturbed reference."""
    for k, v in BRANDT_BRANCHES.items():
        print(f"[brandt] {k}: {v}")
    fallbacks = [k for k, v in BRANDT_BRANCHES.items() if v.startswith("FALLBACK")]
    assert len(fallbacks) <= 1, f"Brandt comparisons that needed the perturbed reference: {fallbacks}"
