"""GPU tests of the steps on either side of the hot path:
noise-profile persistence, streaming-state snapshot, interleaved PCM front end and the
WAV command-line tool.  Reference behaviour they are held to: get/load profile round trip
(specbleach_2d_denoiser.c:145-206), streaming invariance (stft_processor.c:134) and
libsndfile's sample conversions as used by examples/denoiser_demo.c:217-300."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from common import P1D, P2D, assert_parity, run_flow

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def synth(seconds, sr, seed):
    from libspecbleach_b200.synth import speech_plus_noise
    return speech_plus_noise(seconds, sr, seed)


def test_profile_file_round_trip(gpu, tmp_path):
    """a profile learned on one handle, saved, and restored into fresh handles gives the
    same output as learning in place; wrong geometry / corrupt files are refused."""
    sr = 48000
    x = synth(4.0, sr, 21)
    a = gpu.Denoiser2D(sr, 50.0)
    a.load_parameters(learn_noise=1)
    a.process(x[:sr])
    a.load_parameters(**P2D)
    ya = a.process(x[sr:])                     # finalises the profiles, then denoises
    f = tmp_path / "noise.sbnp"
    assert a.save_profile(f)
    b = gpu.Denoiser2D(sr, 50.0)
    assert b.restore_profile(f)
    for m in (1, 2, 3, 4):
        assert np.array_equal(a.get_profile(m), b.get_profile(m))
        assert a.block_count(m) == b.block_count(m) and a.profile_available(m) == b.profile_available(m)
    b.load_parameters(**P2D)
    yb = b.process(x[sr:])
    c = gpu.Denoiser2D(sr, 50.0)               # the reference's own way: load_noise_profile_for_mode
    for m in (1, 2, 3, 4):
        assert c.load_profile(a.get_profile(m), a.block_count(m), m)
    c.load_parameters(**P2D)
    assert np.array_equal(yb, c.process(x[sr:]))
    # same profile, same input: once the stream history has flushed the outputs agree
    assert np.max(np.abs(ya[sr:] - yb[sr:])) <= 1e-5
    assert not gpu.Denoiser2D(44100, 50.0).restore_profile(f)          # other geometry
    assert not gpu.Denoiser1D(sr, 50.0).restore_profile(f)             # other denoiser
    raw = bytearray(f.read_bytes())
    raw[200] ^= 0x40
    g = tmp_path / "bad.sbnp"
    g.write_bytes(bytes(raw))
    assert not gpu.Denoiser2D(sr, 50.0).restore_profile(g)             # checksum
    assert not gpu.Denoiser2D(sr, 50.0).restore_profile(tmp_path / "missing.sbnp")


@pytest.mark.parametrize("two_d,prm", [(True, P2D), (False, P1D),
                                       (True, dict(P2D, adaptive_noise=1, noise_estimation_method=2)),
                                       (True, dict(P2D, adaptive_noise=1, noise_estimation_method=1))],
                         ids=["2d", "1d", "2d_martin", "2d_brandt"])
def test_state_snapshot_resumes_bit_exact(gpu, two_d, prm):
    """stop a stream mid-way (not on a hop boundary), snapshot, restore into a NEW handle and
    continue: identical to the uninterrupted run."""
    sr = 48000
    x = synth(6.0, sr, 33)
    cut = 3 * sr + 777
    full = run_flow(gpu.Denoiser(sr, 50.0, two_d), x, sr, prm, None)
    a = gpu.Denoiser(sr, 50.0, two_d)
    a.load_parameters(learn_noise=1)
    y0 = a.process(x[:sr])
    a.load_parameters(**prm)
    y1 = a.process(x[sr:cut])
    blob = a.save_state()
    a.close()
    b = gpu.Denoiser(sr, 50.0, two_d)
    assert b.load_state(blob)
    y2 = b.process(x[cut:])
    assert np.array_equal(full, np.concatenate([y0, y1, y2]))
    assert not gpu.Denoiser(44100, 50.0, two_d).load_state(blob)
    assert not b.load_state(blob[:-5] + b"12345")


@pytest.mark.parametrize("fmt", ["f32", "s16", "s24", "s32"])
def test_interleaved_pcm_equals_per_channel(gpu, fmt):
    """3 interleaved channels through specbleach_b200_process_interleaved == converting on the
    host the way libsndfile does and calling specbleach_2d_process per channel."""
    from libspecbleach_b200 import api
    sr, ch = 48000, 3
    xs = np.stack([synth(3.0, sr, 40 + c) for c in range(ch)], axis=1)      # [frames, ch]
    if fmt == "f32":
        pcm, xf = xs.copy(), xs
    elif fmt == "s16":
        pcm = np.round(xs * 32767.0).astype(np.int16)
        xf = (pcm.astype(np.float32) / np.float32(32768.0))
    elif fmt == "s32":
        pcm = np.round(xs.astype(np.float64) * 2147483647.0).astype(np.int32)
        xf = (pcm.astype(np.float64) / 2147483648.0).astype(np.float32)
    else:
        q = np.round(xs.astype(np.float64) * 8388607.0).astype(np.int32)
        b = q.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :3]
        pcm = np.ascontiguousarray(b).reshape(-1)
        xf = (q.astype(np.float64) / 8388608.0).astype(np.float32)
    ds = [gpu.Denoiser2D(sr, 50.0) for _ in range(ch)]
    for d in ds:
        d.load_parameters(learn_noise=1)
    n0 = sr * ch * 3 if fmt == "s24" else sr
    head = api.process_interleaved(ds, pcm[:n0], fmt)
    for d in ds:
        d.load_parameters(**P2D)
    tail = api.process_interleaved(ds, pcm[n0:], fmt)
    got = np.concatenate([head, tail])
    want = np.stack([run_flow(gpu.Denoiser2D(sr, 50.0), np.ascontiguousarray(xf[:, c]), sr, P2D, None)
                     for c in range(ch)], axis=1)
    if fmt == "f32":
        assert np.array_equal(got, want)
    elif fmt == "s16":
        assert np.array_equal(got, np.clip(np.rint(want * np.float32(32767.0)), -32768, 32767).astype(np.int16))
    elif fmt == "s32":
        ref = np.clip(np.rint(want.astype(np.float64) * 2147483647.0), -2147483648.0, 2147483647.0)
        assert np.array_equal(got, ref.astype(np.int32))
    else:
        ref = np.clip(np.rint(want * np.float32(8388607.0)), -8388608, 8388607).astype("<i4")
        assert np.array_equal(got, np.ascontiguousarray(ref.view(np.uint8).reshape(-1, 4)[:, :3]).reshape(-1))


def test_wav_tool_matches_api_and_reference(gpu, ref, tmp_path):
    """tools/sb_denoise on a stereo 16-bit WAV: the output file equals the API path, and its
    left channel is within tolerance of the reference build fed the same decoded samples."""
    from scipy.io import wavfile
    tool = ROOT / "tools" / "sb_denoise"
    if not tool.exists():
        pytest.fail("tools/sb_denoise not built (python -c 'import __graft_entry__ as g; g.build()')")
    sr = 44100
    xs = np.stack([synth(4.0, sr, 50 + c) for c in range(2)], axis=1)
    pcm = np.round(xs * 32767.0).astype(np.int16)
    fin, fout, fprof = tmp_path / "in.wav", tmp_path / "out.wav", tmp_path / "p.sbnp"
    wavfile.write(fin, sr, pcm)
    r = subprocess.run([str(tool), "--learn-seconds", "1.0", "--tonal", "6", "--save-profile", str(fprof),
                        str(fin), str(fout)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    sr2, got = wavfile.read(fout)
    assert sr2 == sr and got.shape == pcm.shape and got.dtype == np.int16
    xf = pcm.astype(np.float32) / np.float32(32768.0)
    want = run_flow(ref.RefDenoiser(sr, 50.0, True), np.ascontiguousarray(xf[:, 0]), sr, P2D, 8192)
    assert_parity(got[:, 0].astype(np.float32) / np.float32(32767.0), want, "wav tool, left channel",
                  max_abs=1e-4 + 1.0 / 32767.0, snr_db=60.0)     # + 16-bit quantisation
    # the saved profile drives a second run without a learn pass
    fout2 = tmp_path / "out2.wav"
    r = subprocess.run([str(tool), "--load-profile", str(fprof), "--tonal", "6", str(fin), str(fout2)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert wavfile.read(fout2)[1].shape == pcm.shape
    bad = subprocess.run([str(tool), str(tmp_path / "nope.wav"), str(fout)], capture_output=True, text=True)
    assert bad.returncode != 0


@pytest.mark.parametrize("two_d,prm", [(True, P2D), (False, P1D)], ids=["2d", "1d"])
def test_time_tiled_stream_equals_monolithic(gpu, two_d, prm):
    """N3: the denoise part of a 40 s stream cut into 4 time tiles (fresh handles, 512 warm-up
    frames, profile handed over) stitches to the monolithic result; two 'ranks' owning two
    tiles each give the same pieces."""
    from libspecbleach_b200.sharding import denoise_tiled, stitch
    sr = 48000
    x = synth(40.0, sr, 61)
    mono = run_flow(gpu.Denoiser(sr, 50.0, two_d), x, sr, prm, None)
    mk = lambda: gpu.Denoiser(sr, 50.0, two_d)
    pieces = denoise_tiled(mk, x, sr, prm, tiles=4, warmup_frames=512)
    tiled = stitch(pieces, len(x))
    err = float(np.max(np.abs(tiled - mono)))
    print("tiled vs monolithic: max abs", err, "bit-exact" if np.array_equal(tiled, mono) else "")
    assert err <= 1e-6
    p0 = denoise_tiled(mk, x, sr, prm, tiles=4, warmup_frames=512, world=2, rank=0)
    p1 = denoise_tiled(mk, x, sr, prm, tiles=4, warmup_frames=512, world=2, rank=1)
    assert sorted(list(p0) + list(p1)) == sorted(pieces)
    both = stitch({**p0, **p1}, len(x))
    assert np.array_equal(both, tiled)


def test_distinct_handles_from_concurrent_host_threads(gpu):
    """SURVEY 8b threading contract: distinct handles are independent and may be driven from
    different host threads (ctypes releases the GIL, so these calls really overlap on the
    host); each thread's output equals the single-threaded result."""
    import threading
    sr = 48000
    xs = [synth(4.0, sr, 70 + i) for i in range(4)]
    want = [run_flow(gpu.Denoiser(sr, 50.0, i % 2 == 0), x, sr, P2D if i % 2 == 0 else P1D, 9000)
            for i, x in enumerate(xs)]
    got = [None] * 4
    errs = []

    def work(i):
        try:
            got[i] = run_flow(gpu.Denoiser(sr, 50.0, i % 2 == 0), xs[i], sr, P2D if i % 2 == 0 else P1D, 9000)
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for a, b in zip(want, got):
        assert np.array_equal(a, b)


def test_handles_on_two_devices_of_one_process(gpu):
    """specbleach_b200_set_device: handles created after it live on that device.  One process
    per GPU is the intended deployment (bench.py, sharding.py), but a host that drives two
    GPUs from one process gets the same samples from either (lanes, scratch and geometry
    tables are per device)."""
    import libspecbleach_b200 as sb
    if sb.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    sr = 48000
    x = synth(6.0, sr, 91)
    try:
        assert sb.set_device(0)
        d0 = gpu.Denoiser(sr, 50.0, True)
        assert sb.set_device(1)
        d1 = gpu.Denoiser(sr, 50.0, True)
        y1a = run_flow(d1, x, sr, P2D, 7777)
        y0 = run_flow(d0, x, sr, P2D, 7777)
        d1b = gpu.Denoiser(sr, 50.0, True)
        y1b = run_flow(d1b, x, sr, P2D, 7777)
    finally:
        sb.set_device(0)
    assert np.array_equal(y0, y1a)
    assert np.array_equal(y0, y1b)


def test_config1_speech_wav_demo_flow(gpu, ref, tmp_path):
    """BASELINE configs[0] as SURVEY 8d defines it: the 1D denoiser on the reference's own test
    recording tests/test_data/Speech.wav (44.1 kHz float32 mono, 16.2 s; a git-ignored copy under
    oracle/_ref/test_data travels to the GPU box) with the flow of examples/denoiser_demo.c:91-99,
    234-300: learn on 8 blocks of 512 samples, then reduction 20 dB, whitening 50 %, smoothing 0,
    masking depth 0.5, aggressiveness 1.0 in 512-sample blocks at the demo's 50 ms frames.
    Compared: the reference build, the C API in 512-sample calls, and tools/sb_denoise --1d."""
    import subprocess
    import time
    from pathlib import Path
    from scipy.io import wavfile
    root = Path(__file__).resolve().parent.parent
    wav = root / "oracle" / "_ref" / "test_data" / "Speech.wav"
    tool = root / "tools" / "sb_denoise"
    if not wav.exists() or not tool.exists():
        pytest.skip("oracle/_ref/test_data/Speech.wav or tools/sb_denoise missing (run __graft_entry__.build())")
    sr, x = wavfile.read(str(wav))
    assert sr == 44100 and x.dtype == np.float32 and x.ndim == 1 and x.shape[0] == 714355
    learn, block = 8 * 512, 512
    prm = dict(learn_noise=0, residual_listen=False, reduction_amount=20.0, smoothing_factor=0.0,
               whitening_factor=50.0, masking_depth=0.5, aggressiveness=1.0)
    r = ref.RefDenoiser(sr, 50.0, False)
    t0 = time.perf_counter()
    want = run_flow(r, x, learn, prm, block)
    t_ref = time.perf_counter() - t0
    assert r.profile_available(1)                      # the demo's own check after learning (:268-272)
    d = gpu.Denoiser1D(sr, 50.0)
    t0 = time.perf_counter()
    got = run_flow(d, x, learn, prm, block)
    t_gpu = time.perf_counter() - t0
    assert d.profile_available(1) and d.block_count(1) == r.block_count(1) == learn // 551
    err, snr = assert_parity(got, want, "config 1 (Speech.wav, demo flow)")
    out = tmp_path / "speech_denoised.wav"
    cmd = [str(tool), "--1d", "--frame-size", "50", "--learn-samples", str(learn), "--reduction", "20",
           "--whitening", "50", "--smoothing", "0", "--masking", "0.5", "--aggressiveness", "1.0",
           "--strength", "0", "--tonal", "0",      # the demo leaves suppression_strength / tonal_reduction at 0
           str(wav), str(out)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    sr2, y = wavfile.read(str(out))
    assert sr2 == sr and y.dtype == np.float32 and y.shape == x.shape
    assert np.array_equal(y, got)                      # tool == API, bit for bit (chunking-invariant)
    print(f"config 1: max_abs={err:.3e} snr={snr:.1f} dB; reference {x.shape[0] / sr / t_ref:.1f}x realtime "
          f"(1 thread, 512-sample calls), CUDA path through 512-sample calls {x.shape[0] / sr / t_gpu:.1f}x")
    import json
    rec = root / "gpurun_out"
    if rec.is_dir():
        (rec / "config1_speech_wav.json").write_text(json.dumps({
            "max_abs": err, "snr_db": snr, "audio_seconds": x.shape[0] / sr,
            "reference_x_realtime_1_thread_512_sample_calls": x.shape[0] / sr / t_ref,
            "cuda_x_realtime_512_sample_calls": x.shape[0] / sr / t_gpu}, indent=1))


@pytest.mark.parametrize("two_d", [True, False], ids=["2d", "1d"])
def test_c_abi_tiled_call_matches_reference_and_monolithic(gpu, ref, two_d):
    """N3 as a product path: specbleach_b200_process_tiled (C ABI) on one 60 s stream, 6 tiles over
    the engine's lanes.  (a) against the REFERENCE build, whole output, BASELINE tolerance;
    (b) bit-identical to the monolithic specbleach[_2d]_process; (c) the handle continues the
    stream afterwards exactly as the monolithic handle does (state of the last tile adopted);
    (d) odd call sizes (carry != 0 at the call) and pageable buffers; (e) refusals."""
    sr = 48000
    prm = P2D if two_d else P1D
    x = synth(64.0, sr, 88)
    head, body, rest = x[:sr + 777], x[sr + 777:61 * sr + 123], x[61 * sr + 123:]   # odd cuts: carry != 0
    r = ref.RefDenoiser(sr, 50.0, two_d)
    r.load_parameters(learn_noise=1)
    want = [r.process(head)]
    r.load_parameters(**prm)
    want += [r.process(body, 1 << 18), r.process(rest)]
    want = np.concatenate(want)

    def run(tiled):
        d = gpu.Denoiser(sr, 50.0, two_d)
        d.load_parameters(learn_noise=1)
        ys = [d.process(head)]
        d.load_parameters(**prm)
        ys.append(d.process_tiled(body, tiles=6, warmup_frames=512) if tiled else d.process(body))
        ys.append(d.process(rest, 5000))
        return np.concatenate(ys)

    mono, tiled = run(False), run(True)
    err, snr = assert_parity(tiled, want, f"tiled C-ABI call vs reference, two_d={two_d}")
    same = np.array_equal(tiled, mono)
    print(f"tiled vs reference: {err:.3e} / {snr:.1f} dB; vs monolithic CUDA: "
          f"{'bit-identical' if same else float(np.max(np.abs(tiled - mono)))}")
    assert same
    d = gpu.Denoiser(sr, 50.0, two_d)
    d.load_parameters(learn_noise=1)
    with pytest.raises(RuntimeError):
        d.process_tiled(body, tiles=4)                                   # learn mode
    d.load_parameters(**dict(prm, adaptive_noise=1, noise_estimation_method=0))
    with pytest.raises(RuntimeError):
        d.process_tiled(body, tiles=4)                                   # adaptive needs the flag
    buf = body.copy()
    d.load_parameters(**prm)
    with pytest.raises(RuntimeError):
        d.process_tiled(buf, tiles=4, out=buf)                           # in place


def test_c_abi_tiled_call_adaptive_is_an_approximation(gpu, ref):
    """adaptive SPP-MMSE / Martin through time tiles: allowed by flag, and the error against the
    reference is MEASURED (printed and recorded), not assumed: the estimators' long memories make
    the tiles approximate."""
    from libspecbleach_b200.synth import nonstationary_noise
    import json
    from pathlib import Path
    sr = 44100
    x = nonstationary_noise(60.0, sr, 5)
    res = {}
    for method, name in ((0, "spp_mmse"), (2, "martin")):
        prm = dict(P2D, adaptive_noise=1, noise_estimation_method=method)
        want = run_flow(ref.RefDenoiser(sr, 50.0, True), x, 0, prm, 1 << 18)
        d = gpu.Denoiser2D(sr, 50.0)
        d.load_parameters(**prm)
        got = d.process_tiled(x, tiles=4, warmup_frames=1024, allow_adaptive=True)
        dd = got.astype(np.float64) - want
        err = float(np.max(np.abs(dd)))
        snr = 10 * np.log10(float(np.sum(want.astype(np.float64) ** 2)) / max(1e-300, float(np.sum(dd * dd))))
        res[name] = {"max_abs": err, "snr_db": snr}
        print(f"tiled adaptive {name}: max_abs={err:.3e} snr={snr:.1f} dB (4 tiles, 1024 warm-up frames)")
        assert np.all(np.isfinite(got)) and snr > 20.0
    out = Path(__file__).resolve().parent.parent / "gpurun_out"
    if out.is_dir():
        (out / "tiled_adaptive_error.json").write_text(json.dumps(res, indent=1))


def test_c_abi_tiled_batch_of_two_channels(gpu):
    """specbleach_b200_process_tiled_batch: both channels of a file in one call (all tiles enqueued
    before anything is waited for) == one tiled call per channel == the monolithic calls."""
    from libspecbleach_b200 import api
    sr = 48000
    xs = [synth(40.0, sr, 90 + c) for c in range(2)]

    def start():
        ds = [gpu.Denoiser2D(sr, 50.0) for _ in xs]
        heads = []
        for d, x in zip(ds, xs):
            d.load_parameters(learn_noise=1)
            heads.append(d.process(x[:sr]))
            d.load_parameters(**P2D)
        return ds, heads

    ds, _ = start()
    mono = [d.process(x[sr:]) for d, x in zip(ds, xs)]
    ds, _ = start()
    outs = [np.empty_like(x[sr:]) for x in xs]
    assert api.process_tiled_batch(ds, [np.ascontiguousarray(x[sr:]) for x in xs], outs, tiles=4)
    for a, b in zip(mono, outs):
        assert np.array_equal(a, b)
    # the handles go on from the last tiles' states
    tail = synth(2.0, sr, 99)
    ds2, _ = start()
    for d, x in zip(ds2, xs):
        d.process(x[sr:])
    for d, e in zip(ds, ds2):
        assert np.array_equal(d.process(tail), e.process(tail))


def test_c_abi_tiled_call_across_two_devices(gpu):
    """tiles dealt to the lanes of two GPUs of one process: bit-identical to one GPU, and the
    stream continues from the last tile's state even when that tile ran on the other GPU (the
    state travels as a snapshot blob)."""
    if gpu.device_count() < 2:
        pytest.skip("needs two GPUs")
    sr = 48000
    x = synth(60.0, sr, 89)
    tail = synth(2.0, sr, 98)

    def run(devices, tiles):
        d = gpu.Denoiser2D(sr, 50.0)
        d.load_parameters(learn_noise=1)
        a = d.process(x[:sr])
        d.load_parameters(**P2D)
        b = d.process_tiled(x[sr:], tiles=tiles, warmup_frames=512, devices=devices)
        return np.concatenate([a, b, d.process(tail)])

    want = run(None, 8)
    assert np.array_equal(want, run([0, 1], 8))        # last tile (index 7) on device 1
    assert np.array_equal(want, run([0, 1], 7))        # last tile (index 6) on device 0
