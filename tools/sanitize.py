"""Small pass over every kernel family of the library, about one second of audio per case.

    SB_GUARD=1 python tools/sanitize.py        # guard zones of 0xFF bytes around every device buffer
    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize.py    # where allowed

The script only exercises the product path (no oracle).  It prints one checksum line per case, so
a guarded run can be compared with a plain one: an out-of-bounds or uninitialised float load
under SB_GUARD reads NaNs (0xFF bytes) and changes a checksum or trips the finiteness assert; an
out-of-bounds store shows in the last line's count (specbleach_b200_debug_check_guards).
tests/test_gpu_guard.py runs both and compares."""
import sys
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import libspecbleach_b200 as sb  # noqa: E402
from libspecbleach_b200 import api  # noqa: E402
from libspecbleach_b200.synth import speech_plus_noise  # noqa: E402

DEN = dict(learn_noise=0, reduction_amount=20.0, smoothing_factor=1.25, whitening_factor=50.0,
           masking_depth=0.5, suppression_strength=50.0, aggressiveness=0.5, tonal_reduction=6.0)


def crc(y):
    assert np.isfinite(y).all()
    return f"{zlib.crc32(np.ascontiguousarray(y).tobytes()):08x}"


def learned(cls, sr, ms, x, learn):
    d = cls(sr, ms)
    d.load_parameters(learn_noise=1)
    d.process(x[:learn])
    return d


def main():
    sr = 48000
    x = speech_plus_noise(1.6, sr, 11)
    learn = sr // 2

    d = learned(sb.Denoiser2D, sr, 50.0, x, learn)
    d.load_parameters(**DEN)
    print("2d manual          ", crc(d.process(x[learn:], chunk=7001)))
    d.load_parameters(**{**DEN, "residual_listen": True, "learn_noise": 2})
    print("2d residual/median ", crc(d.process(x[learn:learn + 9000])))
    blob = d.save_state()
    e = sb.Denoiser2D(sr, 50.0)
    assert e.load_state(blob)
    print("2d state round trip", crc(e.process(x[:6000])), crc(d.process(x[:6000])))
    e.close()

    for method in (0, 1, 2):
        a = sb.Denoiser2D(sr, 50.0)
        a.load_parameters(**{**DEN, "adaptive_noise": 1, "noise_estimation_method": method})
        n = 24000 if method == 1 else 48000
        print(f"2d adaptive {method}      ", crc(a.process(x[:n], chunk=16000)))
        a.close()

    o = learned(sb.Denoiser1D, 44100, 50.0, x, 22050)
    o.load_parameters(**DEN)
    print("1d manual          ", crc(o.process(x[22050:66150], chunk=512)))
    o.load_parameters(**{**DEN, "adaptive_noise": 1, "noise_estimation_method": 0})
    print("1d adaptive spp    ", crc(o.process(x[:30000])))
    o.close()

    for two_d, gain, sup, smooth in ((True, 1, 0, 1), (True, 2, 2, 1), (True, 0, 3, 1), (True, 0, 1, 1),
                                     (False, 0, 4, 1), (False, 0, 0, 2)):
        cls = sb.Denoiser2D if two_d else sb.Denoiser1D
        s = learned(cls, sr, 50.0, x, learn)
        assert s.set_strategy(gain, sup, smooth)
        s.load_parameters(**DEN)
        print(f"strategy {int(two_d)} {gain} {sup} {smooth}   ", crc(s.process(x[learn:learn + 30000])))
        s.close()

    hs = [learned(sb.Denoiser2D, sr, 50.0, x, learn) for _ in range(3)]
    for h in hs:
        h.load_parameters(**DEN)
    ins = [np.ascontiguousarray(x[learn + 100 * i:learn + 100 * i + 40000]) for i in range(3)]
    outs = [np.empty_like(a) for a in ins]
    assert api.process_batch(hs, ins, outs), api.last_error()
    print("batch of 3         ", " ".join(crc(y) for y in outs))
    y = hs[0].process_tiled(x[learn:], tiles=2, warmup_frames=30)
    print("tiled x2           ", crc(y))
    pcm = np.stack([x[:30000], x[100:30100]], axis=1)
    print("interleaved s16    ", crc(api.process_interleaved(hs[:2], (pcm * 20000).astype(np.int16), "s16")))
    for h in hs:
        h.close()

    big = speech_plus_noise(0.9, 96000, 5)
    b = learned(sb.Denoiser2D, 96000, 77.0, big, 30000)
    b.load_parameters(**{**DEN, "tonal_reduction": 12.0, "whitening_factor": 100.0})
    print("2d fft 8192        ", crc(b.process(big[30000:])))
    b.close()
    d.close()
    print("guard zones damaged:", api.lib().specbleach_b200_debug_check_guards())


if __name__ == "__main__":
    main()
