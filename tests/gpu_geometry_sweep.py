"""Diagnostic sweep (not a pytest): unusual (sample_rate, frame_ms) geometries against the
reference build, 1D and 2D.  python tests/gpu_geometry_sweep.py"""
import json
import sys
import traceback
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import libspecbleach_b200 as sb  # noqa: E402
from common import P1D, P2D, parity, run_flow  # noqa: E402
from libspecbleach_b200.synth import speech_plus_noise  # noqa: E402
from oracle import ref_binding as rb  # noqa: E402

CASES = [(8000, 32.0), (16000, 20.0), (22050, 46.0), (32000, 64.0), (44100, 23.2), (48000, 10.0),
         (48000, 100.0), (88200, 50.0), (192000, 20.0), (192000, 100.0), (11025, 92.9), (48000, 5.0),
         (48000, 250.0), (96000, 170.0)]
for sr, ms in CASES:
    for two_d in (True, False):
        rec = {"sr": sr, "ms": ms, "two_d": two_d}
        try:
            x = speech_plus_noise(3.0, sr, 7)
            r = rb.RefDenoiser(sr, ms, two_d)
            try:
                d = sb.Denoiser(sr, ms, two_d)
            except RuntimeError as e:
                rec["gpu_init"] = str(e)[:120]
                print(json.dumps(rec), flush=True)
                continue
            rec["geo"] = d.geometry()
            prm = P2D if two_d else P1D
            want = run_flow(r, x, sr, prm, 8192)
            got = run_flow(d, x, sr, prm, 7001)
            err, snr = parity(got, want)
            rec.update(max_abs=err, snr_db=float(snr), lat=(r.latency(), d.latency()),
                       size=(r.profile_size(), d.profile_size()))
            d.close()
            r.close()
        except Exception as e:  # noqa: BLE001
            rec["error"] = repr(e)[:200]
            rec["tb"] = traceback.format_exc()[-400:]
        print(json.dumps(rec), flush=True)
