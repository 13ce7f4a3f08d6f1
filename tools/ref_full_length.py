"""Wall time of the reference CPU build on ONE WHOLE channel of the configs[1] input (600 s,
48 kHz mono, learn 1 s + denoise 599 s) with 1 and with 4 OpenMP threads -- the full-length run
that validates the bounded-sample extrapolation of bench.py's cpu_baseline.
TEST / MEASUREMENT INFRASTRUCTURE: drives oracle/_ref only.

    python tools/ref_full_length.py [seconds]      # prints one JSON line
"""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def run(threads: int, seconds: float):
    import multiprocessing as mp

    def work(q):
        os.environ["OMP_NUM_THREADS"] = str(threads)
        from oracle import ref_binding as rb
        from libspecbleach_b200.synth import speech_plus_noise
        import bench
        x = speech_plus_noise(seconds, 48000, 20260002)
        d = rb.RefDenoiser(48000, 50.0, True)
        t0 = time.perf_counter()
        d.load_parameters(learn_noise=1)
        d.process(x[:48000])
        d.load_parameters(**bench.P2D)
        d.process(x[48000:])
        q.put(time.perf_counter() - t0)

    ctx = mp.get_context("fork")
    q = ctx.Queue()
    p = ctx.Process(target=work, args=(q,))
    p.start()
    dt = q.get(timeout=7200)
    p.join()
    return dt


if __name__ == "__main__":
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
    out = {"audio_seconds": seconds, "host_threads": os.cpu_count()}
    for th in (1, 4):
        dt = run(th, seconds)
        out[f"wall_s_{th}_omp_threads"] = dt
        out[f"x_realtime_mono_{th}_omp_threads"] = seconds / dt
    print(json.dumps(out), flush=True)
