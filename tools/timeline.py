"""Poor man's device timeline of one bench step with both lanes busy (no nsys in this image):
every launch of the library is bracketed by CUDA events on its stream (profile mode) and
SB_TIMELINE makes the engine dump them on one clock.  Prints, per stream, busy time per kernel
class, and the overlap structure: how long 0 / 1 / 2+ streams had a kernel in flight.

    SB_TIMELINE=/tmp/tl.csv python tools/timeline.py [--config 2]
"""
import argparse
import collections
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=600.0)
    args = ap.parse_args()
    path = os.environ.get("SB_TIMELINE")
    if not path:
        raise SystemExit("set SB_TIMELINE=<file>")
    import torch
    import libspecbleach_b200 as sb
    from libspecbleach_b200 import api
    from libspecbleach_b200.synth import speech_plus_noise
    import bench
    sb.set_device(0)
    L = api.lib()
    sr = 48000
    xs = [speech_plus_noise(args.seconds, sr, 20260002 + c) for c in range(2)]
    n = xs[0].shape[0]
    xd = [torch.from_numpy(x).cuda() for x in xs]
    yd = [torch.empty_like(t) for t in xd]
    den = [sb.Denoiser2D(sr, 50.0) for _ in range(2)]

    def step():
        for d in den:
            d.reset_profile()
            d.load_parameters(learn_noise=1)
        sb.process_batch(den, [(t.data_ptr(), sr) for t in xd], [(t.data_ptr(), sr) for t in yd], device=True)
        for d in den:
            d.load_parameters(**bench.P2D)
        sb.process_batch(den, [(t.data_ptr() + 4 * sr, n - sr) for t in xd],
                         [(t.data_ptr() + 4 * sr, n - sr) for t in yd], device=True)

    step()
    step()
    torch.cuda.synchronize()
    if os.path.exists(path):
        os.remove(path)
    L.specbleach_b200_profile(1)
    L.specbleach_b200_stats_reset()
    step()
    torch.cuda.synchronize()
    api.stats()
    L.specbleach_b200_profile(0)
    rows = []
    for ln in open(path):
        if ln.startswith("#"):
            continue
        c, st, a, b = ln.strip().split(",")
        rows.append((c, st, float(a), float(b)))
    t_end = max(r[3] for r in rows)
    print(f"{len(rows)} timed launches, span {t_end:.2f} ms")
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    for c, st, a, b in rows:
        per[st][c] += b - a
    for st, d in per.items():
        tot = sum(d.values())
        print(f"stream {st}: in-flight {tot:.2f} ms:", {k: round(v, 2) for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:10]})
    # overlap histogram
    ev = []
    for c, st, a, b in rows:
        ev.append((a, 1, c))
        ev.append((b, -1, c))
    ev.sort()
    depth, last, hist = 0, 0.0, collections.defaultdict(float)
    active = collections.Counter()
    nlm_alone = nlm_with = 0.0
    for t, d, c in ev:
        hist[depth] += t - last
        if active["nlm"] > 0:
            if depth > active["nlm"]:
                nlm_with += t - last
            else:
                nlm_alone += t - last
        last = t
        depth += d
        active[c] += d
    print("time with k launches in flight:", {k: round(v, 2) for k, v in sorted(hist.items())})
    print(f"NLM in flight alone {nlm_alone:.2f} ms, NLM together with another stream's launch {nlm_with:.2f} ms")


if __name__ == "__main__":
    main()
