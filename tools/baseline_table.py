"""BASELINE.md section 3 table from the committed bench lines (profiles/r02_bench_config*.json,
profiles/r02_scale_*.json).  usage: python tools/baseline_table.py"""
import json
from pathlib import Path

P = Path(__file__).resolve().parent.parent / "profiles"


def line(path):
    try:
        return json.loads(path.read_text().strip().splitlines()[-1])
    except Exception:
        return None


def f(v):
    return "–" if v is None else f"{v:,.0f}".replace(",", " ")


rows = []
names = {2: "2. 2D NLM, 10 min 48 kHz stereo", 3: "3. 4096 x 60 s mono, 2D NLM",
         4: "4. adaptive SPP-MMSE + Martin, 1 h 44.1 kHz", 5: "5. fft 8192, 96 kHz, 8 ch, tonal + veto + whitening"}
print("| Config | reference CPU x realtime (cores) | one stream, 4 OMP threads (mono) | B200 value | e2e (pinned) | "
      "e2e_dropin (22 symbols, pageable) | value / CPU | 8 GPUs value / e2e |")
print("|---|---|---|---|---|---|---|---|")
for c in (2, 3, 4, 5):
    l = line(P / f"r02_bench_config{c}.json")
    if not l:
        continue
    cb = l.get("cpu_baseline") or {}
    s8 = line(P / f"r02_scale_config{c}_8gpu.json")
    s8t = f"{f(s8['value'])} / {f(s8['e2e']['value'])}" if s8 else "–"
    ratio = l["value"] / cb["value"] if cb.get("value") else None
    print(f"| {names[c]} | {f(cb.get('value'))} ({cb.get('cores')}) | "
          f"{f(cb.get('single_stream_4_omp_threads_x_realtime_mono'))} | {f(l['value'])} | {f(l['e2e']['value'])} | "
          f"{f(l['e2e_dropin']['value'])} | {f(ratio)}x | {s8t} |")
