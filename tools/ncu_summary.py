"""Condense `ncu --page raw --csv` / `--page source --csv` exports into the handful of
numbers DESIGN.md and profiles/ quote.  usage: ncu_summary.py raw.csv [src.csv]
       ncu_summary.py --launches launches.csv '<the ncu command line>'"""
import collections
import csv
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "lts__t_sector_hit_rate.pct"]


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")][:90])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:80s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
        for i, k in enumerate(hdr):
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
                v = float(r[i].replace(",", ""))
                if v >= 0.15:
                    print(f"  stall {k.split('issue_stalled_')[1].split('_per_')[0]:40s} {v:8.2f}")


def src(path, top=14):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    iS, iI, iT, iSm = (hdr.index(x) for x in ("Source", "Instructions Executed",
                                               "Thread Instructions Executed", "# Samples"))
    ops, thr, smp = collections.Counter(), collections.Counter(), collections.Counter()
    tot = 0
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break                                 # export holds several kernels: first one only
        if len(r) <= iI or not r[iI].strip().isdigit():
            continue
        parts = r[iS].split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        op = op.split(".")[0]
        n = int(r[iI])
        ops[op] += n
        thr[op] += int(r[iT])
        smp[op] += int(r[iSm])
        tot += n
    ts = sum(smp.values())
    print(f"SASS lines {len(rows) - 2}, warp instructions executed {tot}")
    for op, n in ops.most_common(top):
        print(f"  {op:10s} {100 * n / tot:5.1f}% of instr  {100 * smp[op] / max(ts, 1):5.1f}% of samples  "
              f"avg active threads {thr[op] / max(n, 1):4.1f}")




def by_line(path, top=30):
    """`ncu --page source --csv --print-source cuda,sass` export: instructions and stall
    samples aggregated per CUDA source line."""
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
    hdr = rows[hi]
    iI, iSm = hdr.index("Instructions Executed"), hdr.index("# Samples")
    agg, src = collections.defaultdict(lambda: [0, 0]), {}
    cur = None
    for r in rows[hi + 1:]:
        if len(r) <= iI:
            continue
        if r[0].strip().isdigit():
            cur = int(r[0])
            src[cur] = r[1]
        elif r[0].strip():
            cur = None                            # header of another file's section
            continue
        if r[2].strip() and cur is not None:      # a SASS row belonging to line `cur`
            agg[cur][0] += int(r[iI]) if r[iI].strip().isdigit() else 0
            agg[cur][1] += int(r[iSm]) if r[iSm].strip().isdigit() else 0
    tot = sum(v[0] for v in agg.values())
    ts = sum(v[1] for v in agg.values())
    print(f"total warp instructions {tot}, samples {ts}")
    for ln, (n, sm) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"  {ln:4d} {100 * n / tot:5.1f}% instr {100 * sm / max(ts, 1):5.1f}% samples | {src[ln].strip()[:88]}")


def launches(path, title=""):
    """`ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel and its
    share of the captured launches (cold-cache, serialised: shares, not absolutes)."""
    import re
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[12] == "gpu__time_duration.sum"]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = re.sub(r"^.*::", "", r[4].split("(")[0])
        agg[name][0] += 1
        agg[name][1] += float(r[14]) / 1e3
    tot = sum(v[1] for v in agg.values())
    print(title)
    print(f"(cold-cache, serialised launches: compare SHARES)  launches {len(rows)}, total {tot:.0f} us")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:26s} n={n:4d} total {us:9.1f} us {100 * us / tot:5.1f}%  avg {us / n:8.1f} us")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
    else:
        raw(sys.argv[1])
        if len(sys.argv) > 2:
            src(sys.argv[2])
