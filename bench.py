#!/usr/bin/env python
"""bench.py -- headline benchmark of specbleach-b200 (contract: see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU arm (oracle/_ref)

Workload = BASELINE.json configs[1]: 2D (NLM) denoiser, 48 kHz stereo, 10 min of
synthetic speech-plus-noise per GPU, default 50 ms frames / overlap 4.  One *step*
is the complete job on that file: reset profile -> learn the first second ->
denoise the remaining 599 s, both channels.  Metric: audio-seconds per second
(x realtime) of STEREO audio, whole job over all GPUs (weak scaling: one file per GPU).

  value : inputs already resident in HBM (specbleach_b200_process_batch_device)
  e2e   : reference-facing host API with pinned HOST buffers
          (specbleach_b200_process_batch = specbleach_2d_process per channel),
          host->device and device->host copies inside the timed region.

Prints exactly one JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SR = 48000
FRAME_MS = 50.0
FILE_SECONDS = 600.0
LEARN_SECONDS = 1.0
CHANNELS = 2
METRIC = "audio-seconds/sec per B200 (x realtime), 48kHz stereo 2D NLM"
PARAMS = dict(learn_noise=0, reduction_amount=20.0, smoothing_factor=1.0, whitening_factor=50.0,
              nlm_masking_protection=0.5, suppression_strength=50.0, aggressiveness=0.0,
              tonal_reduction=0.0)
# algorithmic work of the dominant kernel: direct-form NLM,
# ceil(K/8) paste blocks x 357 candidates x 218 FLOP per target frame
NLM_FLOP_PER_FRAME_BLOCK = 357 * 218


def config_dict(n_gpus: int) -> dict:
    return {
        "workload": "configs[1]: 2D NLM denoiser, 10 min synthetic 48 kHz stereo speech+noise, "
                    "50 ms frames, overlap 4 (one file per GPU)",
        "sample_rate": SR, "frame_ms": FRAME_MS, "fft_size": 3200, "bins": 1601,
        "channels": CHANNELS, "seconds": FILE_SECONDS, "learn_seconds": LEARN_SECONDS,
        "params": PARAMS, "files": n_gpus,
        "l2": "inputs (230 MB per step) larger than the 126 MB L2; no explicit flush",
    }


def pin_to_gpu_cpus(local: int) -> str | None:
    """N > 1 only: bind this rank to the CPUs NVML names as local to its GPU before the pinned
    host buffers are allocated, so that the H2D / D2H of eight ranks do not cross sockets.
    Best effort: returns a description, or None when NVML or the affinity call is unavailable."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return f"rank bound to the {len(cpus)} CPUs local to its GPU (NVML)"
    except Exception:  # noqa: BLE001
        return None


def make_file(rank: int) -> list[np.ndarray]:
    from libspecbleach_b200.synth import speech_plus_noise
    return [speech_plus_noise(FILE_SECONDS, SR, 20260002 + 1000 * rank + ch) for ch in range(CHANNELS)]


# ---------------------------------------------------------------------------
# clocks sampling during the timed region
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [a.strip() for a in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
# reference (CPU) arm: oracle/_ref on the host cores
# ---------------------------------------------------------------------------
def _ref_worker(seed, seconds, barrier, q):
    os.environ["OMP_NUM_THREADS"] = "1"      # read at nlm_filter_initialize (nlm_filter.c:127-136)
    from oracle import ref_binding as rb
    from libspecbleach_b200.synth import speech_plus_noise
    x = speech_plus_noise(seconds, SR, seed)
    d = rb.RefDenoiser(SR, FRAME_MS, True)
    barrier.wait()               # all workers enter the timed region together
    t0 = time.perf_counter()
    d.load_parameters(learn_noise=1)
    d.process(x[:int(LEARN_SECONDS * SR)])
    d.load_parameters(**PARAMS)
    d.process(x[int(LEARN_SECONDS * SR):])
    dt = time.perf_counter() - t0
    d.close()
    q.put(dt)


def _ref_single_stream(seconds, q):
    os.environ["OMP_NUM_THREADS"] = "4"      # the library's default NLM thread count
    from oracle import ref_binding as rb
    from libspecbleach_b200.synth import speech_plus_noise
    x = speech_plus_noise(seconds, SR, 9001)
    d = rb.RefDenoiser(SR, FRAME_MS, True)
    t0 = time.perf_counter()
    d.load_parameters(learn_noise=1)
    d.process(x[:int(LEARN_SECONDS * SR)])
    d.load_parameters(**PARAMS)
    d.process(x[int(LEARN_SECONDS * SR):])
    q.put(time.perf_counter() - t0)


def cpu_reference_single_stream(sample_seconds: float):
    """ONE mono stream with the library's default 4 NLM threads, x realtime (mono)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    p = ctx.Process(target=_ref_single_stream, args=(sample_seconds, q))
    p.start()
    dt = q.get(timeout=600)
    p.join()
    return sample_seconds / dt


def cpu_reference_throughput(sample_seconds: float, steps: int, warmup: int, workers: int | None = None):
    """x realtime (stereo) of the reference build with every host thread busy:
    `workers` independent mono streams in parallel processes, OMP_NUM_THREADS=1 each."""
    import multiprocessing as mp
    from oracle import ref_binding as rb
    if not rb.available():
        raise FileNotFoundError("oracle/_ref/libspecbleach_ref.so missing")
    workers = workers or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    times = []
    for s in range(warmup + steps):
        # one process per host thread; step time = slowest worker's timed region
        barrier = ctx.Barrier(workers)
        q = ctx.Queue()
        procs = [ctx.Process(target=_ref_worker, args=(9000 + 17 * w + s, sample_seconds, barrier, q))
                 for w in range(workers)]
        for p in procs:
            p.start()
        dts = [q.get(timeout=1800) for _ in procs]
        for p in procs:
            p.join()
        if s >= warmup:
            times.append(max(dts))
    wall = float(np.sum(times))
    stereo_seconds = steps * workers * sample_seconds / CHANNELS
    flags = (ROOT / "oracle" / "_ref" / "BUILD_FLAGS.txt")
    return {
        "value": stereo_seconds / wall, "unit": "x realtime (stereo audio-s/s)", "cores": workers,
        "kind": "reference",
        "sample": f"{workers} parallel mono streams x {sample_seconds:.0f} s (learn 1 s + denoise), "
                  f"{steps} step(s); reference sources built by oracle/Makefile with the FFTW shim",
        "build": flags.read_text().strip().replace("\n", " | ") if flags.exists() else None,
        "ms_per_step": 1e3 * wall / steps,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_throughput(args.cpu_sample_seconds, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "x realtime",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.gpus),
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "build")},
        "e2e": {"value": r["value"], "unit": "x realtime", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
def run_ours(args):
    import torch
    import libspecbleach_b200 as sb
    from libspecbleach_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # CPU baseline first (forks worker processes: do it before CUDA is initialised)
    cpu_pre = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = cpu_reference_throughput(args.cpu_sample_seconds, 1, 0)
            cpu_pre = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "build")}
            cpu_pre["single_stream_4_omp_threads_x_realtime_mono"] = cpu_reference_single_stream(
                args.cpu_sample_seconds)
        except Exception as e:  # noqa: BLE001
            cpu_pre = {"value": None, "unit": "x realtime", "cores": 0, "kind": "reference",
                       "sample": f"unavailable: {e!r}"}
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if not sb.set_device(local):
        raise SystemExit("bench.py: " + sb.last_error())
    dist = None
    affinity = None
    if world > 1:
        affinity = pin_to_gpu_cpus(local)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = api.lib()
    if args.lanes or args.chunk_frames:
        if not L.specbleach_b200_configure(args.chunk_frames or 0, args.lanes or 4):
            raise SystemExit("bench.py: specbleach_b200_configure failed: " + sb.last_error())

    chans = make_file(rank)
    n = chans[0].shape[0]
    n_learn = int(LEARN_SECONDS * SR)
    den = [sb.Denoiser2D(SR, FRAME_MS) for _ in range(CHANNELS)]

    # device-resident copies (value) and pinned host buffers (e2e)
    x_dev = [torch.from_numpy(c).cuda() for c in chans]
    y_dev = [torch.empty_like(t) for t in x_dev]
    x_pin = [api.PinnedArray(n) for _ in range(CHANNELS)]
    y_pin = [api.PinnedArray(n) for _ in range(CHANNELS)]
    for p, c in zip(x_pin, chans):
        p.array[:] = c

    def step(device: bool, serial: bool = False):
        if serial:
            # one channel after the other: only one lane is busy, so the CUDA events around a
            # launch time that launch alone (roofline pass)
            for c in range(CHANNELS):
                den[c].reset_profile()
                den[c].load_parameters(learn_noise=1)
                if not sb.process_batch([den[c]], [(x_dev[c].data_ptr(), n_learn)],
                                        [(y_dev[c].data_ptr(), n_learn)], device=True):
                    raise RuntimeError(sb.last_error())
                den[c].load_parameters(**PARAMS)
                if not sb.process_batch([den[c]], [(x_dev[c].data_ptr() + 4 * n_learn, n - n_learn)],
                                        [(y_dev[c].data_ptr() + 4 * n_learn, n - n_learn)], device=True):
                    raise RuntimeError(sb.last_error())
            return
        for d in den:
            d.reset_profile()
            d.load_parameters(learn_noise=1)
        if device:
            ins = [(t.data_ptr(), n_learn) for t in x_dev]
            outs = [(t.data_ptr(), n_learn) for t in y_dev]
        else:
            ins = [p.array[:n_learn] for p in x_pin]
            outs = [p.array[:n_learn] for p in y_pin]
        if not sb.process_batch(den, ins, outs, device=device):
            raise RuntimeError(sb.last_error())
        for d in den:
            d.load_parameters(**PARAMS)
        if device:
            ins = [(t.data_ptr() + 4 * n_learn, n - n_learn) for t in x_dev]
            outs = [(t.data_ptr() + 4 * n_learn, n - n_learn) for t in y_dev]
        else:
            ins = [p.array[n_learn:] for p in x_pin]
            outs = [p.array[n_learn:] for p in y_pin]
        if not sb.process_batch(den, ins, outs, device=device):
            raise RuntimeError(sb.last_error())

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(device: bool, steps: int, warmup: int, profile: bool):
        for _ in range(warmup):
            step(device, serial=profile)
        barrier()
        L.specbleach_b200_profile(1 if profile else 0)
        L.specbleach_b200_stats_reset()
        sampler = ClockSampler(local)
        sampler.start()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step(device, serial=profile)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop()
        st = api.stats()
        L.specbleach_b200_profile(0)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, clocks, st

    ms_dev, clocks, st_cnt = timed(True, args.steps, args.warmup, False)
    ms_e2e, clocks_e2e, st_e2e = timed(False, args.steps, max(1, args.warmup // 2), False)
    # roofline pass: the same steps again, one channel at a time, with CUDA events around every
    # launch on its lane stream (its wall time is NOT the reported value)
    ms_prof, _, st_dev = timed(True, args.steps, 0, True)
    classes = api.class_stats()

    # sanity of what was just computed (parity proper is tests/test_gpu_parity.py)
    out_rms = [float(np.sqrt(np.mean(np.square(p.array, dtype=np.float64)))) for p in y_pin]
    finite = all(bool(np.isfinite(p.array).all()) for p in y_pin)

    audio_s = FILE_SECONDS * world * args.steps
    value = audio_s / (ms_dev / 1e3)
    e2e = audio_s / (ms_e2e / 1e3)

    if rank == 0:
        fp32_peak = float(L.specbleach_b200_debug_fp32_peak_tflops())
        peaks = {}
        pf = ROOT / "MEASURED_PEAKS.json"
        if pf.exists():
            peaks = json.loads(pf.read_text())
        nlm_flop = st_dev["nlm_frame_blocks"] * NLM_FLOP_PER_FRAME_BLOCK
        nlm_s = st_dev["nlm_ms"] / 1e3
        achieved = nlm_flop / nlm_s / 1e12 if nlm_s > 0 else 0.0
        traffic = None
        executed = None
        tf = ROOT / "profiles" / "nlm_traffic.json"
        if tf.exists():
            try:
                tfj = json.loads(tf.read_text())
                ex = tfj.get("distance_kernel_executed")
                if ex:   # what the FP32 pipe really does (ncu SASS counts of the same capture)
                    k = tfj["kernels"]["nlm3_kernel"]
                    tflops = ex["fp32_flop_per_chunk"] / (k["time_us"] * 1e-6) / 1e12
                    executed = {"tflops_in_distance_kernel": tflops,
                                "frac_of_peak": tflops / fp32_peak if fp32_peak > 0 else None,
                                "issue_active_pct": ex["issue_active_pct"],
                                "fma_pipe_active_pct": ex["fma_pipe_active_pct"],
                                "source": ex["source"]}
                per_target = tfj.get("dram_bytes_per_target_frame")
                blocks_per_frame = (1601 + 7) // 8
                targets_per_launch = st_dev["nlm_frame_blocks"] / max(1, st_dev["nlm_launches"]) / blocks_per_frame
                traffic = per_target * targets_per_launch
            except Exception:
                traffic = None
        roofline = {
            "kernel": "NLM: sb::nlm3_kernel (row-walk distances) + sb::nlm3_paste_kernel",
            "bound": "fp32",
            "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
            "frac": achieved / fp32_peak if fp32_peak > 0 else None,
            "traffic": traffic,
            "executed": executed,
            "peak_source": "FFMA micro-benchmark run in this process (MEASURED_PEAKS.json has no "
                           "FP32 figure; HBM %.0f GB/s measured is not the bound: ~60 KB/frame)"
                           % peaks.get("hbm_gbs", 0.0),
            "flop_per_launch": nlm_flop / max(1, st_dev["nlm_launches"]),
            "avg_launch_ms": st_dev["nlm_ms"] / max(1, st_dev["nlm_launches"]),
            "launches": st_dev["nlm_launches"],
            "share_of_step": st_dev["nlm_ms"] / ms_prof if ms_prof > 0 else None,
            "note": "achieved = direct-form algorithmic FLOP (SURVEY 8d: 357 x 218 per target "
                    "frame and paste block) / event-timed NLM time; the kernel shares row SSDs "
                    "between 8 targets (~19 instead of 128 FP instructions per candidate), so "
                    "the algorithmic rate may exceed the FP32 pipe peak",
            "device_ms_by_kernel_class_per_step": {k: v[0] / args.steps for k, v in
                                                   sorted(classes.items(), key=lambda kv: -kv[1][0])},
        }
        cpu = cpu_pre
        bytes_step = CHANNELS * n * 4
        line = {
            "metric": METRIC, "value": value, "unit": "x realtime", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": dict(config_dict(world), host_affinity=affinity),
            "e2e": {"value": e2e, "unit": "x realtime", "h2d_bytes_per_step": bytes_step,
                    "d2h_bytes_per_step": bytes_step, "ms_per_step": ms_e2e / args.steps,
                    "api": "specbleach_b200_process_batch (= specbleach_2d_process per channel), "
                           "pinned host buffers"},
            "gpu_launches": int(st_cnt["kernel_launches"]),
            "clocks": clocks, "clocks_e2e": clocks_e2e,
            "roofline": roofline, "cpu_baseline": cpu,
            "frames_per_step": int(st_cnt["frames"] // max(1, args.steps)),
            "output_rms": out_rms, "output_finite": finite,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample-seconds", type=float, default=11.0,
                    help="audio seconds per CPU worker per step in the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=0, help="engine lanes (0 = library default)")
    ap.add_argument("--chunk-frames", type=int, default=0, help="frames per sub-call (0 = default)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
