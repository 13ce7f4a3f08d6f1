#!/usr/bin/env python
"""bench.py -- benchmarks of specbleach-b200 (contract: see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W [--config 2|3|4|5]    # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ... [--config]  # reference CPU arm (oracle/_ref)

Default workload = BASELINE.json configs[1] (the configuration the metric is quoted on): 2D (NLM)
denoiser, 48 kHz stereo, 10 min of synthetic speech-plus-noise per GPU, 50 ms frames / overlap 4.
One *step* is the complete job: reset profile -> learn the first second -> denoise the rest, every
channel / stream of the workload.  Metric: audio-seconds per second (x realtime), whole job over
all GPUs.  --config selects the other BASELINE configs:

  2  configs[1]  10 min 48 kHz stereo, 2D NLM, one file per GPU (weak scaling)          [default]
  3  configs[2]  4096 independent 60 s 48 kHz MONO streams, 2D NLM, sharded over the GPUs (strong)
  4  configs[3]  1 h of non-stationary noise at 44.1 kHz mono, adaptive SPP-MMSE and Martin
  5  configs[4]  96 kHz / 77 ms frames (fft 8192, 4097 bins), 8 channels x 120 s, tonal + veto + whitening

  value      : inputs already resident in HBM (specbleach_b200_process_batch_device)
  e2e        : host API with pinned HOST buffers (specbleach_b200_process_batch =
               specbleach[_2d]_process per stream), H2D and D2H copies inside the timed region
  e2e_dropin : the 22-symbol drop-in API only -- specbleach_2d_process per handle, one after the
               other, plain malloc'd (pageable) numpy buffers, as examples/denoiser_demo.c calls it

Prints exactly one JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

P2D = dict(learn_noise=0, reduction_amount=20.0, smoothing_factor=1.0, whitening_factor=50.0,
           nlm_masking_protection=0.5, suppression_strength=50.0, aggressiveness=0.0,
           tonal_reduction=0.0)
# algorithmic work of the dominant kernel: direct-form NLM,
# ceil(K/8) paste blocks x 357 candidates x 218 FLOP per target frame
NLM_FLOP_PER_FRAME_BLOCK = 357 * 218
METRIC = "audio-seconds/sec per B200 (x realtime), 48kHz stereo 2D NLM"


# ---------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------
class Workload:
    """One BASELINE config: which streams a rank owns, their parameters, how audio is counted."""

    def __init__(self, cid: int, world: int, rank: int):
        self.cid, self.world, self.rank = cid, world, rank
        self.learn_seconds = 1.0
        self.wave = 0                      # streams per process_batch call (0 = all of the rank's)
        if cid == 2:
            self.key, self.sr, self.frame_ms = "configs[1]", 48000, 50.0
            self.seconds, self.channels_per_unit = 600.0, 2
            self.n_streams, self.n_unique = 2, 2
            self.params = [P2D, P2D]
            self.metric = METRIC
            self.scaling = "weak"
            self.what = ("configs[1]: 2D NLM denoiser, 10 min synthetic 48 kHz stereo speech+noise, "
                         "50 ms frames, overlap 4 (one file per GPU)")
            self.counting = "a stereo file counts its duration once"
        elif cid == 3:
            from libspecbleach_b200.sharding import shard_bounds
            self.key, self.sr, self.frame_ms = "configs[2]", 48000, 50.0
            self.seconds, self.channels_per_unit = 60.0, 1
            lo, hi = shard_bounds(4096, world, rank)
            self.n_streams, self.n_unique, self.wave = hi - lo, 128, 128
            self.params = [P2D] * self.n_streams
            self.metric = "audio-seconds/sec (x realtime), 4096 x 60 s 48 kHz MONO streams, 2D NLM"
            self.scaling = "strong"
            self.what = ("configs[2]: batch of 4096 independent 60 s 48 kHz mono streams, 2D NLM, sharded "
                         "over the GPUs (static block partition, no data-path collective); each stream has "
                         "its own handle; waves of 128 streams through specbleach_b200_process_batch; the "
                         "128 distinct synthetic streams of a wave are replayed for every wave")
            self.counting = "MONO: every stream counts its duration"
        elif cid == 4:
            self.key, self.sr, self.frame_ms = "configs[3]", 44100, 50.0
            self.seconds, self.channels_per_unit = 3600.0, 1
            self.learn_seconds = 0.0
            self.n_streams, self.n_unique = 2, 1
            self.params = [dict(P2D, adaptive_noise=1, noise_estimation_method=0),
                           dict(P2D, adaptive_noise=1, noise_estimation_method=2)]
            self.metric = ("audio-seconds/sec (x realtime), 1 h 44.1 kHz mono non-stationary noise, 2D NLM with "
                           "adaptive noise estimation: SPP-MMSE and Martin, one pass each")
            self.scaling = "weak"
            self.what = ("configs[3]: adaptive estimation on 1 h of synthetic non-stationary noise at 44.1 kHz "
                         "(fft 3006, 1504 bins, 288 131 frames): the hour goes through the 2D denoiser twice, "
                         "once with SPP-MMSE and once with Martin (two handles, one batch); no learn phase")
            self.counting = "MONO: each of the two passes counts the hour"
        elif cid == 5:
            self.key, self.sr, self.frame_ms = "configs[4]", 96000, 77.0
            self.seconds, self.channels_per_unit = 120.0, 8
            self.n_streams, self.n_unique = 8, 8
            prm = dict(P2D, tonal_reduction=12.0, nlm_masking_protection=1.0, whitening_factor=100.0)
            self.params = [prm] * 8
            self.metric = ("audio-seconds/sec (x realtime), 96 kHz 8-channel, 77 ms frames (fft 8192), 2D NLM "
                           "with tonal reduction, masking veto and whitening")
            self.scaling = "weak"
            self.what = ("configs[4]: large-frame 2D denoising (96 kHz, 77 ms -> fft 8192, 4097 bins) of 120 s of "
                         "8-channel synthetic speech+noise+60 Hz hum, tonal reduction 12 dB, masking protection "
                         "1.0, whitening 100 %, learned profile (one file per GPU)")
            self.counting = "an 8-channel file counts its duration once"
        else:
            raise SystemExit(f"bench.py: unknown --config {cid}")
        self.n = int(round(self.seconds * self.sr))
        self.n_learn = int(round(self.learn_seconds * self.sr))
        if not self.wave:
            self.wave = self.n_streams
        frame = int(np.float32(self.frame_ms / 1000.0) * np.float32(self.sr))
        n_fft = frame + 800
        n_fft += n_fft & 1
        self.frame, self.fft, self.hop, self.bins = frame, n_fft, frame // 4, n_fft // 2 + 1

    # audio-seconds of the whole job (all ranks) per step, in the config's own accounting
    def job_audio_seconds(self) -> float:
        total_streams = {2: 2 * self.world, 3: 4096, 4: 2 * self.world, 5: 8 * self.world}[self.cid]
        return total_streams * self.seconds / self.channels_per_unit

    def unique_of(self, i: int) -> int:
        return 0 if self.cid == 4 else i % self.n_unique

    def make_unique_streams(self) -> list[np.ndarray]:
        from libspecbleach_b200.synth import nonstationary_noise, speech_plus_noise
        r = self.rank
        if self.cid == 2:
            return [speech_plus_noise(self.seconds, self.sr, 20260002 + 1000 * r + ch) for ch in range(2)]
        if self.cid == 3:
            base = [speech_plus_noise(self.seconds + 1.0, self.sr, 20260003 + 8 * r + i) for i in range(8)]
            out = []
            for i in range(self.n_unique):
                b = base[i % 8]
                off = (37 * i + 11) % self.sr
                # the first second stays noise only: rotate the speech part, keep the head
                seg = np.concatenate([b[:self.n_learn], b[self.n_learn + off:self.n_learn + off + self.n - self.n_learn]])
                if seg.shape[0] < self.n:
                    seg = np.concatenate([seg, b[self.n_learn:self.n_learn + self.n - seg.shape[0]]])
                out.append(np.ascontiguousarray(seg * np.float32(0.6 + 0.4 * ((i * 7) % 11) / 10.0)))
            return out
        if self.cid == 4:
            return [nonstationary_noise(self.seconds, self.sr, 20260004 + r)]
        return [speech_plus_noise(self.seconds, self.sr, 20260005 + 100 * r + c, hum_hz=60.0) for c in range(8)]

    def config_dict(self) -> dict:
        return {
            "workload": self.what, "baseline_config": self.key, "audio_accounting": self.counting,
            "sample_rate": self.sr, "frame_ms": self.frame_ms, "fft_size": self.fft, "bins": self.bins,
            "streams_per_gpu": self.n_streams if self.cid != 3 else f"4096 / {self.world}",
            "seconds_per_stream": self.seconds, "learn_seconds": self.learn_seconds,
            "params": self.params[0] if self.cid != 4 else self.params, "gpus": self.world,
            "l2": "inputs (%d MB per wave) larger than the 126 MB L2; no explicit flush"
                  % (self.wave * self.n * 4 // 1000000),
        }


def pin_to_gpu_cpus(local: int) -> str | None:
    """N > 1 only: bind this rank to the CPUs NVML names as local to its GPU before the pinned
    host buffers are allocated.  Best effort: returns a description, or None."""
    try:
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
def _ref_stream(cid: int, seconds: float, seed: int):
    from libspecbleach_b200.synth import nonstationary_noise, speech_plus_noise
    if cid == 4:
        return nonstationary_noise(seconds, 44100, seed)
    if cid == 5:
        return speech_plus_noise(seconds, 96000, seed, hum_hz=60.0)
    return speech_plus_noise(seconds, 48000, seed)


def _ref_run(d, x, w: Workload, prm: dict):
    n_learn = min(w.n_learn, x.shape[0])
    if n_learn:
        d.load_parameters(learn_noise=1)
        d.process(x[:n_learn])
    d.load_parameters(**prm)
    d.process(x[n_learn:])


def _ref_worker(cid, widx, seed, seconds, omp_threads, barrier, q):
    os.environ["OMP_NUM_THREADS"] = str(omp_threads)   # read at nlm_filter_initialize (nlm_filter.c:127-136)
    from oracle import ref_binding as rb
    w = Workload(cid, 1, 0)
    x = _ref_stream(cid, seconds, seed)
    d = rb.RefDenoiser(w.sr, w.frame_ms, True)
    prm = w.params[widx % len(w.params)]
    if barrier is not None:
        barrier.wait()           # all workers enter the timed region together
    t0 = time.perf_counter()
    _ref_run(d, x, w, prm)
    dt = time.perf_counter() - t0
    d.close()
    q.put(dt)


def cpu_reference_single_stream(cid: int, sample_seconds: float, omp_threads: int = 4):
    """ONE mono stream with the library's default 4 NLM threads, x realtime (mono)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    q = ctx.Queue()
    p = ctx.Process(target=_ref_worker, args=(cid, 0, 9001, sample_seconds, omp_threads, None, q))
    p.start()
    dt = q.get(timeout=1800)
    p.join()
    return sample_seconds / dt


def cpu_reference_throughput(cid: int, sample_seconds: float, steps: int, warmup: int,
                             workers: int | None = None):
    """x realtime (in the config's audio accounting) of the reference build with every host thread
    busy: `workers` independent mono streams in parallel processes, OMP_NUM_THREADS=1 each."""
    import multiprocessing as mp
    from oracle import ref_binding as rb
    if not rb.available():
        raise FileNotFoundError("oracle/_ref/libspecbleach_ref.so missing")
    w = Workload(cid, 1, 0)
    workers = workers or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    times = []
    for s in range(warmup + steps):
        # one process per host thread; step time = slowest worker's timed region
        barrier = ctx.Barrier(workers)
        q = ctx.Queue()
        procs = [ctx.Process(target=_ref_worker,
                             args=(cid, i, 9000 + 17 * i + s, sample_seconds, 1, barrier, q))
                 for i in range(workers)]
        for p in procs:
            p.start()
        dts = [q.get(timeout=3600) for _ in procs]
        for p in procs:
            p.join()
        if s >= warmup:
            times.append(max(dts))
    wall = float(np.sum(times))
    audio = steps * workers * sample_seconds / w.channels_per_unit
    flags = (ROOT / "oracle" / "_ref" / "BUILD_FLAGS.txt")
    return {
        "value": audio / wall, "unit": "x realtime", "cores": workers, "kind": "reference",
        "sample": f"{workers} parallel mono streams x {sample_seconds:.0f} s of the {w.key} workload "
                  f"({w.sr} Hz, {w.frame_ms} ms frames, learn {w.learn_seconds:.0f} s + denoise"
                  + (", SPP-MMSE and Martin alternating over the workers" if cid == 4 else "")
                  + f"), {steps} step(s), OMP_NUM_THREADS=1 each; {w.counting}; reference sources "
                  "built by oracle/Makefile with the double-precision FFTW shim",
        "build": flags.read_text().strip().replace("\n", " | ") if flags.exists() else None,
        "ms_per_step": 1e3 * wall / steps,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = Workload(args.config, max(1, args.gpus), 0)
    r = cpu_reference_throughput(args.config, args.cpu_sample_seconds, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": w.metric, "value": r["value"], "unit": "x realtime",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": w.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": w.config_dict(),
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "build")},
        "e2e": {"value": r["value"], "unit": "x realtime", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# per-kernel-class roofline (north_star: "each kernel reported as a fraction of its roofline")
# ---------------------------------------------------------------------------
def class_rooflines(classes: dict, frames: float, w: Workload, hbm_gbs: float, fp32_tflops: float,
                    sm_mhz: float, two_d: bool = True) -> dict:
    """classes: {name: (ms, launches)} from the event-timed pass over `frames` STFT frames.
    Streaming classes: algorithmic bytes per frame (compulsory reads + writes of that kernel,
    DESIGN.md section 3) / event time vs the measured HBM copy rate.  FFT: 2.5 N log2 N FLOP per
    transform vs the measured FFMA peak, plus its shared-memory traffic.  Bark band / NMR kernels:
    transcendental (MUFU) operations vs 16 per clock per SM.  Scans: latency-bound recurrences,
    reported as ns per frame."""
    K, N, hop, frame = w.bins, w.fft, w.hop, w.frame
    Kp = (K + 3) & ~3
    nfac_passes = max(1, round(math.log(N / 2, 4.5)))          # ~5 Stockham passes at M = 1600
    sms = 148
    clk = (sm_mhz or 1965.0) * 1e6
    smem_peak = sms * 128 * clk / 1e9                          # GB/s
    mufu_peak = sms * 16 * clk / 1e9                           # G ops/s
    fft_flop = 2.5 * N * math.log2(N)
    stream_bytes = {                                           # per frame
        "ola": 4 * frame + 4 * hop,
        "snr": 12 * Kp,
        "gain_pre": 20 * Kp,
        "bin_scan": 8 * Kp,
        "final": 32 * Kp + 128,
        "misc": 4 * Kp,
    }
    out = {}
    for name, (ms, launches) in classes.items():
        if ms <= 0:
            continue
        s = ms / 1e3
        e = {"ms": ms, "launches": launches}
        if name in ("forward_fft", "inverse_fft"):
            tf = fft_flop * frames / s / 1e12
            byt = (4 * hop + 12 * Kp) if name == "forward_fft" else (8 * Kp + 4 * frame)
            smem = 2 * 8 * (N / 2) * (nfac_passes + 1)         # every pass reads and writes M complex
            e.update(bound="smem/fp32", achieved=tf, peak=fp32_tflops, unit="TFLOP/s",
                     frac=tf / fp32_tflops if fp32_tflops else None,
                     hbm_gbs=byt * frames / s / 1e9, hbm_frac=byt * frames / s / 1e9 / hbm_gbs,
                     smem_gbs=smem * frames / s / 1e9, smem_frac=smem * frames / s / 1e9 / smem_peak)
        elif name in stream_bytes:
            g = stream_bytes[name] * frames / s / 1e9
            e.update(bound="hbm", achieved=g, peak=hbm_gbs, unit="GB/s", frac=g / hbm_gbs)
        elif name in ("band", "nmr"):
            ops = (2 * (25 * 25 * 2 + 25 * 4) + 2 * K) if name == "band" else (3 * K + 25 * 3)
            g = ops * frames / s / 1e9
            byt = (12 * Kp + 128) if name == "band" else (4 * Kp + 256)
            e.update(bound="mufu", achieved=g, peak=mufu_peak, unit="Gop/s", frac=g / mufu_peak,
                     hbm_gbs=byt * frames / s / 1e9, hbm_frac=byt * frames / s / 1e9 / hbm_gbs)
        elif name in ("tscan", "pscan", "adaptive"):
            e.update(bound="latency", ns_per_frame=1e6 * ms / max(1.0, frames))
        elif name in ("nlm", "nlm_tail"):
            nb = (K + 7) // 8 if name == "nlm" else 1
            tf = nb * NLM_FLOP_PER_FRAME_BLOCK * frames / s / 1e12
            e.update(bound="fp32", achieved=tf, peak=fp32_tflops, unit="TFLOP/s (direct-form algorithmic)",
                     frac=tf / fp32_tflops if fp32_tflops else None)
        else:
            # learn / finalize / tonal mask / whitening rows: per profile or per learn frame, a
            # handful of small launches per stream -- no per-frame roofline
            e.update(bound="n/a (per-profile rows)")
        out[name] = e
    return out


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
    w = Workload(args.config, world, rank)
    # CPU baseline first (forks worker processes: do it before CUDA is initialised)
    cpu_pre = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = cpu_reference_throughput(args.config, args.cpu_sample_seconds, 1, 0)
            cpu_pre = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "build")}
            cpu_pre["single_stream_4_omp_threads_x_realtime_mono"] = cpu_reference_single_stream(
                args.config, args.cpu_sample_seconds)
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

    uniq = w.make_unique_streams()
    n, n_learn = w.n, w.n_learn
    den = [sb.Denoiser2D(w.sr, w.frame_ms) for _ in range(w.n_streams)]
    n_waves = (w.n_streams + w.wave - 1) // w.wave

    # device-resident copies (value) and pinned host buffers (e2e): one wave's worth, reused by
    # every wave (config 3 replays its 128 distinct streams)
    slots = min(w.wave, w.n_streams)
    x_dev = [torch.from_numpy(u).cuda() for u in uniq]
    y_dev = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(slots)]
    x_pin = [api.PinnedArray(n) for _ in uniq]
    y_pin = [api.PinnedArray(n) for _ in range(slots)]
    for p, c in zip(x_pin, uniq):
        p.array[:] = c

    def run_wave(idx: list[int], device: bool):
        hs = [den[i] for i in idx]
        if device:
            src = [x_dev[w.unique_of(i)].data_ptr() for i in idx]
            dst = [y_dev[j].data_ptr() for j in range(len(idx))]
        if n_learn:
            for d in hs:
                d.reset_profile()
                d.load_parameters(learn_noise=1)
            if device:
                ins = [(p, n_learn) for p in src]
                outs = [(p, n_learn) for p in dst]
            else:
                ins = [x_pin[w.unique_of(i)].array[:n_learn] for i in idx]
                outs = [y_pin[j].array[:n_learn] for j in range(len(idx))]
            if not sb.process_batch(hs, ins, outs, device=device):
                raise RuntimeError(sb.last_error())
        for i, d in zip(idx, hs):
            d.load_parameters(**w.params[i])
        if device:
            ins = [(p + 4 * n_learn, n - n_learn) for p in src]
            outs = [(p + 4 * n_learn, n - n_learn) for p in dst]
        else:
            ins = [x_pin[w.unique_of(i)].array[n_learn:] for i in idx]
            outs = [y_pin[j].array[n_learn:] for j in range(len(idx))]
        if not sb.process_batch(hs, ins, outs, device=device):
            raise RuntimeError(sb.last_error())

    prof_streams = list(range(min(w.n_streams, 8)))

    def step(device: bool, serial: bool = False):
        if serial:
            # one stream after the other: only one lane is busy, so the CUDA events around a
            # launch time that launch alone (roofline pass)
            for i in prof_streams:
                run_wave([i], True)
            return
        for k in range(n_waves):
            run_wave(list(range(k * w.wave, min(w.n_streams, (k + 1) * w.wave))), device)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps: int, warmup: int, profile: bool = False):
        for _ in range(warmup):
            fn()
        barrier()
        L.specbleach_b200_profile(1 if profile else 0)
        L.specbleach_b200_stats_reset()
        sampler = ClockSampler(local)
        sampler.start()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
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

    ms_dev, clocks, st_cnt = timed(lambda: step(True), args.steps, args.warmup)
    ms_e2e, clocks_e2e, st_e2e = timed(lambda: step(False), args.steps, max(1, args.warmup // 2))

    # sanity of what was just computed (parity proper is tests/test_gpu_*.py)
    out_rms = [float(np.sqrt(np.mean(np.square(p.array, dtype=np.float64)))) for p in y_pin[:8]]
    finite = all(bool(np.isfinite(p.array).all()) for p in y_pin[:8])

    # drop-in e2e: the reference's own 22 symbols only, pageable buffers, one handle after the other
    drop_idx = list(range(min(w.n_streams, 16 if w.cid == 3 else w.n_streams)))
    x_page = [np.array(uniq[w.unique_of(i)], copy=True) for i in drop_idx]
    y_page = [np.empty(n, dtype=np.float32) for _ in drop_idx]

    def step_dropin():
        for j, i in enumerate(drop_idx):
            d = den[i]
            if n_learn:
                d.reset_profile()
                d.load_parameters(learn_noise=1)
                if not d.process_raw(n_learn, x_page[j].ctypes.data, y_page[j].ctypes.data):
                    raise RuntimeError(sb.last_error())
            d.load_parameters(**w.params[i])
            if not d.process_raw(n - n_learn, x_page[j][n_learn:].ctypes.data, y_page[j][n_learn:].ctypes.data):
                raise RuntimeError(sb.last_error())

    drop_steps = max(1, min(args.steps, 3))
    ms_drop, _, _ = timed(step_dropin, drop_steps, 1)
    drop_same = all(np.array_equal(y_page[j], y_pin[j].array) for j in range(min(len(drop_idx), slots))
                    ) if (n_waves == 1 and w.cid != 4) else None

    # roofline pass: the same work again, one stream at a time, with CUDA events around every
    # launch on its lane stream (its wall time is NOT the reported value)
    prof_steps = 1 if w.cid != 2 else args.steps
    ms_prof, _, st_dev = timed(lambda: step(True, serial=True), prof_steps, 0, profile=True)
    classes = api.class_stats()
    prof_frames = float(st_dev["frames"])
    step_frames = float(st_cnt["frames"]) / max(1, args.steps)
    prof_scale = step_frames / prof_frames * prof_steps if prof_frames else 0.0   # profile pass -> one full step

    audio_s = w.job_audio_seconds() * args.steps
    value = audio_s / (ms_dev / 1e3)
    e2e = audio_s / (ms_e2e / 1e3)
    drop_audio = len(drop_idx) * w.seconds / w.channels_per_unit * drop_steps * world
    e2e_dropin = drop_audio / (ms_drop / 1e3)

    if rank == 0:
        fp32_peak = float(L.specbleach_b200_debug_fp32_peak_tflops())
        peaks = {}
        pf = ROOT / "MEASURED_PEAKS.json"
        if pf.exists():
            peaks = json.loads(pf.read_text())
        hbm = float(peaks.get("hbm_gbs", 6552.3))
        nlm_flop = st_dev["nlm_frame_blocks"] * NLM_FLOP_PER_FRAME_BLOCK
        nlm_s = st_dev["nlm_ms"] / 1e3
        achieved = nlm_flop / nlm_s / 1e12 if nlm_s > 0 else 0.0
        traffic = None
        executed = None
        tf = ROOT / "profiles" / "nlm_traffic.json"
        if tf.exists() and w.bins == 1601:
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
                blocks_per_frame = (w.bins + 7) // 8
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
                           "FP32 figure; HBM %.0f GB/s measured is not the bound: ~60 KB/frame)" % hbm,
            "flop_per_launch": nlm_flop / max(1, st_dev["nlm_launches"]),
            "avg_launch_ms": st_dev["nlm_ms"] / max(1, st_dev["nlm_launches"]),
            "launches": st_dev["nlm_launches"],
            "share_of_step": st_dev["nlm_ms"] / ms_prof if ms_prof > 0 else None,
            "note": "achieved = direct-form algorithmic FLOP (SURVEY 8d: 357 x 218 per target "
                    "frame and paste block) / event-timed NLM time; the kernel shares row SSDs "
                    "between 8 targets (~19 instead of 128 FP instructions per candidate), so "
                    "the algorithmic rate may exceed the FP32 pipe peak",
            "device_ms_by_kernel_class_per_step": {k: v[0] * prof_scale / prof_steps for k, v in
                                                   sorted(classes.items(), key=lambda kv: -kv[1][0])},
        }
        by_class = class_rooflines(classes, prof_frames, w, hbm, fp32_peak, clocks.get("sm_mhz") or 1965.0)
        bytes_step = w.n_streams * n * 4
        line = {
            "metric": w.metric, "value": value, "unit": "x realtime", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": w.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": w.config_dict(),
            "e2e": {"value": e2e, "unit": "x realtime", "h2d_bytes_per_step": bytes_step,
                    "d2h_bytes_per_step": bytes_step, "ms_per_step": ms_e2e / args.steps,
                    "api": "specbleach_b200_process_batch (= specbleach_2d_process per stream), "
                           "pinned host buffers"},
            "e2e_dropin": {"value": e2e_dropin, "unit": "x realtime",
                           "ms_per_step": ms_drop / drop_steps, "streams": len(drop_idx),
                           "api": "specbleach_2d_process only (the reference's 22 symbols), one handle "
                                  "after the other, pageable numpy buffers",
                           "bit_identical_to_batch_output": drop_same},
            "gpu_launches": int(st_cnt["kernel_launches"]),
            "collective": "none on the data path (NCCL only for the barrier and the max-reduce of the timing)",
            "host_affinity": affinity,
            "clocks": clocks, "clocks_e2e": clocks_e2e,
            "roofline": roofline, "roofline_by_class": by_class, "cpu_baseline": cpu_pre,
            "frames_per_step": int(step_frames),
            "output_rms": out_rms, "output_finite": finite,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------
# N3: one file, time tiles over several GPUs of ONE process (strong scaling)
# ---------------------------------------------------------------------------
def run_tiled(args):
    import torch
    import libspecbleach_b200 as sb
    from libspecbleach_b200 import api
    if args.config != 2 or "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) > 1:
        raise SystemExit("bench.py --tiled: config 2, one process (no torchrun); --gpus N selects the devices")
    if not torch.cuda.is_available() or torch.cuda.device_count() < args.gpus:
        raise SystemExit("bench.py --tiled: not enough CUDA devices -- the product path has no CPU fallback")
    w = Workload(2, 1, 0)
    if args.file_seconds > 0:          # a longer file amortises the per-tile host work
        w.seconds = args.file_seconds
        w.n = int(round(w.seconds * w.sr))
        w.what = w.what.replace("10 min", f"{w.seconds / 60:.0f} min")
    devices = list(range(args.gpus))
    sb.set_device(0)
    torch.cuda.set_device(0)
    L = api.lib()
    lanes = args.lanes or 4
    if not L.specbleach_b200_configure(args.chunk_frames or 0, lanes):
        raise SystemExit("bench.py: specbleach_b200_configure failed: " + sb.last_error())
    tiles = args.tiles or lanes * args.gpus
    uniq = w.make_unique_streams()
    n, n_learn = w.n, w.n_learn
    den = [sb.Denoiser2D(w.sr, w.frame_ms) for _ in range(2)]
    x_pin = [api.PinnedArray(n) for _ in range(2)]
    y_pin = [api.PinnedArray(n) for _ in range(2)]
    for p, c in zip(x_pin, uniq):
        p.array[:] = c

    def step(tiled: bool):
        for d in den:
            d.reset_profile()
            d.load_parameters(learn_noise=1)
        if not sb.process_batch(den, [p.array[:n_learn] for p in x_pin], [p.array[:n_learn] for p in y_pin]):
            raise RuntimeError(sb.last_error())
        for c, d in enumerate(den):
            d.load_parameters(**w.params[c])
        if tiled:
            if not api.process_tiled_batch(den, [p.array[n_learn:] for p in x_pin],
                                           [p.array[n_learn:] for p in y_pin], tiles, args.warmup_frames, devices):
                raise RuntimeError(sb.last_error())
        elif not sb.process_batch(den, [p.array[n_learn:] for p in x_pin], [p.array[n_learn:] for p in y_pin]):
            raise RuntimeError(sb.last_error())

    def sync_all():
        for d in devices:
            torch.cuda.synchronize(d)

    def timed(tiled: bool, steps: int, warmup: int):
        for _ in range(warmup):
            step(tiled)
        sync_all()
        L.specbleach_b200_stats_reset()
        samplers = [ClockSampler(d) for d in devices]
        for sm in samplers:
            sm.start()
        t0 = time.perf_counter()
        for _ in range(steps):
            step(tiled)
        sync_all()
        wall = (time.perf_counter() - t0) * 1e3
        clocks = [sm.stop() for sm in samplers]
        return wall, clocks, api.stats()

    # untiled run first (one GPU, the e2e path of the default bench): reference output + v_1
    ms_1, _, _ = timed(False, args.steps, args.warmup)
    ref_out = [p.array.copy() for p in y_pin]
    ms_t, clocks, st = timed(True, args.steps, args.warmup)
    same = all(np.array_equal(a, p.array) for a, p in zip(ref_out, y_pin))
    maxdiff = max(float(np.max(np.abs(a - p.array))) for a, p in zip(ref_out, y_pin))
    audio = w.seconds * args.steps
    v1 = audio / (ms_1 / 1e3)
    vt = audio / (ms_t / 1e3)
    line = {
        "metric": w.metric + " -- ONE file time-tiled over the GPUs (N3)", "value": vt, "unit": "x realtime",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_t / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(w.config_dict(), workload=w.what.replace("(one file per GPU)", "(ONE file for all GPUs)"),
                       gpus=args.gpus, tiles_per_channel=tiles, warmup_frames=args.warmup_frames, lanes_per_gpu=lanes,
                       timing="wall clock of the single driving process around synchronous "
                              "specbleach_b200_process_tiled calls, every device synchronised on both sides "
                              "(CUDA events do not span devices)"),
        "e2e": {"value": vt, "unit": "x realtime", "h2d_bytes_per_step": 2 * n * 4, "d2h_bytes_per_step": 2 * n * 4,
                "ms_per_step": ms_t / args.steps,
                "api": "specbleach_b200_process_tiled_batch over both channels (page-locked host buffers, tiles "
                       "dealt to the lanes of every GPU), learn phase through specbleach_b200_process_batch"},
        "untiled_one_gpu": {"value": v1, "ms_per_step": ms_1 / args.steps,
                            "api": "specbleach_b200_process_batch, same host buffers, device 0"},
        "speedup_over_untiled_one_gpu": vt / v1,
        "strong_scaling_efficiency": vt / (v1 * args.gpus),
        "bit_identical_to_untiled": same, "max_abs_diff_to_untiled": maxdiff,
        "gpu_launches": int(st["kernel_launches"]),
        "collective": "none: every tile reads its slice of the input and writes its slice of the output",
        "clocks": clocks[0], "clocks_all_gpus": clocks,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE config (1-based as in SURVEY 8d); 2 = the headline workload")
    ap.add_argument("--cpu-sample-seconds", type=float, default=11.0,
                    help="audio seconds per CPU worker per step in the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tiled", action="store_true",
                    help="config 2 only: ONE 10-min stereo file, each channel cut into time tiles "
                         "(specbleach_b200_process_tiled) over the lanes of --gpus GPUs driven by this one "
                         "process (strong scaling; run WITHOUT torchrun)")
    ap.add_argument("--tiles", type=int, default=0, help="time tiles per channel (0 = one per lane and GPU)")
    ap.add_argument("--warmup-frames", type=int, default=512)
    ap.add_argument("--file-seconds", type=float, default=0.0,
                    help="--tiled only: length of the one stereo file (0 = the config's 600 s)")
    ap.add_argument("--lanes", type=int, default=0, help="engine lanes (0 = library default)")
    ap.add_argument("--chunk-frames", type=int, default=0, help="frames per sub-call (0 = default)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.tiled:
        run_tiled(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
