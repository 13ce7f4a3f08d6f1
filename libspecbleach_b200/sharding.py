"""Host-side sharding of independent units (stream, channel) across ranks.

libspecbleach handles are mono and share nothing (reference:
src/processors/specbleach_2d_denoiser.c:30-38), so multi-GPU is pure data
parallelism: every rank owns a contiguous block of units, no collective touches
the data path; only the timing is reduced (max over ranks)."""
from __future__ import annotations


def shard_bounds(n_units: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced partition: the first ``n_units % world`` ranks get one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def shard_units(units, world: int, rank: int):
    lo, hi = shard_bounds(len(units), world, rank)
    return units[lo:hi]


def job_throughput(local_units_seconds: float, local_ms: float, dist=None, device=None):
    """Whole-job audio-seconds per second: sum of audio over ranks / max time over ranks."""
    if dist is None or not dist.is_initialized():
        return local_units_seconds / (local_ms / 1e3), local_ms
    import torch
    t = torch.tensor([local_ms], dtype=torch.float64, device=device)
    a = torch.tensor([local_units_seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(a, op=dist.ReduceOp.SUM)
    return float(a.item()) / (float(t.item()) / 1e3), float(t.item())


# ---------------------------------------------------------------------------
# N3: one long stream cut into time tiles that run on different
# handles / GPUs.  The product path for this is the C ABI's specbleach_b200_process_tiled[_batch]
# (tiles over the lanes and GPUs of ONE process, csrc/sb_engine.cu); the helpers below are the
# host-side counterpart for ranks that are separate processes (one per GPU): the same tile plan,
# the profile handed over as four small arrays.
# A tile starts `warmup_frames` hops early on a fresh handle that was given the
# finalised noise profile; the warm-up output is thrown away.  What a
# handle carries between frames forgets its start geometrically:
# the 21-frame NLM ring and the 4-frame delay line are exact after 25 frames, the
# veto EMAs (0.5 per frame) after ~150, the forward-masking recursion (<= 0.91 per
# frame at 48 kHz / 50 ms) after ~400 -- so with >= 512 warm-up frames the stitched
# output equals the monolithic one to float rounding in manual mode.  Adaptive
# estimators (Martin: 96-frame minimum history, Brandt: 150-frame window plus the
# "hold last accepted" state) are only approximately reproduced.
# ---------------------------------------------------------------------------
def time_tiles(n_samples: int, hop: int, tiles: int, warmup_frames: int = 512, first: int = 0):
    """[(in_lo, out_lo, out_hi)] -- tile i denoises input[in_lo:out_hi] and owns
    output[out_lo:out_hi]; all bounds are ``first`` plus a multiple of ``hop`` so that
    the tile's STFT frames coincide with the monolithic run's (frames start at multiples
    of hop from the start of the stream).  Tile 0 starts at ``first`` with no warm-up
    (it continues the handle that learned the profile)."""
    if tiles < 1 or hop < 1 or first < 0 or first > n_samples:
        raise ValueError("bad tiling")
    body = n_samples - first
    frames = body // hop
    out = []
    for i in range(tiles):
        lo_f, hi_f = shard_bounds(frames, tiles, i)
        out_lo = first + lo_f * hop
        out_hi = n_samples if i == tiles - 1 else first + hi_f * hop
        in_lo = out_lo if i == 0 else max(first, out_lo - warmup_frames * hop)
        if out_hi > out_lo:
            out.append((in_lo, out_lo, out_hi))
    return out


def denoise_tiled(make_denoiser, x, learn_samples: int, params: dict, tiles: int,
                  warmup_frames: int = 512, hop: int | None = None, world: int = 1, rank: int = 0):
    """Reference flow (learn on x[:learn_samples], then denoise) with the denoise part cut
    into time tiles.  ``make_denoiser()`` returns a fresh handle wrapper (api.Denoiser).
    With ``world`` > 1 every rank computes only its own tiles (static block partition) and
    returns {out_lo: array}; rank 0's dict also holds the learn-phase output under key 0.
    The finalised profile travels as four small arrays (no data-path collective)."""
    import numpy as np
    d0 = make_denoiser()
    if hop is None:
        hop = d0.geometry()["hop"]
    if learn_samples % hop:
        raise ValueError("learn_samples must be a multiple of the hop for frame-aligned tiles")
    pieces = {}
    if learn_samples:
        d0.load_parameters(learn_noise=1)
        head = d0.process(x[:learn_samples])
        if rank == 0:
            pieces[0] = head
    plan = time_tiles(len(x), hop, tiles, warmup_frames, first=learn_samples)
    lo, hi = shard_bounds(len(plan), world, rank)
    profile = None
    finalised = False
    for i in range(lo, hi):
        in_lo, out_lo, out_hi = plan[i]
        if i == 0:
            d = d0                                   # continues the learning handle, as the reference does
            d.load_parameters(**params)
            y = d.process(x[in_lo:out_hi])
            finalised = True
        else:
            if profile is None:
                if not finalised:
                    # the first denoise-mode frame finalises the learned profiles
                    # (noise_estimator.c:143-162); a rank that does not own tile 0
                    # triggers it with one throw-away hop
                    d0.load_parameters(**params)
                    d0.process(np.zeros(hop, dtype=np.float32))
                    finalised = True
                profile = [(d0.get_profile(m), d0.block_count(m), d0.profile_available(m))
                           for m in (1, 2, 3, 4)]
            d = make_denoiser()
            for m, (arr, blocks, avail) in zip((1, 2, 3, 4), profile):
                if avail:
                    d.load_profile(arr, blocks, m)
            d.load_parameters(**params)
            y = d.process(x[in_lo:out_hi])
            d.close()
        pieces[out_lo] = y[out_lo - in_lo:]
    return pieces


def stitch(pieces: dict, n_samples: int):
    import numpy as np
    out = np.zeros(n_samples, dtype=np.float32)
    for lo in sorted(pieces):
        out[lo:lo + len(pieces[lo])] = pieces[lo]
    return out
