"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's denoising hot path.

This file is the *oracle* of specbleach-b200: an independent, readable, float32
restatement of what ``/root/reference`` computes for ``specbleach_process`` and
``specbleach_2d_process`` (learn mode, manual profile, SPP-MMSE, Martin
and Brandt adaptive estimators).  It is pinned against the reference itself:
``tests/test_oracle.py`` checks it against ``oracle/_ref`` (the unmodified reference
sources compiled by ``oracle/Makefile``) and against the committed golden vectors in
``tests/golden/`` that were generated from that build.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import it; the product
never does.

Every function cites the reference lines it restates (paths relative to the
reference tree).  The FFT (FFTW3f in the reference, un-vendored third party) is
evaluated with numpy's double-precision pocketfft and rounded to float32 once.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

f32 = np.float32
EPS = f32(1e-12)               # configurations.h:50
FLT_MIN = f32(np.finfo(np.float32).tiny)
BARK = np.array([100, 200, 300, 400, 510, 630, 770, 920, 1080, 1270, 1480, 1720, 2000, 2320,
                 2700, 3150, 3700, 4400, 5300, 6400, 7700, 9500, 12000, 15500, 19000], dtype=f32)


# ---------------------------------------------------------------------------
# geometry and tables
# ---------------------------------------------------------------------------
@dataclass
class Geometry:
    sr: int
    frame: int
    N: int
    hop: int
    K: int
    cp: int
    win: np.ndarray
    scale: np.float32
    band_start: np.ndarray
    band_end: np.ndarray
    centre: np.ndarray
    fwd: np.ndarray
    bwd: np.float32
    spl_ref: np.float32
    abs_thr: np.ndarray


def make_geometry(sr: int, frame_ms: float) -> Geometry:
    # stft_processor.c:57-58,68 ; fft_transform.c:59,119-123
    frame = int(f32(f32(frame_ms) / f32(1000.0)) * f32(sr))
    N = frame + 800
    N += N & 1
    hop = frame // 4
    K = N // 2 + 1
    cp = N // 2 - frame // 2
    # spectral_utils.c:34-37 (periodic Hann over `frame`), general_utils.c:26-31
    k = np.arange(frame, dtype=f32)
    p = k / f32(frame)
    win = (f32(0.5) - f32(0.5) * np.cos(f32(2.0) * f32(math.pi) * p, dtype=f32)).astype(f32)
    win[~np.isfinite(win) | (np.abs(win) < FLT_MIN)] = 0
    # stft_windows.c:103-118
    s = f32(0)
    for v in win:
        s = f32(s + v * v)
    scale = f32(s * (f32(N) / f32(hop)))
    # critical_bands.c:96-116,138-149 (bin edges computed with K as the fft size)
    nyq = f32(sr) / f32(2.0)
    nb = int(np.max(np.nonzero(BARK < nyq)[0])) + 1
    delim = np.zeros(nb, dtype=np.int64)
    nbins = np.zeros(nb, dtype=np.int64)
    for b in range(nb):
        bi = int(f32(BARK[b] / (f32(sr) / f32(K))) + f32(0.5))
        if b == 0:
            nbins[b] = bi
            delim[b] = bi
        elif b == nb - 1:
            delim[b] = K
            nbins[b] = K - delim[b - 1]
        else:
            nbins[b] = bi - delim[b - 1]
            delim[b] = bi
    start = delim - nbins
    centre = (start.astype(f32) + (delim - start).astype(f32) / f32(2.0)).astype(f32)  # masking_veto.c:103-110
    # masking_estimator.c:146-160
    hop_time = f32(N) / (f32(4.0) * f32(sr))
    bark = np.minimum(np.arange(nb, dtype=f32), f32(24.0))
    w = bark / f32(24.0)
    tau = (f32(1.0) - w) * f32(0.100) + w * f32(0.025)
    fwd = np.exp(-hop_time / tau, dtype=f32)
    bwd = f32(np.exp(f32(-hop_time / f32(0.010))))
    # absolute_hearing_thresholds.c:86-89,118-139 (Vorbis-windowed 1 kHz sine, peak power)
    n = np.arange(N, dtype=f32)
    sine = np.sin((f32(2.0) * f32(math.pi) * n * f32(1000.0)) / f32(sr), dtype=f32)
    pv = n / f32(N)
    vorbis = np.sin(f32(math.pi) / f32(2.0) * np.sin(f32(math.pi) * pv, dtype=f32) ** 2, dtype=f32)
    X = np.fft.rfft((sine * vorbis).astype(np.float64))
    P = (X.real ** 2 + X.imag ** 2).astype(f32)
    spl_ref = f32(90.0) - f32(10.0) * np.log10(f32(P.max()) + EPS, dtype=f32)
    # absolute_hearing_thresholds.c:161-170
    fr = np.maximum(np.arange(K, dtype=f32) * (f32(sr) / f32(N)), f32(20.0)) / f32(1000.0)
    abs_thr = (f32(3.64) * fr ** f32(-0.8) - f32(6.5) * np.exp(f32(-0.6) * (fr - f32(3.3)) ** 2, dtype=f32)
               + f32(1e-3) * fr ** 4).astype(f32)
    return Geometry(sr, frame, N, hop, K, cp, win, scale, start, delim.copy(), centre, fwd, bwd,
                    f32(spl_ref), abs_thr)


# ---------------------------------------------------------------------------
# small vector helpers
# ---------------------------------------------------------------------------
def smooth3(p: np.ndarray, f: float) -> np.ndarray:
    """spectral_utils.c:198-222 (all taps read the ORIGINAL values)."""
    f = f32(f)
    if p.size < 2 or f <= 0:
        return p.copy()
    y = np.empty_like(p)
    y[0] = (p[0] + p[1]) / f32(2) * f + p[0] * (f32(1) - f)
    y[1:-1] = (p[:-2] + p[1:-1] + p[2:]) / f32(3) * f + p[1:-1] * (f32(1) - f)
    y[-1] = (p[-2] + p[-1]) / f32(2) * f + p[-1] * (f32(1) - f)
    return y.astype(f32)


def interp_gaps(p: np.ndarray, thr: float) -> np.ndarray:
    """spectral_utils.c:224-251."""
    s = p.copy()
    n = s.size
    if n < 3:
        return s
    i = 1
    while i < n - 1:
        if s[i] < thr:
            j = i + 1
            while j < n and s[j] < thr:
                j += 1
            if j < n:
                a, b = s[i - 1], s[j]
                step = f32((b - a) / f32(j - (i - 1)))
                for k in range(i, j):
                    s[k] = f32(a + step * f32(k - (i - 1)))
                i = j
        i += 1
    return s


def morph(prof: np.ndarray, a: float) -> np.ndarray:
    """get_morphed_profile, spectral_utils.c:253-279. prof rows: mean, median, max, min."""
    a = f32(a)
    if a < 0:
        t = -a
        v = prof[0] * (f32(1) - t) + prof[1] * t
    else:
        v = prof[0] * (f32(1) - a) + prof[2] * a
    return np.maximum(v.astype(f32), prof[3])


def fast_exp(x: np.ndarray) -> np.ndarray:
    """sb_fast_expf, simd_utils.h:670-684 (Schraudolph bit trick, 0 below -20)."""
    x = np.asarray(x, dtype=f32)
    dead = ~(x >= f32(-20.0))                       # x < -20 (or NaN) -> weight 0
    xs = np.where(dead, f32(0), x).astype(f32)
    v = (f32(12102203.0) * xs + f32(1064866805.0)).astype(f32)
    out = v.astype(np.int64).astype(np.uint32).view(f32).copy()
    out[dead] = 0
    return out


def db_to_coeff(db: float) -> np.float32:
    return f32(np.exp(f32(f32(db) / f32(20.0) * np.log(f32(10.0)))))   # general_utils.c:33-35


# ---------------------------------------------------------------------------
# NLM (nlm_filter_avx.c:87-248 / nlm_filter.c:261-430)
# ---------------------------------------------------------------------------
def nlm_frame(ring: np.ndarray, h: float) -> np.ndarray:
    """``ring[i]`` = SNR frame at offset i-16 from the target (21 rows).  Patch rows go
    through the 21-slot ring: offset x -> ((x+16) mod 21) (nlm_filter_internal.h:74-86)."""
    K = ring.shape[1]
    h = f32(h)
    target = ring[16]
    nblk = (K + 7) // 8
    bs = np.arange(nblk) * 8
    bc = np.minimum(bs + 4, K - 1)
    lim = np.minimum(8, K - bs)
    q = np.arange(8)
    clip = lambda a: np.clip(a, 0, K - 1)
    tcols = clip(bc[:, None] - 4 + q[None, :])                           # (nblk, 8)
    dfs = np.arange(-8, 9)
    cc = clip(bc[:, None] + dfs[None, :])                                # (nblk, 17)
    ccols = clip(cc[:, :, None] - 4 + q[None, None, :])                  # (nblk, 17, 8)
    pcols = clip(bs[:, None, None] + q[None, None, :] + dfs[None, :, None])  # paste bins (nblk,17,8)
    nf = bc.astype(f32) / f32(K - 1)
    fs = f32(1.0) + f32(3.0) * nf * nf
    hv = (h * fs).astype(f32)
    h2 = (hv * hv).astype(f32)
    inv = (f32(1.0) / h2).astype(f32)
    thr = (f32(4.0) * h2).astype(f32)
    live = np.array([np.sum(target[bs[b]:bs[b] + lim[b]], dtype=f32) >= f32(1e-6) for b in range(nblk)])
    acc = np.zeros((nblk, 8), dtype=f32)
    wsum = np.zeros(nblk, dtype=f32)
    trow = [ring[(r - 4 + 16) % 21][tcols] for r in range(8)]           # target patch rows (nblk, 8)
    for dt in range(-16, 5):
        lanes = np.zeros((nblk, 17, 8), dtype=f32)
        for r in range(8):
            crow = ring[(dt + r - 4 + 16) % 21]
            d = trow[r][:, None, :] - crow[ccols]
            lanes = (lanes + d * d).astype(f32)
        # horizontal sum ((a0+a4)+(a1+a5))+((a2+a6)+(a3+a7))  (simd_utils.h:260-276)
        dist = ((lanes[..., 0] + lanes[..., 4]) + (lanes[..., 1] + lanes[..., 5])) + \
               ((lanes[..., 2] + lanes[..., 6]) + (lanes[..., 3] + lanes[..., 7]))
        w = fast_exp((-dist * inv[:, None]).astype(f32))
        w = np.where((dist > thr[:, None]) | (w < f32(1e-10)) | ~live[:, None], f32(0), w).astype(f32)
        cand = ring[dt + 16][pcols]                                      # (nblk, 17, 8)
        for j in range(17):                                              # df order
            acc = (acc + w[:, j, None] * cand[:, j, :]).astype(f32)
            wsum = (wsum + w[:, j]).astype(f32)
    out = target.copy()
    for b in range(nblk):
        if wsum[b] > f32(1e-10):
            out[bs[b]:bs[b] + lim[b]] = acc[b, :lim[b]] / wsum[b]
    return out.astype(f32)


# ---------------------------------------------------------------------------
# the denoiser (streaming-equivalent, whole-call form)
# ---------------------------------------------------------------------------
@dataclass
class Params:
    learn_noise: int = 0
    residual_listen: bool = False
    reduction: np.float32 = f32(1.0)
    smoothing: np.float32 = f32(0.0)
    whitening: np.float32 = f32(0.0)
    adaptive: int = 0
    method: int = 0
    depth: np.float32 = f32(0.0)
    strength: np.float32 = f32(0.0)
    aggressiveness: np.float32 = f32(0.0)
    tonal: np.float32 = f32(1.0)


@dataclass
class OracleDenoiser:
    """Mirrors one handle of the reference (1D: specbleach_denoiser.c, 2D:
    specbleach_2d_denoiser.c).  ``process`` may be called with any chunking."""
    sr: int
    frame_ms: float
    two_d: bool = True
    g: Geometry = field(init=False)

    def __post_init__(self):
        g = self.g = make_geometry(self.sr, self.frame_ms)
        K = g.K
        self.prm = Params()
        self.prof = np.zeros((4, K), dtype=f32)           # mean, median, max, min
        self.count = 0
        self.available = [False] * 4
        self.med_hist = np.zeros((25, K), dtype=f32)
        self.med_wr = 0
        self.was_learning = False
        self.fifo = np.zeros(g.frame - g.hop, dtype=f32)  # input FIFO tail (stft_buffer.c)
        self.hopbuf = np.zeros(g.hop, dtype=f32)
        self.fill = 0
        self.acc = np.zeros(2 * g.frame, dtype=f32)       # OLA accumulator (stft_processor.c:162-170)
        self.out_fifo = np.zeros(g.hop, dtype=f32)
        self.Xhist = [np.zeros(K, dtype=np.complex64) for _ in range(4)]
        self.nhist = [np.zeros(K, dtype=f32) for _ in range(4)]
        self.Sring = [np.zeros(K, dtype=f32) for _ in range(21)]
        self.pushed = 0
        self.stable = np.zeros(K, dtype=f32)
        self.sp_prev = np.zeros(K, dtype=f32)
        nb = g.centre.size
        self.T_prev = np.zeros(nb, dtype=f32)
        self.p_mem = np.zeros(nb, dtype=f32)
        self.nlm_h = f32(1.0)
        self.est = None
        self.last_adaptive_state = 0
        self.floor = np.zeros(K, dtype=f32)

    # -- parameters (specbleach_2d_denoiser.c:223-255, specbleach_denoiser.c:199-229) --
    def load_parameters(self, **kw):
        p = Params()
        p.learn_noise = int(kw.get("learn_noise", 0))
        p.residual_listen = bool(kw.get("residual_listen", False))
        p.reduction = db_to_coeff(-float(kw.get("reduction_amount", 0.0)))
        sm = f32(kw.get("smoothing_factor", 0.0))
        p.smoothing = sm if self.two_d else f32(1.0) - f32(np.exp(f32(-3.0) * (sm / f32(100.0))))
        p.whitening = f32(kw.get("whitening_factor", 0.0)) / f32(100.0)
        p.adaptive = int(kw.get("adaptive_noise", 0))
        p.method = int(kw.get("noise_estimation_method", 0))
        p.depth = f32(kw.get("nlm_masking_protection", kw.get("masking_depth", 0.0)))
        p.strength = f32(kw.get("suppression_strength", 0.0)) / f32(100.0)
        p.aggressiveness = f32(kw.get("aggressiveness", 0.0))
        p.tonal = db_to_coeff(-float(kw.get("tonal_reduction", 0.0)))
        if p.adaptive and (self.est is None or self.est["method"] != p.method):
            K = self.g.K   # spectral_2d_denoiser.c:296-311
            self.est = {"method": p.method, "first": True, "psd": np.zeros(K, f32),
                        "sspp": np.zeros(K, f32), "cmin": np.zeros(K, f32),
                        "hist": np.zeros((8, K), f32), "fc": 0, "idx": 0}
            if p.method == 1:
                self.est.update(brandt_init(K, self.sr, self.g.N))
            self.last_adaptive_state = 0
        self.prm = p
        if self.two_d and p.smoothing > 0:
            self.nlm_h = p.smoothing

    # -- streaming shell: stft_processor.c:126-175, stft_buffer.c:76-103 --
    def process(self, x: np.ndarray) -> np.ndarray:
        g = self.g
        x = np.asarray(x, dtype=f32)
        out = np.empty_like(x)
        i, n = 0, x.size
        while i < n:
            take = min(g.hop - self.fill, n - i)
            out[i:i + take] = self.out_fifo[self.fill:self.fill + take]   # latency = frame
            self.hopbuf[self.fill:self.fill + take] = x[i:i + take]
            self.fill += take
            i += take
            if self.fill == g.hop:
                frame = np.concatenate([self.fifo, self.hopbuf])
                self.fifo = frame[g.hop:].copy()
                y = self._frame(frame)
                self.acc[:g.frame] = (self.acc[:g.frame] + y).astype(f32)
                self.out_fifo = self.acc[:g.hop].copy()
                self.acc[:g.frame] = self.acc[g.hop:g.hop + g.frame].copy()
                self.fill = 0
        return out

    # -- one STFT frame: analysis, spectral processing, synthesis --
    def _frame(self, frame: np.ndarray) -> np.ndarray:
        g = self.g
        xt = np.zeros(g.N, dtype=f32)
        xt[g.cp:g.cp + g.frame] = frame * g.win                          # fft_transform.c:165-182
        Xd = np.fft.rfft(xt.astype(np.float64))
        X = (Xd.real.astype(f32) + 1j * Xd.imag.astype(f32)).astype(np.complex64)
        Y = self._spectral(X)
        z = np.fft.irfft(Y.astype(np.complex128), n=g.N) * g.N           # unnormalised HC2R
        z = z.astype(f32)
        return ((z[g.cp:g.cp + g.frame] * g.win) / g.scale).astype(f32)  # stft_windows.c:120-137

    def _spectral(self, X: np.ndarray) -> np.ndarray:
        g, p = self.g, self.prm
        re, im = X.real.astype(f32), X.imag.astype(f32)
        P = (re * re + im * im).astype(f32)                              # spectral_features.c:80-107
        # ---- learn mode (denoiser_profile_core.c:28-51, noise_estimator.c:83-141) ----
        if p.learn_noise > 0:
            c = self.count
            self.prof[0] = P if c <= 1 else (self.prof[0] + (P - self.prof[0]) / f32(c)).astype(f32)
            self.count += 1
            if self.count > 5:
                self.available[0] = True
            self.med_hist[self.med_wr] = P
            self.med_wr = (self.med_wr + 1) % 25
            self.prof[1] = np.sort(self.med_hist, axis=0)[12]
            self.prof[2] = np.maximum(self.prof[2], P)
            self.prof[3] = np.minimum(self.prof[3], P)
            self.available[1] = self.available[2] = self.available[3] = True
            self.was_learning = True
            return X
        if self.was_learning:
            for m in range(4):                                           # noise_estimator.c:143-162
                if self.available[m]:
                    self.prof[m] = smooth3(interp_gaps(self.prof[m], 1e-9), 0.5)
            self.was_learning = False
        # ---- noise for this frame (denoiser_profile_core.c:53-102) ----
        if p.adaptive and self.est is not None:
            if self.last_adaptive_state == 0:
                self.floor = morph(self.prof, p.aggressiveness)
                self._seed(self.floor)
                self.last_adaptive_state = 1
            noise = self._estimate(P)
            noise = smooth3(interp_gaps(noise, 1e-9), 0.5)               # adaptive_noise_estimator.c:174-187
            self._apply_floor(self.floor)
            noise = np.maximum(noise, self.floor)
        else:
            self.last_adaptive_state = 0
            noise = morph(self.prof, p.aggressiveness)
        if not self.two_d:
            return self._gains_1d(X, P, noise)
        # ---- 2D: delay line, SNR, NLM (spectral_2d_denoiser.c:362-446) ----
        self.Xhist.append(X)
        self.nhist.append(noise)
        Xt = self.Xhist.pop(0)
        nt = self.nhist.pop(0)
        S = (P / np.maximum(noise, f32(1e-9))).astype(f32)               # nlm_filter.c:453-479
        self.Sring.append(S)
        self.Sring.pop(0)
        self.pushed = min(self.pushed + 1, 21)
        if self.pushed < 21:
            return Xt
        Shat = nlm_frame(np.stack(self.Sring), self.nlm_h)
        Mg = (Shat * np.maximum(nt, f32(1e-9))).astype(f32)              # nlm_filter.c:481-508
        alpha = self._alpha(Mg, nt)
        alpha = self._veto(Mg, nt, Xt.real.astype(f32), alpha)
        gain = self._gain(Mg, nt, alpha)
        return self._apply(Xt, gain)

    def _gains_1d(self, X, P, noise):
        p = self.prm                                                     # spectral_denoiser.c:258-343
        alpha = self._alpha(P, noise)
        s = p.smoothing
        sp = np.where(P > self.sp_prev, s * self.sp_prev + (f32(1) - s) * P, P).astype(f32)
        self.sp_prev = sp                                                # spectral_smoother.c:111-139,183-192
        alpha = self._veto(sp, noise, None, alpha)
        gain = self._gain(sp, noise, alpha)
        return self._apply(X, gain)

    # -- Berouti alpha + tonal boost (suppression_engine.c:108-133, tonal_reducer.c:54-114) --
    def _tonal_mask(self, noise):
        K = self.g.K
        det = self.prof[2] if np.any(self.prof[1] > 0) else noise        # tonal_detector.c:35-44
        idx = np.clip(np.arange(K)[:, None] + np.arange(-25, 26)[None, :], 0, K - 1)
        fl = np.sort(det[idx], axis=1)[:, 25]
        ok = ~((det - fl) < f32(1e-7)) & ((det / (fl + f32(1e-20))) > f32(1.41))
        return np.where(ok, (det - fl) / (det + f32(1e-20)), f32(0)).astype(f32)

    def _alpha(self, spec, noise):
        p = self.prm
        amax = f32(1.0) + p.strength * f32(3.0)
        with np.errstate(divide="ignore", invalid="ignore"):
            snr = (f32(10.0) * np.log10(spec / (noise + EPS), dtype=f32)).astype(f32)
        ns = (snr + f32(5.0)) / f32(25.0)
        a = np.where(snr <= f32(-5.0), amax, np.where(snr >= f32(20.0), f32(1.0),
                                                       amax - ns * (amax - f32(1.0)))).astype(f32)
        self.mask = self._tonal_mask(noise)
        if not (p.tonal >= f32(0.999)):
            need = f32(1.0) + (f32(1.0) - p.tonal) * f32(9.0)
            tgt = f32(1.0) + self.mask * (need - f32(1.0))
            a = np.where(self.mask > 0, np.maximum(a, tgt), a).astype(f32)
        return a

    # -- masking veto (masking_veto.c:127-267, masking_estimator.c:188-349) --
    def _thresholds(self, spec):
        g = self.g
        nb = g.centre.size
        E = np.array([np.sum(spec[s:e], dtype=f32) for s, e in zip(g.band_start, g.band_end)], dtype=f32)
        L = (f32(10.0) * np.log10(E + EPS, dtype=f32) + f32(96.0)).astype(f32)
        s_up = np.clip(f32(15.0) - (L - f32(40.0)) * f32(0.2), f32(5.0), f32(15.0)).astype(f32)
        i = np.arange(nb, dtype=f32)
        y = (i[:, None] - i[None, :] + f32(0.474)).astype(f32)           # dz = i - j
        sf = f32(15.81) + ((f32(25.0) - s_up) / f32(2))[None, :] * y - \
            ((f32(25.0) + s_up) / f32(2))[None, :] * np.sqrt(f32(1.0) + y * y, dtype=f32)
        gain = np.power(f32(10.0), (sf / f32(10.0)).astype(f32), dtype=f32)
        spr = np.power(np.sum(np.power((E[None, :] * gain).astype(f32), f32(0.4), dtype=f32), axis=1, dtype=f32),
                       f32(2.5), dtype=f32)
        T = np.empty(nb, dtype=f32)
        for b in range(nb):
            s, e = g.band_start[b], g.band_end[b]
            v = np.maximum(spec[s:e], EPS)
            n = f32(e - s)
            if n <= 1:
                ton = f32(1.0)
            else:
                sfm = f32(10.0) * (np.sum(np.log10(v, dtype=f32), dtype=f32) / n -
                                   np.log10(np.sum(v, dtype=f32) / n, dtype=f32))
                ton = f32(np.clip(sfm / f32(-60.0), 0.0, 1.0))
            off = ton * (f32(14.5) + f32(min(b + 1, 25))) + f32(6.0) * (f32(1.0) - ton)
            T[b] = np.power(f32(10.0), f32(np.log10(spr[b] + EPS, dtype=f32) - off / f32(10.0)), dtype=f32)
        return T

    def _veto(self, Mg, noise, future, alpha):
        g, p = self.g, self.prm
        if p.depth < 0:
            return alpha
        clean = np.maximum(Mg - noise, f32(0))
        self.stable = (f32(0.5) * clean + f32(0.5) * self.stable).astype(f32)
        T = np.power(self._thresholds(self.stable), f32(0.6), dtype=f32)
        T = T + np.power((self.T_prev * g.fwd).astype(f32), f32(0.6), dtype=f32)
        if future is not None:
            fut = np.maximum(future - noise, f32(0))                     # masking_veto.c:151-157 (Re X_t!)
            T = T + np.power((self._thresholds(fut) * g.bwd).astype(f32), f32(0.6), dtype=f32)
        T = np.power(T.astype(f32), f32(1.0) / f32(0.6), dtype=f32)
        self.T_prev = T.copy()
        nb = T.size
        prot = np.zeros(nb, dtype=f32)
        for b in range(nb):
            s, e = g.band_start[b], g.band_end[b]
            if e <= s:
                continue
            # absolute_hearing_thresholds.c:141-159
            spl = f32(10.0) * np.log10(T[b] + EPS, dtype=f32) + g.spl_ref
            mx = np.maximum(spl, g.abs_thr[s:e])
            thr = np.maximum(np.power(f32(10.0), ((mx - g.spl_ref) / f32(10.0)).astype(f32), dtype=f32) - EPS, f32(0))
            nmr = f32(10.0) * np.log10((np.sum(noise[s:e], dtype=f32) + EPS) / (np.sum(thr, dtype=f32) + EPS), dtype=f32)
            prot[b] = 1.0 if nmr <= 0 else (0.0 if nmr >= 20 else f32(1.0) - nmr / f32(20.0))
        self.p_mem = (f32(0.5) * prot + f32(0.5) * self.p_mem).astype(f32)  # spectral_smoother.c:213-223
        pb = self.p_mem.copy()
        ps = pb.copy()                                                   # spectral_smoother.c:194-211
        for b in range(1, nb):
            nxt = pb[b + 1] if b < nb - 1 else pb[b]
            ps[b] = f32(0.25) * pb[b - 1] + f32(0.5) * pb[b] + f32(0.25) * nxt
        K = g.K
        k = np.arange(K, dtype=f32)
        cb = np.zeros(K, dtype=np.int64)
        cur = 0
        for kk in range(K):                                              # masking_veto.c:233-236
            while cur < nb - 1 and f32(kk) > g.centre[cur + 1]:
                cur += 1
            cb[kk] = cur
        nxtb = np.minimum(cb + 1, nb - 1)
        x0, x1 = g.centre[cb], g.centre[nxtb]
        with np.errstate(divide="ignore", invalid="ignore"):
            t = np.clip((k - x0) / (x1 - x0), 0, 1).astype(f32)
        bp = np.where(cb < nb - 1, np.where(x1 > x0, ps[cb] * (f32(1) - t) + ps[nxtb] * t, ps[cb]),
                      ps[nb - 1]).astype(f32)
        sc = np.minimum(self.stable / (noise + EPS), f32(1.0))
        return (alpha + (f32(1.0) - alpha) * (bp * sc * p.depth)).astype(f32)

    # -- Wiener gain + whitened floor (gain_calculator.c:27-69, noise_floor_manager.c:77-163) --
    def _gain(self, spec, noise, alpha):
        p = self.prm
        sn = (noise * alpha).astype(f32)
        with np.errstate(divide="ignore", invalid="ignore"):
            gw = np.where(spec > sn, (spec - sn) / spec, f32(0))
        gain = np.where(sn > FLT_MIN, gw, f32(1.0)).astype(f32)
        if p.reduction >= f32(0.999) and p.tonal >= f32(0.999):
            return np.ones_like(gain)
        sm = smooth3(noise, 0.5)                                         # spectral_whitening.c:59-111
        anchor = max(np.sort(sm)[sm.size // 2], EPS)
        with np.errstate(divide="ignore", invalid="ignore"):
            wt = np.where((p.whitening > 0) & (sm > EPS), np.power(anchor / sm, p.whitening, dtype=f32), f32(1.0))
        m4 = np.where(self.mask > 0, np.sqrt(np.sqrt(self.mask, dtype=f32), dtype=f32), self.mask)
        dual = p.reduction * (f32(1) - m4) + p.tonal * m4
        target = dual + p.whitening * (p.reduction - dual)
        fl = np.minimum(target * (f32(1) + (f32(1) - target) * (wt - f32(1))), f32(1.0))
        return np.maximum(fl, gain).astype(f32)

    def _apply(self, X, gain):
        if self.prm.residual_listen:                                     # denoiser_post_process.c:36-50
            re = X.real - X.real * gain
            im = X.imag - X.imag * gain
        else:
            re, im = X.real * gain, X.imag * gain
        return (re.astype(f32) + 1j * im.astype(f32)).astype(np.complex64)

    # -- adaptive estimators --
    def _seed(self, fl):
        e = self.est
        if e["method"] == 1:
            return brandt_set_state(e, fl)
        v = np.maximum(fl, FLT_MIN)
        if e["method"] == 0:                                             # spp_mmse:169-180
            e["psd"] = v.copy()
            e["sspp"] = np.zeros_like(v)
        else:                                                            # martin:167-187
            v = (v / f32(1.5)).astype(f32)
            e["psd"], e["cmin"] = v.copy(), v.copy()
            e["hist"] = np.tile(v, (8, 1))
            e["first"] = False
            e["fc"] = 0

    def _apply_floor(self, fl):
        e = self.est
        if e["method"] == 1:
            return brandt_apply_floor(e, fl)
        e["psd"] = np.maximum(e["psd"], fl)
        if e["method"] != 0:
            e["cmin"] = np.maximum(e["cmin"], fl)
            e["hist"] = np.maximum(e["hist"], fl[None, :])

    def _estimate(self, P):
        e = self.est
        K = P.size
        energy = f32(np.sum(P, dtype=f32) / f32(K))
        silent = energy < f32(1e-10)
        if e["method"] == 1:
            return brandt_run(e, P, energy)
        if e["method"] == 0:                                             # spp_mmse:98-166
            if e["first"]:
                if silent:
                    return np.zeros(K, f32)
                e["psd"], e["sspp"], e["first"] = P.copy(), np.zeros(K, f32), False
                return P.copy()
            if silent:
                return e["psd"].copy()
            prev = np.maximum(e["psd"], EPS)
            ex = np.exp((-(P / prev) * (f32(31.62) / f32(32.62))).astype(f32), dtype=f32)
            spp = np.clip(f32(1.0) / (f32(1.0) + f32(32.62) * ex), 0, 1).astype(f32)
            spp = np.where(e["sspp"] > f32(0.99), np.minimum(spp, f32(0.99)), spp)
            mm = (f32(1.0) - spp) * P + spp * e["psd"]
            out = (f32(0.8) * e["psd"] + (f32(1.0) - f32(0.8)) * mm).astype(f32)
            e["sspp"] = (f32(0.9) * e["sspp"] + f32(0.1) * spp).astype(f32)
            e["psd"] = out.copy()
            return out
        # Martin (martin_noise_estimator.c:82-164)
        if e["first"]:
            if silent:
                return np.zeros(K, f32)
            v = (P / f32(1.5)).astype(f32)
            e["psd"], e["cmin"], e["hist"] = v.copy(), v.copy(), np.tile(v, (8, 1))
            e["first"], e["fc"] = False, 1
            return P.copy()
        if not silent:
            e["psd"] = (f32(0.8) * e["psd"] + (f32(1.0) - f32(0.8)) * P).astype(f32)
            e["cmin"] = np.minimum(e["cmin"], e["psd"])
            if e["fc"] >= 12:
                e["hist"][e["idx"]] = e["cmin"]
                e["cmin"] = e["psd"].copy()
                e["idx"] = (e["idx"] + 1) % 8
                e["fc"] = 0
        out = (np.minimum(e["cmin"], e["hist"].min(axis=0)) * f32(1.5)).astype(f32)
        e["fc"] += 1
        return out


# ---------------------------------------------------------------------------
# Brandt trimmed-mean estimator (brandt_noise_estimator.c)
# ---------------------------------------------------------------------------
BRANDT_P = (f32(0.1), f32(0.25), f32(0.5), f32(0.75), f32(1.0))


def brandt_correction(p) -> np.float32:
    """calculate_correction_factor, brandt_noise_estimator.c:44-56."""
    p = f32(p)
    if p <= 0 or p >= 1:
        return f32(1.0)
    den = f32(1.0) + (f32(1.0) - p) / p * f32(np.log(f32(1.0) - p))
    return f32(1.0) if abs(den) < f32(1e-6) else f32(f32(1.0) / f32(den))


def brandt_init(K: int, sr: int, N: int) -> dict:
    """brandt_noise_estimator_initialize, brandt_noise_estimator.c:65-128: the history
    length assumes a hop of fft_size/2 (quirk Q6)."""
    frame_dur = max(f32(f32(N) * f32(1000.0) / f32(sr)) * f32(0.5), f32(0.1))
    H = max(int(f32(5000.0) / frame_dur), 5)
    return {"H": H, "bh": np.zeros((K, H), f32), "bidx": 0, "last": np.zeros(K, f32),
            "bfirst": True, "cf": brandt_correction(0.5),
            "cfs": [brandt_correction(p) for p in BRANDT_P]}


def brandt_set_state(e: dict, profile: np.ndarray) -> None:
    """brandt_noise_estimator_set_state / _update_seed, :252-282: history = profile /
    factor(0.5) with a deterministic +-1 % jitter indexed by (t + k) % 11."""
    K, H = e["bh"].shape
    val = (profile * (f32(1.0) / e["cf"])).astype(f32)
    t = np.arange(H)[None, :] + np.arange(K)[:, None]
    jitter = (f32(1.0) + f32(0.01) * ((t % 11) - 5).astype(f32) / f32(5.0)).astype(f32)
    e["bh"] = (val[:, None] * jitter).astype(f32)
    e["last"] = profile.astype(f32).copy()
    e["bfirst"] = False


def brandt_apply_floor(e: dict, fl: np.ndarray) -> None:
    """brandt_noise_estimator_apply_floor, :284-299."""
    fv = (fl * (f32(1.0) / e["cf"])).astype(f32)
    e["bh"] = np.maximum(e["bh"], fv[:, None])


def _gcc_vector_sum(s: np.ndarray, q: int) -> np.ndarray:
    """sum of s[:, :q] in the order gcc -O3 -ffast-math -mavx2 emits for the scalar loop at
    brandt_noise_estimator.c:218-221 (8-lane partial sums, fold to 4 lanes, one more
    4-vector, horizontal add, <= 3 scalar leftovers); verified on the oracle build."""
    K = s.shape[0]
    n8 = q // 8
    acc = np.zeros((K, 8), f32)
    for j in range(n8):
        acc = acc + s[:, 8 * j:8 * j + 8]
    a4 = acc[:, :4] + acc[:, 4:]
    done = 8 * n8
    if q - done >= 4:
        a4 = a4 + s[:, done:done + 4]
        done += 4
    tot = (a4[:, 1] + a4[:, 3]) + (a4[:, 0] + a4[:, 2])
    for j in range(done, q):
        tot = tot + s[:, j]
    return tot.astype(f32)


def brandt_run(e: dict, P: np.ndarray, energy) -> np.ndarray:
    """brandt_noise_estimator_run, :134-250 (all bins at once)."""
    K, H = e["bh"].shape
    thr = f32(1e-10)
    if e["bfirst"] and energy > thr:
        brandt_set_state(e, P)                  # same fill as :147-160 (factor(0.5) == cf)
        e["last"] = P.copy()
    if energy < thr:
        return e["last"].copy()
    e["bh"][:, e["bidx"]] = P
    e["bidx"] = (e["bidx"] + 1) % H
    s = np.sort(e["bh"], axis=1)
    min_ad = np.full(K, f32(2.0))
    best = e["last"].copy()
    for ci, p in enumerate(BRANDT_P):
        q = int(f32(p) * f32(H))
        if q < 10:
            continue
        b = s[:, q - 1]
        mu_trunc = _gcc_vector_sum(s, q) / f32(q)
        ok = mu_trunc > thr
        mu = (mu_trunc * e["cfs"][ci]).astype(f32)
        ad = np.ones(K, f32)                    # calculate_ad_norm, :113-132
        live = ok & (mu >= f32(1e-15))
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            mu_inv = (f32(1.0) / mu).astype(f32)
            denom = (f32(1.0) - np.exp(-(b * mu_inv).astype(f32), dtype=f32)).astype(f32)
            live &= ~(np.abs(denom) < f32(1e-12))
            q_inv = f32(1.0) / f32(q)
            denom_inv = (f32(1.0) / denom).astype(f32)
            e1 = (f32(1.0) - np.exp(-(s[:, :q] * mu_inv[:, None]).astype(f32), dtype=f32)).astype(f32)
            f_emp = (np.arange(1, q + 1, dtype=f32) * q_inv).astype(f32)
            # gcc contracts f_emp - e1*denom_inv into one vfnmadd (single rounding)
            term = np.abs((f_emp[None, :].astype(np.float64)
                           - e1.astype(np.float64) * denom_inv[:, None].astype(np.float64))
                          .astype(f32))
            tot = np.zeros(K, f32)
            for i in range(q):                  # sequential scalar accumulation
                tot = tot + term[:, i]
            adv = ((tot + tot) * q_inv).astype(f32)
        ad = np.where(live, adv, ad).astype(f32)
        upd = ok & (ad < min_ad)
        best = np.where(upd, mu, best)
        min_ad = np.where(ok, np.minimum(min_ad, ad), min_ad)
    accept = ~(f32(1.0) - f32(0.9) < min_ad)    # 1 - min_ad >= 0.9 as gcc folds it
    e["last"] = np.where(accept, best, e["last"]).astype(f32)
    return e["last"].copy()


def run(x: np.ndarray, sr: int, frame_ms: float, two_d: bool, learn_samples: int, params: dict,
        chunk: int | None = None) -> np.ndarray:
    """Learn on the first ``learn_samples`` samples, then denoise the rest (demo flow)."""
    d = OracleDenoiser(sr, frame_ms, two_d)
    outs = []
    chunk = chunk or len(x)
    if learn_samples:
        d.load_parameters(learn_noise=1)
        for i in range(0, learn_samples, chunk):
            outs.append(d.process(x[i:min(i + chunk, learn_samples)]))
    d.load_parameters(**params)
    for i in range(learn_samples, len(x), chunk):
        outs.append(d.process(x[i:min(i + chunk, len(x))]))
    return np.concatenate(outs)


def denoise_2d(x, sr, frame_ms, learn_samples, params, chunk=None):
    return run(x, sr, frame_ms, True, learn_samples, params, chunk)


def denoise_1d(x, sr, frame_ms, learn_samples, params, chunk=None):
    return run(x, sr, frame_ms, False, learn_samples, params, chunk)
