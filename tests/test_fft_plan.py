"""The FFT plan of odd frame sizes, restated in numpy (no GPU): which (sample rate, frame) pairs go
through Bluestein's convolution (libspecbleach_b200/csrc/sb_geometry.cu, build_fft_plan) and how
accurate that form is in float32.  The kernel itself is checked on the GPU by
tests/test_gpu_parity.py::test_rfft_kernels; this file pins the arithmetic the kernel implements:
    Z[k] = w[k] * sum_n (z[n] w[n]) conj(w)[k - n],  w[n] = exp(-i pi n^2 / M)
as FFT_L, product with FFT_L(conj(w) wrapped to +-n) / L, inverse FFT_L, for L = the smallest
2^a, 3*2^a or 5*2^a >= 2M - 1."""
import math

import numpy as np
import pytest


def factorize(m):
    f = []
    for p in (4, 2, 5, 3):
        while m % p == 0:
            f.append(p)
            m //= p
    p = 7
    while p * p <= m:
        while m % p == 0:
            f.append(p)
            m //= p
        p += 2
    if m > 1:
        f.append(m)
    return f


def bluestein_length(need):
    best = 0
    for odd in (1, 3, 5):
        L = odd * 16
        while L < need:
            L <<= 1
        best = L if best == 0 or L < best else best
    return best


def plan(sr, ms):
    frame = int(np.float32(np.float32(ms) / np.float32(1000.0)) * np.float32(sr))
    n = frame + 800
    n += n & 1
    m = n // 2
    work = sum(4.0 * m * r for r in factorize(m) if r not in (2, 3, 4, 5))
    L = bluestein_length(2 * m - 1)
    use = work > 2.0 * 5.0 * L * math.log2(L) and 2 * 8 * L <= 200 * 1024
    return m, L, use


@pytest.mark.parametrize("sr,ms,m,L,use", [
    (44100, 50.0, 1503, 3072, True),      # BASELINE configs 0 and 3: 3^2 * 167
    (22050, 46.0, 907, 2048, True),       # prime M
    (88200, 50.0, 2605, 6144, True),      # 5 * 521
    (48000, 50.0, 1600, 4096, False),     # the headline geometry stays mixed radix 4 4 4 5 5
    (44100, 20.0, 841, 2048, False),      # 29 * 29: two cheap generic passes
    (176400, 50.0, 4810, 10240, False),   # a factor 37 is cheaper than two 10240-point transforms
    (96000, 170.0, 8560, 20480, False),   # factor 107, but 2 x 20480 points do not fit in shared memory
])
def test_plan_choice(sr, ms, m, L, use):
    assert plan(sr, ms) == (m, L, use)


@pytest.mark.parametrize("m", [1503, 907, 2605, 2557])
def test_bluestein_float32_accuracy(m):
    L = bluestein_length(2 * m - 1)
    n = np.arange(m, dtype=np.int64)
    w = np.exp(-1j * np.pi * ((n * n) % (2 * m)) / m)
    b = np.zeros(L, dtype=np.complex128)
    b[:m] = np.conj(w)
    b[L - m + 1:] = np.conj(w[1:][::-1])
    B = (np.fft.fft(b) / L).astype(np.complex64)
    # the table as the library builds it: b is symmetric, so a cosine sum
    k = np.arange(L)[:64, None]
    Bc = (np.conj(w[0]) + 2.0 * (np.conj(w[1:])[None, :] * np.cos(2 * np.pi * ((k * n[1:][None, :]) % L) / L)).sum(1)) / L
    assert np.abs(Bc - np.fft.fft(b)[:64] / L).max() < 1e-15
    w32 = w.astype(np.complex64)
    rng = np.random.default_rng(m)
    worst = 0.0
    for _ in range(8):
        x = (0.1 * rng.standard_normal(2 * m) + 0.5 * np.sin(2 * np.pi * 0.003 * np.arange(2 * m))).astype(np.float32)
        z = (x[0::2] + 1j * x[1::2]).astype(np.complex64)
        a = np.zeros(L, dtype=np.complex64)
        a[:m] = z * w32
        c = np.fft.ifft(np.fft.fft(a) * B) * np.float32(L)
        assert c.dtype == np.complex64                      # float32 arithmetic throughout
        got = w32 * c[:m].astype(np.complex64)
        ref = np.fft.fft(z.astype(np.complex128))
        worst = max(worst, float(np.abs(got - ref).max() / np.abs(ref).max()))
        # the inverse (unnormalised) conjugates the chirp and the table
        ai = np.zeros(L, dtype=np.complex64)
        ai[:m] = z * np.conj(w32)
        ci = np.fft.ifft(np.fft.fft(ai) * np.conj(B)) * np.float32(L)
        goti = np.conj(w32) * ci[:m].astype(np.complex64)
        refi = np.fft.ifft(z.astype(np.complex128)) * m
        worst = max(worst, float(np.abs(goti - refi).max() / np.abs(refi).max()))
    assert worst < 5e-7, worst
