// STFT kernels: framing + Hann + centred zero-pad + real FFT (forward), the
// inverse real FFT + synthesis window, and the ordered overlap-add.
//
// Replaces, for a whole batch of frames at once, the per-hop loop of
//   stft_processor.c:126-175   (FIFO, analysis, synthesis, OLA)
//   fft_transform.c:165-221    (centred load/crop, FFTW R2HC / HC2R)
//   stft_windows.c:103-139     (window and 1/scale)
//   spectral_features.c:80-107 (power spectrum)
//
// FFT: one CTA per frame.  The N real samples are packed into M = N/2 complex
// points, transformed by a mixed-radix Stockham autosort FFT that ping-pongs
// between two shared-memory buffers (radix 4, 2, 5 and 3 hard-wired, any other
// prime factor through a generic O(r) per output butterfly in double), and split into the
// N/2+1 bins of the real transform.  The frame sizes of the library are not
// powers of two (N = frame + 800), hence the runtime factor list.  When M has a
// large prime factor (44.1 kHz / 50 ms: M = 1503 = 3^2 167) the M points go
// through Bluestein's convolution over a length 2^a, 3 2^a or 5 2^a instead
// (fft_points below; plan and tables: build_fft_plan in sb_geometry.cu).
#include "sb_common.h"

namespace sb {

namespace {

constexpr int kFftMaxThreads = 512;   // the launch uses g.fft_threads (sb_geometry.cu)

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// Mixed-radix Stockham DIF passes over M complex points held in shared memory.
// Returns the buffer that holds the result.  INV conjugates every twiddle.
//
// Pass f (radix r, s = product of the radices before it, l = M / (r s), nbf = M / r
// butterflies): butterfly idx = j s + q reads x[idx + k nbf] (k < r; s l == nbf) and writes
// y[idx + j s (r - 1) + k s] after a twist by tw[k j tstep].  In the last pass l == 1, so
// j == 0 and every twist is 1: LAST drops the loads and the multiplies.
template <bool INV, bool LAST>
__device__ __forceinline__ void pass_radix4(const float2* __restrict__ x, float2* __restrict__ y,
                                            const float2* __restrict__ tw, int nbf, int s,
                                            int tstep, unsigned smagic) {
  const int sr1 = 3 * s;
  for (int idx = threadIdx.x; idx < nbf; idx += blockDim.x) {
    const int j = s == 1 ? idx : (int)__umulhi((unsigned)idx, smagic);
    const float2* xp = x + idx;
    const float2 a = xp[0];
    const float2 b = xp[nbf];
    const float2 c = xp[2 * nbf];
    const float2 d = xp[3 * nbf];
    const float2 apc = cadd(a, c), amc = csub(a, c);
    const float2 bpd = cadd(b, d), bmd = csub(b, d);
    // forward: -i*(b-d); inverse: +i*(b-d)
    const float2 jb = INV ? make_float2(-bmd.y, bmd.x) : make_float2(bmd.y, -bmd.x);
    float2* o = y + idx + j * sr1;
    if (LAST) {
      o[0] = cadd(apc, bpd);
      o[s] = cadd(amc, jb);
      o[2 * s] = csub(apc, bpd);
      o[3 * s] = csub(amc, jb);
    } else {
      const int jt = j * tstep;
      float2 w1 = __ldg(&tw[jt]);
      float2 w2 = __ldg(&tw[2 * jt]);
      float2 w3 = __ldg(&tw[3 * jt]);
      if (INV) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
      o[0] = cadd(apc, bpd);
      o[s] = cmul(cadd(amc, jb), w1);
      o[2 * s] = cmul(csub(apc, bpd), w2);
      o[3 * s] = cmul(csub(amc, jb), w3);
    }
  }
}

template <bool INV, bool LAST>
__device__ __forceinline__ void pass_radix2(const float2* __restrict__ x, float2* __restrict__ y,
                                            const float2* __restrict__ tw, int nbf, int s,
                                            int tstep, unsigned smagic) {
  for (int idx = threadIdx.x; idx < nbf; idx += blockDim.x) {
    const int j = s == 1 ? idx : (int)__umulhi((unsigned)idx, smagic);
    const float2 a = x[idx];
    const float2 b = x[idx + nbf];
    float2* o = y + idx + j * s;
    o[0] = cadd(a, b);
    if (LAST) {
      o[s] = csub(a, b);
    } else {
      float2 w1 = __ldg(&tw[j * tstep]);
      if (INV) w1.y = -w1.y;
      o[s] = cmul(csub(a, b), w1);
    }
  }
}

// 5-point DFT from the symmetric / antisymmetric pairs (x1 +- x4, x2 +- x3)
template <bool INV, bool LAST>
__device__ __forceinline__ void pass_radix5(const float2* __restrict__ x, float2* __restrict__ y,
                                            const float2* __restrict__ tw, int nbf, int s,
                                            int tstep, unsigned smagic) {
  const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;   // cos(2pi/5), cos(4pi/5)
  const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;    // sin(2pi/5), sin(4pi/5)
  const int sr1 = 4 * s;
  for (int idx = threadIdx.x; idx < nbf; idx += blockDim.x) {
    const int j = s == 1 ? idx : (int)__umulhi((unsigned)idx, smagic);
    const float2* xp = x + idx;
    const float2 x0 = xp[0];
    const float2 x1 = xp[nbf];
    const float2 x2 = xp[2 * nbf];
    const float2 x3 = xp[3 * nbf];
    const float2 x4 = xp[4 * nbf];
    const float2 t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
    const float2 a1 = make_float2(x0.x + c1 * t1.x + c2 * t2.x, x0.y + c1 * t1.y + c2 * t2.y);
    const float2 a2 = make_float2(x0.x + c2 * t1.x + c1 * t2.x, x0.y + c2 * t1.y + c1 * t2.y);
    const float2 b1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
    const float2 b2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
    // forward: y1 = a1 - i b1, y4 = a1 + i b1, y2 = a2 - i b2, y3 = a2 + i b2; inverse: conjugate roots
    const float2 ib1 = INV ? make_float2(-b1.y, b1.x) : make_float2(b1.y, -b1.x);
    const float2 ib2 = INV ? make_float2(-b2.y, b2.x) : make_float2(b2.y, -b2.x);
    float2* o = y + idx + j * sr1;
    o[0] = cadd(x0, cadd(t1, t2));
    if (LAST) {
      o[s] = cadd(a1, ib1);
      o[2 * s] = cadd(a2, ib2);
      o[3 * s] = csub(a2, ib2);
      o[4 * s] = csub(a1, ib1);
    } else {
      const int jt = j * tstep;
      float2 w1 = __ldg(&tw[jt]);
      float2 w2 = __ldg(&tw[2 * jt]);
      float2 w3 = __ldg(&tw[3 * jt]);
      float2 w4 = __ldg(&tw[4 * jt]);
      if (INV) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; w4.y = -w4.y; }
      o[s] = cmul(cadd(a1, ib1), w1);
      o[2 * s] = cmul(cadd(a2, ib2), w2);
      o[3 * s] = cmul(csub(a2, ib2), w3);
      o[4 * s] = cmul(csub(a1, ib1), w4);
    }
  }
}

template <bool INV, bool LAST>
__device__ __forceinline__ void pass_radix3(const float2* __restrict__ x, float2* __restrict__ y,
                                            const float2* __restrict__ tw, int nbf, int s,
                                            int tstep, unsigned smagic) {
  const float h3 = 0.86602540378443865f;   // sin(2pi/3)
  const int sr1 = 2 * s;
  for (int idx = threadIdx.x; idx < nbf; idx += blockDim.x) {
    const int j = s == 1 ? idx : (int)__umulhi((unsigned)idx, smagic);
    const float2* xp = x + idx;
    const float2 x0 = xp[0];
    const float2 x1 = xp[nbf];
    const float2 x2 = xp[2 * nbf];
    const float2 t1 = cadd(x1, x2), t3 = csub(x1, x2);
    const float2 a = make_float2(x0.x - 0.5f * t1.x, x0.y - 0.5f * t1.y);
    const float2 b = make_float2(h3 * t3.x, h3 * t3.y);
    const float2 ib = INV ? make_float2(-b.y, b.x) : make_float2(b.y, -b.x);
    float2* o = y + idx + j * sr1;
    o[0] = cadd(x0, t1);
    if (LAST) {
      o[s] = cadd(a, ib);
      o[2 * s] = csub(a, ib);
    } else {
      const int jt = j * tstep;
      float2 w1 = __ldg(&tw[jt]);
      float2 w2 = __ldg(&tw[2 * jt]);
      if (INV) { w1.y = -w1.y; w2.y = -w2.y; }
      o[s] = cmul(cadd(a, ib), w1);
      o[2 * s] = cmul(csub(a, ib), w2);
    }
  }
}

template <bool INV>
__device__ float2* stockham(float2* x, float2* y, const GeoK& g) {
  const float2* __restrict__ tw = g.tw;
  for (int f = 0; f < g.nfac; f++) {
    const int r = g.fac[f];
    const int l = g.pass_l[f];
    const int s = g.pass_s[f];
    const int tstep = g.pass_tstep[f];
    const int nbf = g.pass_nbf[f];               // butterflies of this pass: M / r
    const unsigned smagic = g.pass_smagic[f];    // j = idx / s without a division (idx < M <= 2^16)
    const bool last = l == 1;
    if (r == 4) {
      if (last) pass_radix4<INV, true>(x, y, tw, nbf, s, tstep, smagic);
      else pass_radix4<INV, false>(x, y, tw, nbf, s, tstep, smagic);
    } else if (r == 2) {
      if (last) pass_radix2<INV, true>(x, y, tw, nbf, s, tstep, smagic);
      else pass_radix2<INV, false>(x, y, tw, nbf, s, tstep, smagic);
    } else if (r == 5) {
      if (last) pass_radix5<INV, true>(x, y, tw, nbf, s, tstep, smagic);
      else pass_radix5<INV, false>(x, y, tw, nbf, s, tstep, smagic);
    } else if (r == 3) {
      if (last) pass_radix3<INV, true>(x, y, tw, nbf, s, tstep, smagic);
      else pass_radix3<INV, false>(x, y, tw, nbf, s, tstep, smagic);
    } else {
      // generic (odd prime) radix, in double with double roots of unity, rounded to float once.
      // A work item is a butterfly and a PAIR of outputs t, r - t: with S_u = x_u + x_{r-u},
      // D_u = x_u - x_{r-u} (u = 1 .. (r-1)/2) and w = root^(u t),
      //   y_t     = x_0 + sum_u (S_u Re w + i D_u Im w),   y_{r-t} = x_0 + sum_u (S_u Re w - i D_u Im w)
      // share their four real sums: (r-1)^2 real multiplies per butterfly instead of 4 r (r-1)
      // (the radix-167 pass of N = 3006 was half the time of the whole adaptive 44.1 kHz path).
      // Consecutive threads take consecutive butterflies of the SAME t, so a warp needs only a
      // few distinct roots per term.
      const double2* __restrict__ root = g.twg + g.twg_off[f];
      const unsigned nbfmagic = g.pass_nbfmagic[f];
      const int half = (r - 1) >> 1;
      const int items = nbf * (half + 1);
      for (int item = threadIdx.x; item < items; item += blockDim.x) {
        const int t = nbf == 1 ? item : (int)__umulhi((unsigned)item, nbfmagic);   // 0 .. half
        const int idx = item - t * nbf;                           // butterfly
        const int j = s == 1 ? idx : (int)__umulhi((unsigned)idx, smagic);
        const int q = idx - j * s;
        const float2* xb = x + q + s * j;
        const float2 v0 = xb[0];
        if (t == 0) {
          double ar = (double)v0.x, ai = (double)v0.y;
          for (int u = 1; u < r; u++) {
            const float2 v = xb[s * u * l];
            ar += (double)v.x;
            ai += (double)v.y;
          }
          y[q + s * (r * j)] = make_float2((float)ar, (float)ai);
        } else {
          double p1 = 0.0, p2 = 0.0, p3 = 0.0, p4 = 0.0;
          int ut = 0;
          for (int u = 1; u <= half; u++) {
            ut += t;
            if (ut >= r) ut -= r;
            double2 w = root[ut];
            if (INV) w.y = -w.y;
            const float2 va = xb[s * u * l];
            const float2 vb = xb[s * (r - u) * l];
            const double sx = (double)va.x + (double)vb.x, sy = (double)va.y + (double)vb.y;
            const double dx = (double)va.x - (double)vb.x, dy = (double)va.y - (double)vb.y;
            p1 = fma(sx, w.x, p1);
            p2 = fma(dy, w.y, p2);
            p3 = fma(sy, w.x, p3);
            p4 = fma(dx, w.y, p4);
          }
          const double ar = (double)v0.x + p1, ai = (double)v0.y + p3;
          const double y1r = ar - p2, y1i = ai + p4;              // output t
          const double y2r = ar + p2, y2i = ai - p4;              // output r - t
          float2 w1 = __ldg(&tw[j * t * tstep]);
          float2 w2 = __ldg(&tw[j * (r - t) * tstep]);
          if (INV) { w1.y = -w1.y; w2.y = -w2.y; }
          y[q + s * (r * j + t)] = make_float2((float)(y1r * (double)w1.x - y1i * (double)w1.y),
                                               (float)(y1r * (double)w1.y + y1i * (double)w1.x));
          y[q + s * (r * j + r - t)] = make_float2((float)(y2r * (double)w2.x - y2i * (double)w2.y),
                                                   (float)(y2r * (double)w2.y + y2i * (double)w2.x));
        }
      }
    }
    __syncthreads();
    float2* tmp = x; x = y; y = tmp;
  }
  return x;
}

// DFT of the M packed points in `a` (b: second buffer of the same size), either directly or,
// for an M with a large prime factor, as Bluestein's convolution over the power of two L:
//   Z[k] = w[k] * sum_n (z[n] w[n]) conj(w)[k - n],   w[n] = exp(-i pi n^2 / M)
// (forward; the unnormalised inverse conjugates w).  The convolution runs through two L-point
// Stockham transforms; bl_B = FFT_L(conj(w) wrapped to +-n) / L is real-symmetric in n, so the
// inverse's spectrum is conj(bl_B).  Returns the buffer holding the M results.
template <bool INV>
__device__ __forceinline__ float2* fft_points(float2* a, float2* b, const GeoK& g) {
  if (g.bl_L == 0) return stockham<INV>(a, b, g);
  const int M = g.M, L = g.bl_L;
  const float2* __restrict__ cw = g.bl_w;
  const float2* __restrict__ cB = g.bl_B;
  for (int n = threadIdx.x; n < L; n += blockDim.x) {
    float2 v = make_float2(0.f, 0.f);
    if (n < M) {
      float2 w = __ldg(&cw[n]);
      if (INV) w.y = -w.y;
      v = cmul(a[n], w);
    }
    a[n] = v;
  }
  __syncthreads();
  float2* A = stockham<false>(a, b, g);
  float2* other = (A == a) ? b : a;
  for (int k = threadIdx.x; k < L; k += blockDim.x) {
    float2 Bk = __ldg(&cB[k]);
    if (INV) Bk.y = -Bk.y;
    A[k] = cmul(A[k], Bk);
  }
  __syncthreads();
  float2* c = stockham<true>(A, other, g);
  for (int k = threadIdx.x; k < M; k += blockDim.x) {
    float2 w = __ldg(&cw[k]);
    if (INV) w.y = -w.y;
    c[k] = cmul(c[k], w);
  }
  __syncthreads();
  return c;
}

// ---------------------------------------------------------------------------
// forward: frame f = xbuf[f*hop, f*hop+frame) -> Hann -> centre in N -> rFFT
//
// VEC (geometries with cp % 4 == 0, frame % 4 == 0, hop even, M % 4 == 0 -- 48 kHz / 50 ms,
// 96 kHz / 77 ms ...): sample pairs (forward) and bin / sample quads (inverse) move as 64 /
// 128-bit accesses.  Same arithmetic per element in both forms.
// ---------------------------------------------------------------------------
// one bin of the real spectrum from the packed transform Z (k in [0, Kp))
__device__ __forceinline__ void split_bin(const float2* Z, int M, const float2* __restrict__ twn, int k,
                                          float* xr_out, float* xi_out) {
  float xr = 0.f, xi = 0.f;
  if (k == 0) {
    xr = Z[0].x + Z[0].y;
  } else if (k == M) {
    xr = Z[0].x - Z[0].y;
  } else if (k < M) {
    const float2 z1 = Z[k];
    const float2 z2 = Z[M - k];
    const float er = 0.5f * (z1.x + z2.x), ei = 0.5f * (z1.y - z2.y);
    const float dr = 0.5f * (z1.x - z2.x), di = 0.5f * (z1.y + z2.y);
    // O = D / i = (di, -dr);  X = E + w_N^k * O
    const float2 w = __ldg(&twn[k]);
    xr = er + (di * w.x + dr * w.y);
    xi = ei + (di * w.y - dr * w.x);
  }
  *xr_out = xr;
  *xi_out = xi;
}

template <bool VEC>
__global__ void __launch_bounds__(kFftMaxThreads)
forward_kernel(const GeoK g, const float* __restrict__ xbuf, float* __restrict__ Xre,
               float* __restrict__ Xim, float* __restrict__ P) {
  extern __shared__ __align__(16) float2 smem[];
  float2* a = smem;
  float2* b = smem + g.fft_len;
  const int f = blockIdx.x;
  const float* __restrict__ src = xbuf + (size_t)f * g.hop;
  const float* __restrict__ win = g.win;
  // fft_load_input_samples + INPUT window (fft_transform.c:165-182, stft_windows.c:120-127)
  // four rounds of loads are requested before the first is used (one round per DRAM / L2
  // latency otherwise: the runtime trip count keeps ptxas from batching them itself)
  for (int k0 = threadIdx.x; k0 < g.M; k0 += 4 * blockDim.x) {
    float v0[4], v1[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i0 = 2 * (k0 + u * (int)blockDim.x) - g.cp;
      if (VEC) {   // i0 even, frame even: the pair is inside the frame or outside it as a whole
        float2 sv = make_float2(0.f, 0.f), wv = make_float2(0.f, 0.f);
        if (i0 >= 0 && i0 < g.frame) {
          sv = *reinterpret_cast<const float2*>(src + i0);
          wv = __ldg(reinterpret_cast<const float2*>(win + i0));
        }
        v0[u] = sv.x * wv.x;
        v1[u] = sv.y * wv.y;
      } else {
        const int i1 = i0 + 1;
        v0[u] = (i0 >= 0 && i0 < g.frame) ? src[i0] * __ldg(&win[i0]) : 0.f;
        v1[u] = (i1 >= 0 && i1 < g.frame) ? src[i1] * __ldg(&win[i1]) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int k = k0 + u * (int)blockDim.x;
      if (k < g.M) a[k] = make_float2(v0[u], v1[u]);
    }
  }
  __syncthreads();
  const float2* Z = fft_points<false>(a, b, g);
  // split into the real spectrum; halfcomplex bin k <-> (Xre[k], Xim[k])
  float* __restrict__ re = Xre + (size_t)f * g.Kp;
  float* __restrict__ im = Xim + (size_t)f * g.Kp;
  float* __restrict__ pw = P + (size_t)f * g.Kp;
  const int M = g.M;
  // (one bin per thread also in the VEC form: four consecutive bins per thread put the Z
  // reads 32 B apart, an 8-way bank conflict that cost more than the 128-bit stores saved)
  for (int k = threadIdx.x; k < g.Kp; k += blockDim.x) {
    float xr, xi;
    split_bin(Z, M, g.twn, k, &xr, &xi);
    re[k] = xr;
    im[k] = xi;
    // spectral_features.c:80-107 (DC and Nyquist have xi == 0 exactly)
    pw[k] = fmaf(xr, xr, xi * xi);
  }
}

// ---------------------------------------------------------------------------
// inverse: Y -> unnormalised HC2R -> OUTPUT window / scale -> centre crop
// ---------------------------------------------------------------------------
// packed input point k of the inverse transform from bins k and M - k
__device__ __forceinline__ float2 merge_bins(float xr, float xi, float yr, float yi, float2 w) {
  const float er = xr + yr, ei = xi + yi;
  const float dr = xr - yr, di = xi - yi;
  // t = conj(w) * D ; Z = E + i*t
  const float tr = dr * w.x + di * w.y;
  const float ti = di * w.x - dr * w.y;
  return make_float2(er - ti, ei + tr);
}

template <bool VEC>
__global__ void __launch_bounds__(kFftMaxThreads)
inverse_kernel(const GeoK g, const float* __restrict__ Yre, const float* __restrict__ Yim,
               float* __restrict__ synth) {
  extern __shared__ __align__(16) float2 smem[];
  float2* a = smem;
  float2* b = smem + g.fft_len;
  const int f = blockIdx.x;
  const int M = g.M;
  const float* __restrict__ re = Yre + (size_t)f * g.Kp;
  const float* __restrict__ im = Yim + (size_t)f * g.Kp;
  if (VEC) {
    // quads of bins k .. k+3 (128-bit) against the mirrored bins M-k .. M-k-3 (scalar: they
    // sit at odd offsets), two quads per thread requested before the first is used
    for (int k0 = 4 * threadIdx.x; k0 < M; k0 += 8 * blockDim.x) {
      float4 r4[2], i4[2], w01[2], w23[2];
      float yr[2][4], yi[2][4];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int k = k0 + 4 * h * (int)blockDim.x;
        const bool in = k < M;
        r4[h] = in ? *reinterpret_cast<const float4*>(re + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        i4[h] = in ? *reinterpret_cast<const float4*>(im + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        w01[h] = in ? __ldg(reinterpret_cast<const float4*>(g.twn + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
        w23[h] = in ? __ldg(reinterpret_cast<const float4*>(g.twn + k + 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; u++) {
          yr[h][u] = in ? re[M - k - u] : 0.f;
          yi[h][u] = (in && k + u != 0) ? -im[M - k - u] : 0.f;   // conj X[M-k]
        }
      }
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int k = k0 + 4 * h * (int)blockDim.x;
        if (k < M) {
          const float xi0 = (k == 0) ? 0.f : i4[h].x;
          const float2 z0 = merge_bins(r4[h].x, xi0, yr[h][0], yi[h][0], make_float2(w01[h].x, w01[h].y));
          const float2 z1 = merge_bins(r4[h].y, i4[h].y, yr[h][1], yi[h][1], make_float2(w01[h].z, w01[h].w));
          const float2 z2 = merge_bins(r4[h].z, i4[h].z, yr[h][2], yi[h][2], make_float2(w23[h].x, w23[h].y));
          const float2 z3 = merge_bins(r4[h].w, i4[h].w, yr[h][3], yi[h][3], make_float2(w23[h].z, w23[h].w));
          *reinterpret_cast<float4*>(a + k) = make_float4(z0.x, z0.y, z1.x, z1.y);
          *reinterpret_cast<float4*>(a + k + 2) = make_float4(z2.x, z2.y, z3.x, z3.y);
        }
      }
    }
  } else {
    for (int k0 = threadIdx.x; k0 < M; k0 += 4 * blockDim.x) {   // loads batched as in forward_kernel
      float xr[4], xi[4], yr[4], yi[4];
      float2 w[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int k = k0 + u * (int)blockDim.x;
        const bool in = k < M;
        xr[u] = in ? re[k] : 0.f;
        xi[u] = (in && k != 0) ? im[k] : 0.f;
        yr[u] = in ? re[M - k] : 0.f;
        yi[u] = (in && k != 0) ? -im[M - k] : 0.f;   // conj X[M-k]
        w[u] = in ? __ldg(&g.twn[k]) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int k = k0 + u * (int)blockDim.x;
        if (k < M) a[k] = merge_bins(xr[u], xi[u], yr[u], yi[u], w[u]);
      }
    }
  }
  __syncthreads();
  const float2* z = fft_points<true>(a, b, g);
  // OUTPUT window then divide by scale (stft_windows.c:120-137), crop (fft_transform.c:184-201)
  float* __restrict__ out = synth + (size_t)f * g.frame_p;
  const float* __restrict__ win = g.win;
  const float* zf = reinterpret_cast<const float*>(z);
  // the quotient by the constant scale: reciprocal, one residual correction (the IEEE quotient
  // for every normal operand; the generic division sequence was 12 % of this kernel)
  const float scale = g.scale, inv_scale = 1.0f / scale;
  auto quot = [&](float num) {
    const float q0 = num * inv_scale;
    const float rem = fmaf(-q0, scale, num);
    return fmaf(rem, inv_scale, q0);
  };
  if (VEC) {
    for (int i = 4 * threadIdx.x; i < g.frame; i += 4 * blockDim.x) {
      const float4 v = *reinterpret_cast<const float4*>(zf + g.cp + i);
      const float4 w = __ldg(reinterpret_cast<const float4*>(win + i));
      *reinterpret_cast<float4*>(out + i) =
          make_float4(quot(v.x * w.x), quot(v.y * w.y), quot(v.z * w.z), quot(v.w * w.w));
    }
  } else {
#pragma unroll 4
    for (int i = threadIdx.x; i < g.frame; i += blockDim.x) out[i] = quot(zf[g.cp + i] * __ldg(&win[i]));
  }
}

// ---------------------------------------------------------------------------
// overlap-add, in the reference's summation order (stft_processor.c:162-170):
// the accumulator receives frames oldest first.  i runs over the n output
// samples of this call followed by the `frame` samples of the new tail.
// ---------------------------------------------------------------------------
template <typename I>
__global__ void ola_kernel(const GeoK g, const float* __restrict__ synth, int frames,
                           int carry, unsigned n, const float* __restrict__ tail_in,
                           float* __restrict__ tail_out, float* __restrict__ y) {
  // I = int when n + frame + carry < 2^31 (the usual case), long long otherwise: the two
  // divisions by hop per sample are what this kernel spends its time on
  const unsigned total = n + (unsigned)g.frame;
  const I hop = (I)g.hop, frame = (I)g.frame;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += gridDim.x * blockDim.x) {
    float acc = (i < (unsigned)g.frame) ? tail_in[i] : 0.f;
    const I ic = (I)i + (I)carry;
    I fmax = ic / hop - 1;
    if (fmax > (I)frames - 1) fmax = (I)frames - 1;
    const I lo = ic - frame + 1;
    I fmin = lo <= 0 ? 0 : (lo + hop - 1) / hop - 1;
    if (fmin < 0) fmin = 0;
    for (I f = fmin; f <= fmax; f++) {
      const I o = (f + 1) * hop - (I)carry;   // first output sample of frame f
      acc += synth[(size_t)f * g.frame_p + (size_t)((I)i - o)];
    }
    if (i < n) y[i] = acc;
    else tail_out[i - n] = acc;
  }
}

}  // namespace

// geometries whose frame / padding / hop let the FFT kernels move pairs and quads (see forward_kernel)
static bool fft_vec_ok(const GeoK& g) {
  return g.cp % 4 == 0 && g.frame % 4 == 0 && g.hop % 2 == 0 && g.M % 4 == 0;
}

static bool set_fft_smem(const void* fn, size_t bytes) {
  if (bytes > 48 * 1024) {
    SB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  }
  return true;
}

bool launch_forward(const GeoK& g, const float* xbuf, int frames, float* Xre, float* Xim,
                    float* P, cudaStream_t st) {
  KTimer kt_(kcForward, st);
  if (frames <= 0) return true;
  const size_t smem = 2 * sizeof(float2) * (size_t)g.fft_len;
  if (fft_vec_ok(g) && ((uintptr_t)xbuf & 7) == 0) {
    if (!set_fft_smem((const void*)forward_kernel<true>, smem)) return false;
    forward_kernel<true><<<frames, g.fft_threads, smem, st>>>(g, xbuf, Xre, Xim, P);
  } else {
    if (!set_fft_smem((const void*)forward_kernel<false>, smem)) return false;
    forward_kernel<false><<<frames, g.fft_threads, smem, st>>>(g, xbuf, Xre, Xim, P);
  }
  SB_LAUNCH_CHECK();
  return true;
}

bool launch_inverse(const GeoK& g, const float* Yre, const float* Yim, int frames,
                    float* synth, cudaStream_t st) {
  KTimer kt_(kcInverse, st);
  if (frames <= 0) return true;
  const size_t smem = 2 * sizeof(float2) * (size_t)g.fft_len;
  if (fft_vec_ok(g)) {
    if (!set_fft_smem((const void*)inverse_kernel<true>, smem)) return false;
    inverse_kernel<true><<<frames, g.fft_threads, smem, st>>>(g, Yre, Yim, synth);
  } else {
    if (!set_fft_smem((const void*)inverse_kernel<false>, smem)) return false;
    inverse_kernel<false><<<frames, g.fft_threads, smem, st>>>(g, Yre, Yim, synth);
  }
  SB_LAUNCH_CHECK();
  return true;
}

bool launch_ola(const GeoK& g, const float* synth, int frames, uint32_t carry, uint32_t n,
                const float* tail_in, float* tail_out, float* y, cudaStream_t st) {
  KTimer kt_(kcOla, st);
  const unsigned total = n + (unsigned)g.frame;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if ((unsigned long long)n + (unsigned long long)g.frame + carry + (unsigned long long)g.hop < 0x7fffffffull)
    ola_kernel<int><<<blocks, 256, 0, st>>>(g, synth, frames, (int)carry, n, tail_in, tail_out, y);
  else
    ola_kernel<long long><<<blocks, 256, 0, st>>>(g, synth, frames, (int)carry, n, tail_in, tail_out, y);
  SB_LAUNCH_CHECK();
  return true;
}

}  // namespace sb
