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
// powers of two (N = frame + 800), hence the runtime factor list.
#include "sb_common.h"

namespace sb {

namespace {

constexpr int kFftThreads = 256;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// Mixed-radix Stockham DIF passes over M complex points held in shared memory.
// Returns the buffer that holds the result.  INV conjugates every twiddle.
template <bool INV>
__device__ float2* stockham(float2* x, float2* y, const GeoK& g) {
  const int M = g.M;
  const float2* __restrict__ tw = g.tw;
  for (int f = 0; f < g.nfac; f++) {
    const int r = g.fac[f];
    const int l = g.pass_l[f];
    const int s = g.pass_s[f];
    const int tstep = g.pass_tstep[f];
    const int nbf = g.pass_nbf[f];               // butterflies of this pass: M / r
    const unsigned smagic = g.pass_smagic[f];    // j = idx / s without a division (idx < M <= 2^16)
    if (r == 4) {
      for (int idx = threadIdx.x; idx < nbf; idx += blockDim.x) {
        const int j = s == 1 ? idx : (int)__umulhi((unsigned)idx, smagic);
        const int q = idx - j * s;
        const float2 a = x[q + s * j];
        const float2 b = x[q + s * (j + l)];
        const float2 c = x[q + s * (j + 2 * l)];
        const float2 d = x[q + s * (j + 3 * l)];
        const float2 apc = cadd(a, c), amc = csub(a, c);
        const float2 bpd = cadd(b, d), bmd = csub(b, d);
        // forward: -i*(b-d); inverse: +i*(b-d)
        const float2 jb = INV ? make_float2(-bmd.y, bmd.x) : make_float2(bmd.y, -bmd.x);
        float2 w1 = __ldg(&tw[j * tstep]);
        float2 w2 = __ldg(&tw[2 * j * tstep]);
        float2 w3 = __ldg(&tw[3 * j * tstep]);
        if (INV) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
        float2* o = &y[q + s * (4 * j)];
        o[0] = cadd(apc, bpd);
        o[s] = cmul(cadd(amc, jb), w1);
        o[2 * s] = cmul(csub(apc, bpd), w2);
        o[3 * s] = cmul(csub(amc, jb), w3);
      }
    } else if (r == 2) {
      for (int idx = threadIdx.x; idx < nbf; idx += blockDim.x) {
        const int j = s == 1 ? idx : (int)__umulhi((unsigned)idx, smagic);
        const int q = idx - j * s;
        const float2 a = x[q + s * j];
        const float2 b = x[q + s * (j + l)];
        float2 w1 = __ldg(&tw[j * tstep]);
        if (INV) w1.y = -w1.y;
        y[q + s * (2 * j)] = cadd(a, b);
        y[q + s * (2 * j + 1)] = cmul(csub(a, b), w1);
      }
    } else if (r == 5) {
      // 5-point DFT from the symmetric / antisymmetric pairs (x1 +- x4, x2 +- x3)
      const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;   // cos(2pi/5), cos(4pi/5)
      const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;    // sin(2pi/5), sin(4pi/5)
      for (int idx = threadIdx.x; idx < nbf; idx += blockDim.x) {
        const int j = s == 1 ? idx : (int)__umulhi((unsigned)idx, smagic);
        const int q = idx - j * s;
        const float2 x0 = x[q + s * j];
        const float2 x1 = x[q + s * (j + l)];
        const float2 x2 = x[q + s * (j + 2 * l)];
        const float2 x3 = x[q + s * (j + 3 * l)];
        const float2 x4 = x[q + s * (j + 4 * l)];
        const float2 t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
        const float2 a1 = make_float2(x0.x + c1 * t1.x + c2 * t2.x, x0.y + c1 * t1.y + c2 * t2.y);
        const float2 a2 = make_float2(x0.x + c2 * t1.x + c1 * t2.x, x0.y + c2 * t1.y + c1 * t2.y);
        const float2 b1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
        const float2 b2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
        // forward: y1 = a1 - i b1, y4 = a1 + i b1, y2 = a2 - i b2, y3 = a2 + i b2; inverse: conjugate roots
        const float2 ib1 = INV ? make_float2(-b1.y, b1.x) : make_float2(b1.y, -b1.x);
        const float2 ib2 = INV ? make_float2(-b2.y, b2.x) : make_float2(b2.y, -b2.x);
        float2 w1 = __ldg(&tw[j * tstep]);
        float2 w2 = __ldg(&tw[2 * j * tstep]);
        float2 w3 = __ldg(&tw[3 * j * tstep]);
        float2 w4 = __ldg(&tw[4 * j * tstep]);
        if (INV) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; w4.y = -w4.y; }
        float2* o = &y[q + s * (5 * j)];
        o[0] = cadd(x0, cadd(t1, t2));
        o[s] = cmul(cadd(a1, ib1), w1);
        o[2 * s] = cmul(cadd(a2, ib2), w2);
        o[3 * s] = cmul(csub(a2, ib2), w3);
        o[4 * s] = cmul(csub(a1, ib1), w4);
      }
    } else if (r == 3) {
      const float h3 = 0.86602540378443865f;   // sin(2pi/3)
      for (int idx = threadIdx.x; idx < nbf; idx += blockDim.x) {
        const int j = s == 1 ? idx : (int)__umulhi((unsigned)idx, smagic);
        const int q = idx - j * s;
        const float2 x0 = x[q + s * j];
        const float2 x1 = x[q + s * (j + l)];
        const float2 x2 = x[q + s * (j + 2 * l)];
        const float2 t1 = cadd(x1, x2), t3 = csub(x1, x2);
        const float2 a = make_float2(x0.x - 0.5f * t1.x, x0.y - 0.5f * t1.y);
        const float2 b = make_float2(h3 * t3.x, h3 * t3.y);
        const float2 ib = INV ? make_float2(-b.y, b.x) : make_float2(b.y, -b.x);
        float2 w1 = __ldg(&tw[j * tstep]);
        float2 w2 = __ldg(&tw[2 * j * tstep]);
        if (INV) { w1.y = -w1.y; w2.y = -w2.y; }
        float2* o = &y[q + s * (3 * j)];
        o[0] = cadd(x0, t1);
        o[s] = cmul(cadd(a, ib), w1);
        o[2 * s] = cmul(csub(a, ib), w2);
      }
    } else {
      // generic radix: one work item per output point; the r-term sum runs in
      // double with double roots of unity and is rounded to float once
      // consecutive threads take consecutive butterflies of the SAME output index t, so a warp
      // needs only a few distinct roots per term (the root loads were the bottleneck of the
      // radix-167 pass of N = 3006 when every lane asked for a different one)
      const double2* __restrict__ root = g.twg + g.twg_off[f];
      const unsigned nbfmagic = g.pass_nbfmagic[f];
      for (int item = threadIdx.x; item < M; item += blockDim.x) {
        const int t = nbf == 1 ? item : (int)__umulhi((unsigned)item, nbfmagic);   // output index
        const int idx = item - t * nbf;                           // butterfly
        const int j = s == 1 ? idx : (int)__umulhi((unsigned)idx, smagic);
        const int q = idx - j * s;
        const float2 v0 = x[q + s * j];
        double ar = (double)v0.x, ai = (double)v0.y;
        int ut = 0;
        for (int u = 1; u < r; u++) {
          ut += t;
          if (ut >= r) ut -= r;
          double2 w = root[ut];
          if (INV) w.y = -w.y;
          const float2 v = x[q + s * (j + u * l)];
          ar = fma((double)v.x, w.x, fma(-(double)v.y, w.y, ar));
          ai = fma((double)v.x, w.y, fma((double)v.y, w.x, ai));
        }
        float2 wo = __ldg(&tw[j * t * tstep]);
        if (INV) wo.y = -wo.y;
        const double yr = ar * (double)wo.x - ai * (double)wo.y;
        const double yi = ar * (double)wo.y + ai * (double)wo.x;
        y[q + s * (r * j + t)] = make_float2((float)yr, (float)yi);
      }
    }
    __syncthreads();
    float2* tmp = x; x = y; y = tmp;
  }
  return x;
}

// ---------------------------------------------------------------------------
// forward: frame f = xbuf[f*hop, f*hop+frame) -> Hann -> centre in N -> rFFT
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kFftThreads)
forward_kernel(const GeoK g, const float* __restrict__ xbuf, float* __restrict__ Xre,
               float* __restrict__ Xim, float* __restrict__ P) {
  extern __shared__ float2 smem[];
  float2* a = smem;
  float2* b = smem + g.M;
  const int f = blockIdx.x;
  const float* __restrict__ src = xbuf + (size_t)f * g.hop;
  const float* __restrict__ win = g.win;
  // fft_load_input_samples + INPUT window (fft_transform.c:165-182, stft_windows.c:120-127)
  for (int k = threadIdx.x; k < g.M; k += blockDim.x) {
    const int i0 = 2 * k - g.cp;
    const int i1 = i0 + 1;
    float v0 = 0.f, v1 = 0.f;
    if (i0 >= 0 && i0 < g.frame) v0 = src[i0] * __ldg(&win[i0]);
    if (i1 >= 0 && i1 < g.frame) v1 = src[i1] * __ldg(&win[i1]);
    a[k] = make_float2(v0, v1);
  }
  __syncthreads();
  const float2* Z = stockham<false>(a, b, g);
  // split into the real spectrum; halfcomplex bin k <-> (Xre[k], Xim[k])
  float* __restrict__ re = Xre + (size_t)f * g.Kp;
  float* __restrict__ im = Xim + (size_t)f * g.Kp;
  float* __restrict__ pw = P + (size_t)f * g.Kp;
  const int M = g.M;
  for (int k = threadIdx.x; k < g.Kp; k += blockDim.x) {
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
      const float2 w = __ldg(&g.twn[k]);
      xr = er + (di * w.x + dr * w.y);
      xi = ei + (di * w.y - dr * w.x);
    }
    re[k] = xr;
    im[k] = xi;
    // spectral_features.c:80-107 (DC and Nyquist have xi == 0 exactly)
    pw[k] = fmaf(xr, xr, xi * xi);
  }
}

// ---------------------------------------------------------------------------
// inverse: Y -> unnormalised HC2R -> OUTPUT window / scale -> centre crop
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kFftThreads)
inverse_kernel(const GeoK g, const float* __restrict__ Yre, const float* __restrict__ Yim,
               float* __restrict__ synth) {
  extern __shared__ float2 smem[];
  float2* a = smem;
  float2* b = smem + g.M;
  const int f = blockIdx.x;
  const int M = g.M;
  const float* __restrict__ re = Yre + (size_t)f * g.Kp;
  const float* __restrict__ im = Yim + (size_t)f * g.Kp;
  for (int k = threadIdx.x; k < M; k += blockDim.x) {
    float xr = re[k], xi = (k == 0) ? 0.f : im[k];
    float yr = re[M - k], yi = (k == 0) ? 0.f : -im[M - k];   // conj X[M-k]
    const float er = xr + yr, ei = xi + yi;
    const float dr = xr - yr, di = xi - yi;
    const float2 w = __ldg(&g.twn[k]);
    // t = conj(w) * D ; Z = E + i*t
    const float tr = dr * w.x + di * w.y;
    const float ti = di * w.x - dr * w.y;
    a[k] = make_float2(er - ti, ei + tr);
  }
  __syncthreads();
  const float2* z = stockham<true>(a, b, g);
  // OUTPUT window then divide by scale (stft_windows.c:120-137), crop (fft_transform.c:184-201)
  float* __restrict__ out = synth + (size_t)f * g.frame_p;
  const float* __restrict__ win = g.win;
  const float* zf = reinterpret_cast<const float*>(z);
  for (int i = threadIdx.x; i < g.frame; i += blockDim.x) {
    const float v = zf[g.cp + i];
    out[i] = (v * __ldg(&win[i])) / g.scale;
  }
}

// ---------------------------------------------------------------------------
// overlap-add, in the reference's summation order (stft_processor.c:162-170):
// the accumulator receives frames oldest first.  i runs over the n output
// samples of this call followed by the `frame` samples of the new tail.
// ---------------------------------------------------------------------------
__global__ void ola_kernel(const GeoK g, const float* __restrict__ synth, int frames,
                           int carry, unsigned n, const float* __restrict__ tail_in,
                           float* __restrict__ tail_out, float* __restrict__ y) {
  const unsigned total = n + (unsigned)g.frame;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += gridDim.x * blockDim.x) {
    float acc = (i < (unsigned)g.frame) ? tail_in[i] : 0.f;
    const long long ic = (long long)i + carry;
    long long fmax = ic / g.hop - 1;
    if (fmax > frames - 1) fmax = frames - 1;
    long long lo = ic - g.frame + 1;
    long long fmin = lo <= 0 ? 0 : (lo + g.hop - 1) / g.hop - 1;
    if (fmin < 0) fmin = 0;
    for (long long f = fmin; f <= fmax; f++) {
      const long long o = (f + 1) * g.hop - carry;   // first output sample of frame f
      acc += synth[(size_t)f * g.frame_p + (size_t)(i - o)];
    }
    if (i < n) y[i] = acc;
    else tail_out[i - n] = acc;
  }
}

}  // namespace

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
  const size_t smem = 2 * sizeof(float2) * (size_t)g.M;
  if (!set_fft_smem((const void*)forward_kernel, smem)) return false;
  forward_kernel<<<frames, kFftThreads, smem, st>>>(g, xbuf, Xre, Xim, P);
  SB_LAUNCH_CHECK();
  return true;
}

bool launch_inverse(const GeoK& g, const float* Yre, const float* Yim, int frames,
                    float* synth, cudaStream_t st) {
  KTimer kt_(kcInverse, st);
  if (frames <= 0) return true;
  const size_t smem = 2 * sizeof(float2) * (size_t)g.M;
  if (!set_fft_smem((const void*)inverse_kernel, smem)) return false;
  inverse_kernel<<<frames, kFftThreads, smem, st>>>(g, Yre, Yim, frames ? synth : synth);
  SB_LAUNCH_CHECK();
  return true;
}

bool launch_ola(const GeoK& g, const float* synth, int frames, uint32_t carry, uint32_t n,
                const float* tail_in, float* tail_out, float* y, cudaStream_t st) {
  KTimer kt_(kcOla, st);
  const unsigned total = n + (unsigned)g.frame;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  ola_kernel<<<blocks, 256, 0, st>>>(g, synth, frames, (int)carry, n, tail_in, tail_out, y);
  SB_LAUNCH_CHECK();
  return true;
}

}  // namespace sb
