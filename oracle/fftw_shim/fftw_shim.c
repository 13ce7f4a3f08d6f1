/*
 * TEST INFRASTRUCTURE ONLY -- not part of the shipped product.
 *
 * Double-precision implementation of the FFTW3f calls used by the reference
 * (R2HC / HC2R r2r plans, fft_transform.c:98-103).  Layout and scaling follow
 * the FFTW manual: R2HC output is [r0, r1, ..., r_{n/2}, i_{(n+1)/2-1}, ..., i_1]
 * with forward sign exp(-2*pi*i*j*k/n); HC2R is the unnormalised inverse, so
 * HC2R(R2HC(x)) = n*x.
 *
 * Algorithm: mixed-radix Stockham autosort FFT (radix 4/2/3/5 butterflies and a
 * generic O(p^2) butterfly for other prime factors) on n/2 complex points plus
 * the usual real split/merge step; all arithmetic in double, one rounding to
 * float at the end.
 */
#include "fftw3.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  double re, im;
} cpx;

struct sb_shim_plan_s {
  int n;       /* real transform length                     */
  int m;       /* complex FFT length (n/2 if n even else n) */
  int even;    /* n even -> packed real algorithm           */
  int nfac;
  int fac[32];
  fftwf_r2r_kind kind;
  float* in;
  float* out;
  cpx* tw_m; /* exp(-2 pi i k / m), k < m */
  cpx* tw_n; /* exp(-2 pi i k / n), k <= n/2 (split step) */
  cpx* a;
  cpx* b;
};

void* fftwf_malloc(size_t n) {
  void* p = NULL;
  if (posix_memalign(&p, 64, n ? n : 64) != 0) {
    return NULL;
  }
  return p;
}

void fftwf_free(void* p) { free(p); }

static int factorize(int m, int* fac) {
  int nf = 0;
  while (m % 4 == 0) {
    fac[nf++] = 4;
    m /= 4;
  }
  while (m % 2 == 0) {
    fac[nf++] = 2;
    m /= 2;
  }
  for (int p = 3; p * p <= m; p += 2) {
    while (m % p == 0) {
      fac[nf++] = p;
      m /= p;
    }
  }
  if (m > 1) {
    fac[nf++] = m;
  }
  return nf;
}

fftwf_plan fftwf_plan_r2r_1d(int n, float* in, float* out, fftwf_r2r_kind kind,
                             unsigned flags) {
  (void)flags;
  if (n <= 0 || !in || !out) {
    return NULL;
  }
  fftwf_plan p = (fftwf_plan)calloc(1, sizeof(*p));
  if (!p) {
    return NULL;
  }
  p->n = n;
  p->even = (n % 2 == 0);
  p->m = p->even ? n / 2 : n;
  p->kind = kind;
  p->in = in;
  p->out = out;
  p->nfac = factorize(p->m, p->fac);
  p->tw_m = (cpx*)malloc(sizeof(cpx) * (size_t)p->m);
  p->tw_n = (cpx*)malloc(sizeof(cpx) * (size_t)(n / 2 + 1));
  p->a = (cpx*)malloc(sizeof(cpx) * (size_t)p->m);
  p->b = (cpx*)malloc(sizeof(cpx) * (size_t)p->m);
  if (!p->tw_m || !p->tw_n || !p->a || !p->b) {
    fftwf_destroy_plan(p);
    return NULL;
  }
  const long double two_pi = 6.283185307179586476925286766559L;
  for (int k = 0; k < p->m; k++) {
    long double ang = -two_pi * (long double)k / (long double)p->m;
    p->tw_m[k].re = (double)cosl(ang);
    p->tw_m[k].im = (double)sinl(ang);
  }
  for (int k = 0; k <= n / 2; k++) {
    long double ang = -two_pi * (long double)k / (long double)n;
    p->tw_n[k].re = (double)cosl(ang);
    p->tw_n[k].im = (double)sinl(ang);
  }
  return p;
}

void fftwf_destroy_plan(fftwf_plan p) {
  if (!p) {
    return;
  }
  free(p->tw_m);
  free(p->tw_n);
  free(p->a);
  free(p->b);
  free(p);
}

/* One Stockham DIF pass of radix r: current length len = r*l, stride s.
 * sign = -1 forward, +1 inverse (inverse conjugates the table). */
static void pass(const struct sb_shim_plan_s* P, int r, int len, int s,
                 const cpx* x, cpx* y, int sign) {
  const int l = len / r;
  const int M = P->m;
  const cpx* tw = P->tw_m;
  const int tstep = M / len; /* w_len^k = tw[k * tstep] */
  if (r == 2) {
    for (int j = 0; j < l; j++) {
      cpx w = tw[(size_t)j * tstep];
      if (sign > 0) w.im = -w.im;
      for (int q = 0; q < s; q++) {
        cpx a = x[q + s * j];
        cpx b = x[q + s * (j + l)];
        cpx d = {a.re - b.re, a.im - b.im};
        y[q + s * (2 * j)].re = a.re + b.re;
        y[q + s * (2 * j)].im = a.im + b.im;
        y[q + s * (2 * j + 1)].re = d.re * w.re - d.im * w.im;
        y[q + s * (2 * j + 1)].im = d.re * w.im + d.im * w.re;
      }
    }
    return;
  }
  if (r == 4) {
    for (int j = 0; j < l; j++) {
      cpx w1 = tw[(size_t)j * tstep];
      cpx w2 = tw[((size_t)2 * j * tstep) % M];
      cpx w3 = tw[((size_t)3 * j * tstep) % M];
      if (sign > 0) {
        w1.im = -w1.im;
        w2.im = -w2.im;
        w3.im = -w3.im;
      }
      for (int q = 0; q < s; q++) {
        cpx a = x[q + s * j];
        cpx b = x[q + s * (j + l)];
        cpx c = x[q + s * (j + 2 * l)];
        cpx d = x[q + s * (j + 3 * l)];
        cpx apc = {a.re + c.re, a.im + c.im};
        cpx amc = {a.re - c.re, a.im - c.im};
        cpx bpd = {b.re + d.re, b.im + d.im};
        cpx bmd = {b.re - d.re, b.im - d.im};
        /* forward: -i*(b-d) ; inverse: +i*(b-d) */
        cpx jbmd;
        if (sign < 0) {
          jbmd.re = bmd.im;
          jbmd.im = -bmd.re;
        } else {
          jbmd.re = -bmd.im;
          jbmd.im = bmd.re;
        }
        cpx t0 = {apc.re + bpd.re, apc.im + bpd.im};
        cpx t1 = {amc.re + jbmd.re, amc.im + jbmd.im};
        cpx t2 = {apc.re - bpd.re, apc.im - bpd.im};
        cpx t3 = {amc.re - jbmd.re, amc.im - jbmd.im};
        cpx* o = &y[q + s * (4 * j)];
        o[0] = t0;
        o[s].re = t1.re * w1.re - t1.im * w1.im;
        o[s].im = t1.re * w1.im + t1.im * w1.re;
        o[2 * s].re = t2.re * w2.re - t2.im * w2.im;
        o[2 * s].im = t2.re * w2.im + t2.im * w2.re;
        o[3 * s].re = t3.re * w3.re - t3.im * w3.im;
        o[3 * s].im = t3.re * w3.im + t3.im * w3.re;
      }
    }
    return;
  }
  /* generic radix (3, 5, any prime): O(r^2) butterfly */
  cpx v[r];
  const int rstep = M / r; /* w_r^k = tw[k * rstep] */
  for (int j = 0; j < l; j++) {
    for (int q = 0; q < s; q++) {
      for (int u = 0; u < r; u++) {
        v[u] = x[q + s * (j + u * l)];
      }
      for (int t = 0; t < r; t++) {
        double sr = v[0].re, si = v[0].im;
        for (int u = 1; u < r; u++) {
          cpx w = tw[((size_t)((u * t) % r)) * rstep];
          if (sign > 0) w.im = -w.im;
          sr += v[u].re * w.re - v[u].im * w.im;
          si += v[u].re * w.im + v[u].im * w.re;
        }
        cpx w = tw[((size_t)j * t * tstep) % M];
        if (sign > 0) w.im = -w.im;
        y[q + s * (r * j + t)].re = sr * w.re - si * w.im;
        y[q + s * (r * j + t)].im = sr * w.im + si * w.re;
      }
    }
  }
}

/* in-place (result returned in P->a) complex FFT of P->a, length m */
static void cfft(struct sb_shim_plan_s* P, int sign) {
  cpx* x = P->a;
  cpx* y = P->b;
  int len = P->m;
  int s = 1;
  for (int f = 0; f < P->nfac; f++) {
    const int r = P->fac[f];
    pass(P, r, len, s, x, y, sign);
    len /= r;
    s *= r;
    cpx* t = x;
    x = y;
    y = t;
  }
  if (x != P->a) {
    memcpy(P->a, x, sizeof(cpx) * (size_t)P->m);
  }
}

static void exec_r2hc(struct sb_shim_plan_s* P) {
  const int n = P->n;
  const float* in = P->in;
  float* out = P->out;
  if (!P->even) {
    for (int k = 0; k < n; k++) {
      P->a[k].re = in[k];
      P->a[k].im = 0.0;
    }
    cfft(P, -1);
    out[0] = (float)P->a[0].re;
    for (int k = 1; k <= n / 2; k++) {
      out[k] = (float)P->a[k].re;
      out[n - k] = (float)P->a[k].im;
    }
    return;
  }
  const int m = P->m;
  for (int k = 0; k < m; k++) {
    P->a[k].re = in[2 * k];
    P->a[k].im = in[2 * k + 1];
  }
  cfft(P, -1);
  const cpx* Z = P->a;
  out[0] = (float)(Z[0].re + Z[0].im);
  out[m] = (float)(Z[0].re - Z[0].im);
  for (int k = 1; k < m; k++) {
    /* E = (Z[k] + conj Z[m-k])/2, O = (Z[k] - conj Z[m-k])/(2i) */
    const double ar = Z[k].re, ai = Z[k].im;
    const double br = Z[m - k].re, bi = -Z[m - k].im;
    const double er = 0.5 * (ar + br), ei = 0.5 * (ai + bi);
    const double dr = 0.5 * (ar - br), di = 0.5 * (ai - bi);
    const double orr = di, oi = -dr; /* (dr + i di)/i */
    const cpx w = P->tw_n[k];
    const double xr = er + (orr * w.re - oi * w.im);
    const double xi = ei + (orr * w.im + oi * w.re);
    out[k] = (float)xr;
    out[n - k] = (float)xi;
  }
}

static void exec_hc2r(struct sb_shim_plan_s* P) {
  const int n = P->n;
  const float* in = P->in;
  float* out = P->out;
  if (!P->even) {
    P->a[0].re = in[0];
    P->a[0].im = 0.0;
    for (int k = 1; k <= n / 2; k++) {
      P->a[k].re = in[k];
      P->a[k].im = in[n - k];
      P->a[n - k].re = in[k];
      P->a[n - k].im = -(double)in[n - k];
    }
    cfft(P, +1);
    for (int k = 0; k < n; k++) {
      out[k] = (float)P->a[k].re;
    }
    return;
  }
  const int m = P->m;
  /* X[k], k = 0..m ; X[0], X[m] purely real */
  for (int k = 0; k < m; k++) {
    double xr, xi, yr, yi; /* X[k], conj X[m-k] */
    if (k == 0) {
      xr = in[0];
      xi = 0.0;
      yr = in[m];
      yi = 0.0;
    } else {
      xr = in[k];
      xi = in[n - k];
      yr = in[m - k];
      yi = -(double)in[n - (m - k)];
    }
    const double er = xr + yr, ei = xi + yi;
    const double dr = xr - yr, di = xi - yi;
    /* i * conj(w_n^k) * D */
    const cpx w = P->tw_n[k];
    const double tr = dr * w.re + di * w.im;
    const double ti = di * w.re - dr * w.im;
    P->a[k].re = er - ti;
    P->a[k].im = ei + tr;
  }
  cfft(P, +1);
  for (int k = 0; k < m; k++) {
    out[2 * k] = (float)P->a[k].re;
    out[2 * k + 1] = (float)P->a[k].im;
  }
}

void fftwf_execute(const fftwf_plan p) {
  if (!p) {
    return;
  }
  if (p->kind == FFTW_R2HC) {
    exec_r2hc(p);
  } else {
    exec_hc2r(p);
  }
}
