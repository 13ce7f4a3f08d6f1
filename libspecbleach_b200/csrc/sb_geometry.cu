// Geometry tables: everything derived from (sample_rate, frame_ms) once, on the
// host, and uploaded to the device.  Host arithmetic deliberately follows the
// reference's float expressions so that integer sizes and table entries agree.
#include "sb_engine.h"

#include <math.h>
#include <string.h>

#include <map>
#include <mutex>
#include <tuple>

namespace sb {

static std::mutex g_geo_mu;
static std::map<std::tuple<int, uint32_t, uint32_t>, Geometry*> g_geo_cache;

template <typename T>
static T* upload(Geometry* G, const std::vector<T>& v) {
  T* d = nullptr;
  if (dev_malloc(&d, v.size() * sizeof(T)) != cudaSuccess) return nullptr;
  if (cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) {
    dev_free(d);
    return nullptr;
  }
  G->dev_allocs.push_back(d);
  return d;
}

// Bark band upper edges in Hz (critical_bands.c:25-28)
static const float kBark[25] = {100.f,  200.f,  300.f,  400.f,  510.f,  630.f,  770.f,
                                920.f,  1080.f, 1270.f, 1480.f, 1720.f, 2000.f, 2320.f,
                                2700.f, 3150.f, 3700.f, 4400.f, 5300.f, 6400.f, 7700.f,
                                9500.f, 12000.f, 15500.f, 19000.f};

static void fill_passes(GeoK& g) {
  const int FL = g.fft_len;
  int len = FL, s = 1;
  for (int f = 0; f < g.nfac; f++) {
    const int r = g.fac[f];
    g.pass_l[f] = len / r;
    g.pass_s[f] = s;
    g.pass_tstep[f] = FL / len;
    g.pass_nbf[f] = FL / r;
    g.pass_smagic[f] = (unsigned)((0x100000000ull + (unsigned long long)s - 1) / (unsigned long long)s);
    g.pass_nbfmagic[f] = (unsigned)((0x100000000ull + (unsigned long long)g.pass_nbf[f] - 1) /
                                    (unsigned long long)g.pass_nbf[f]);
    len /= r;
    s *= r;
  }
  // CTA size of the FFT kernels: the passes have M / r butterflies each, which 256 threads
  // rarely divide (M = 1600: 400 and 320 -> every second round half empty).  Take the size
  // with the fewest warp rounds over all passes and the load / split loops, among those that
  // keep >= 24 warps resident per SM (2 x 8M bytes of shared memory and ~56 registers).
  double best = 1e30;
  g.fft_threads = 256;
  for (int T = 96; T <= 512; T += 8) {
    const int warps = (T + 31) / 32;
    double cost = 0.0;
    for (int f = 0; f < g.nfac; f++) {
      const int r = g.fac[f];
      const bool hard = r == 2 || r == 3 || r == 4 || r == 5;
      // generic radix: a work item is a butterfly and a pair of outputs, (r - 1) / 2 terms of
      // four double FMAs each (sb_stft.cu)
      const int items = hard ? g.pass_nbf[f] : g.pass_nbf[f] * ((r + 1) / 2);
      cost += (double)((items + T - 1) / T) * warps * (hard ? r : 2.0 * r);
    }
    cost += 2.0 * (double)((g.M + T - 1) / T) * warps * 2.0;      // load and split / crop loops
    cost += 0.25 * g.nfac * warps;                                 // barrier per pass
    if (g.bl_L) cost *= 2.0;                                       // both transforms of the convolution
    const size_t smem = 2 * sizeof(float2) * (size_t)FL;
    int ctas = (int)((227 * 1024) / (smem + 1024));
    const int by_regs = 65536 / (56 * warps * 32);
    if (by_regs < ctas) ctas = by_regs;
    if (ctas > 32) ctas = 32;
    if (ctas < 1) ctas = 1;
    const int resident = ctas * warps;
    if (resident < 24) cost *= 24.0 / resident;
    if (cost < best - 1e-9) { best = cost; g.fft_threads = T; }
  }
  // Bluestein lengths are 2^a, 3 2^a or 5 2^a: 256 threads divide every pass's butterfly count
  // (measured at L = 3072: 192 -> 0.278 ms per 4800 frames, 256 -> 0.255, 384 -> 0.255, 512 -> 0.274)
  if (g.bl_L >= 2048) g.fft_threads = 256;
  // measurement knob: SB_FFT_THREADS=<32..512> overrides the choice
  if (const char* e = getenv("SB_FFT_THREADS")) {
    const int t = atoi(e);
    if (t >= 32 && t <= 512) g.fft_threads = t;
  }
}

static void factorize(int m, int* fac, int* nfac) {
  int n = 0;
  while (m % 4 == 0) { fac[n++] = 4; m /= 4; }
  while (m % 2 == 0) { fac[n++] = 2; m /= 2; }
  while (m % 5 == 0) { fac[n++] = 5; m /= 5; }
  while (m % 3 == 0) { fac[n++] = 3; m /= 3; }
  for (int p = 7; p * p <= m; p += 2)
    while (m % p == 0) { fac[n++] = p; m /= p; }
  if (m > 1) fac[n++] = m;
  *nfac = n;
}

// Roots of unity in double for every pass whose radix is not 2 or 4: the generic
// butterfly accumulates r terms, so it runs in double and rounds once (keeps the
// odd frame sizes, e.g. N = 3006 = 2*3*3*167, as accurate as the 2^a 5^b ones).
static std::vector<double2> generic_roots(GeoK& g) {
  std::vector<double2> t;
  for (int f = 0; f < g.nfac; f++) {
    const int r = g.fac[f];
    if (r == 2 || r == 4) { g.twg_off[f] = -1; continue; }
    g.twg_off[f] = (int)t.size();
    for (int k = 0; k < r; k++) {
      long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)r;
      t.push_back(make_double2((double)cosl(a), (double)sinl(a)));
    }
  }
  if (t.empty()) t.push_back(make_double2(1.0, 0.0));
  return t;
}

// Smallest length >= need of the form 2^a, 3 2^a or 5 2^a (a >= 4): the hard-wired radices only.
static int bluestein_length(int need) {
  int best = 0;
  for (int odd : {1, 3, 5}) {
    int L = odd * 16;
    while (L < need) L <<= 1;
    if (best == 0 || L < best) best = L;
  }
  return best;
}

// Chooses the transform of the M packed points (mixed radix, or Bluestein over a power of two
// when M has a large prime factor), fills the pass tables and uploads the twiddles.
// SB_FFT_BLUESTEIN=0 / 1 forces the choice (measurement and tests).
static bool build_fft_plan(Geometry* G) {
  GeoK& g = G->k;
  const int L = bluestein_length(2 * g.M - 1);
  // Work of the generic passes (a pair of outputs per (r - 1) / 2 terms of four double FMAs, FP64
  // at half the FP32 rate: ~4 M r FP32-equivalent FMAs per pass of radix r) against the two
  // L-point transforms (~5 L log2 L), with the factor 2 that the measured kernels show between
  // the two forms (M = 1503: model 5.6x, kernel 2.6x).  M = 1503 (167), 907 (prime), 2605 (521):
  // Bluestein; 841 (29 x 29), 4810 (37): generic.
  double generic_work = 0.0;
  {
    int fac[14], nfac = 0;
    factorize(g.M, fac, &nfac);
    for (int f = 0; f < nfac; f++)
      if (fac[f] != 2 && fac[f] != 3 && fac[f] != 4 && fac[f] != 5) generic_work += 4.0 * g.M * fac[f];
  }
  bool bluestein = generic_work > 2.0 * 5.0 * L * log2((double)L);
  if (const char* e = getenv("SB_FFT_BLUESTEIN")) bluestein = *e == '1' ? true : (*e == '0' ? false : bluestein);
  if (2 * sizeof(float2) * (size_t)L > 200 * 1024 || L > 65536) bluestein = false;
  g.bl_L = bluestein ? L : 0;
  g.fft_len = bluestein ? L : g.M;
  factorize(g.fft_len, g.fac, &g.nfac);
  fill_passes(g);
  const int FL = g.fft_len;
  std::vector<float2> tw(FL);
  for (int k = 0; k < FL; k++) {
    long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)FL;
    tw[k] = make_float2((float)cosl(a), (float)sinl(a));
  }
  g.tw = upload(G, tw);
  g.twg = upload(G, generic_roots(g));
  if (!g.tw || !g.twg) return false;
  if (bluestein) {
    const int M = g.M;
    std::vector<float2> w(M);
    std::vector<double> cr(M), ci(M), cosL(L);
    for (int k = 0; k < L; k++)
      cosL[k] = (double)cosl(2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)L);
    for (int n = 0; n < M; n++) {
      const long long q = ((long long)n * n) % (2LL * M);      // n^2 mod 2M: the phase stays exact
      const long double a = -3.14159265358979323846264338327950288L * (long double)q / (long double)M;
      const double c = (double)cosl(a), sn = (double)sinl(a);
      w[n] = make_float2((float)c, (float)sn);
      cr[n] = c; ci[n] = -sn;                                  // b[+-n] = conj(w[n])
    }
    // b is symmetric in n, so its DFT is b[0] + 2 sum_{n >= 1} b[n] cos(2 pi n k / L): L M terms
    std::vector<float2> B(L);
    for (int k = 0; k < L; k++) {
      double sr = cr[0], si = ci[0];
      long long idx = 0;
      for (int n = 1; n < M; n++) {
        idx += k;
        if (idx >= L) idx -= L;
        sr += 2.0 * cr[n] * cosL[idx];
        si += 2.0 * ci[n] * cosL[idx];
      }
      B[k] = make_float2((float)(sr / L), (float)(si / L));
    }
    g.bl_w = upload(G, w);
    g.bl_B = upload(G, B);
    if (!g.bl_w || !g.bl_B) return false;
  }
  return true;
}

// spl_reference_value: absolute_hearing_thresholds.c:86-89,118-139.  The
// reference takes the peak of a float FFT of a Vorbis-windowed 1 kHz sine; the
// peak lies in the main lobe, so a double-precision DFT over the bins around it
// gives the same maximum (to ~1e-7 relative).
static float compute_spl_reference(uint32_t sr, int N) {
  std::vector<float> x(N);
  const float pi = 3.14159265358979323846f;
  for (int k = 0; k < N; k++) {
    float s = 1.f * sinf((2.f * pi * (float)k * 1000.f) / (float)sr);
    float p = (float)k / (float)N;
    float w = sinf(pi / 2.f * powf(sinf(pi * p), 2.f));   // spectral_utils.c:44-47
    if (!isnormal(w)) w = 0.f;
    x[k] = s * w;
  }
  int k0 = (int)lround(1000.0 * (double)N / (double)sr);
  int lo = k0 - 24 < 0 ? 0 : k0 - 24;
  int hi = k0 + 24 > N / 2 ? N / 2 : k0 + 24;
  double best = 0.0;
  for (int k = lo; k <= hi; k++) {
    double re = 0.0, im = 0.0;
    const double w = -2.0 * M_PI * (double)k / (double)N;
    for (int n = 0; n < N; n++) {
      double a = w * (double)n;
      re += x[n] * cos(a);
      im += x[n] * sin(a);
    }
    double p = (k == 0 || k == N / 2) ? re * re : re * re + im * im;
    if (p > best) best = p;
  }
  return 90.f - (10.f * log10f((float)best + kSpectralEps));
}

const Geometry* get_geometry(uint32_t sample_rate, float frame_ms) {
  int dev = 0;
  cudaGetDevice(&dev);
  uint32_t ms_bits;
  memcpy(&ms_bits, &frame_ms, 4);
  std::lock_guard<std::mutex> lock(g_geo_mu);
  auto key = std::make_tuple(dev, sample_rate, ms_bits);
  auto it = g_geo_cache.find(key);
  if (it != g_geo_cache.end()) return it->second;

  Geometry* G = new Geometry();
  G->sample_rate = sample_rate;
  G->frame_ms = frame_ms;
  GeoK& g = G->k;
  memset(&g, 0, sizeof(g));
  g.sr = (int)sample_rate;
  // stft_processor.c:57-58,68 ; fft_transform.c:59,119-123
  const uint32_t frame = (uint32_t)((frame_ms / 1000.f) * (float)sample_rate);
  if (frame < 4) { delete G; return nullptr; }
  uint32_t N = frame + kZeroPad;
  N += (N & 1u);
  g.frame = (int)frame;
  g.N = (int)N;
  g.hop = (int)(frame / kOverlap);
  g.K = (int)(N / 2 + 1);
  g.Kp = (g.K + 3) & ~3;
  g.M = (int)(N / 2);
  g.cp = (int)(N / 2 - frame / 2);
  g.frame_p = (g.frame + 3) & ~3;

  // periodic Hann over `frame` (spectral_utils.c:34-37, general_utils.c:26-31)
  const float pi = 3.14159265358979323846f;
  std::vector<float> win(frame);
  for (uint32_t k = 0; k < frame; k++) {
    float p = (float)k / (float)frame;
    float v = 0.5f - (0.5f * cosf(2.f * pi * p));
    if (!isnormal(v)) v = 0.f;
    win[k] = v;
  }
  // stft_windows.c:103-118
  float sum = 0.f;
  for (uint32_t k = 0; k < frame; k++) sum += win[k] * win[k];
  g.scale = (sum < 1.1920929e-7f) ? 1.f : sum * ((float)N / (float)g.hop);

  // twiddles of the real <-> packed complex split (the transform's own: build_fft_plan)
  std::vector<float2> twn(g.M + 1);
  for (int k = 0; k <= g.M; k++) {
    long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)g.N;
    twn[k] = make_float2((float)cosl(a), (float)sinl(a));
  }

  // Bark bands (critical_bands.c:96-116,138-149; quirk: bin edges use K as fft size)
  const float nyq = (float)sample_rate / 2.f;
  int last_valid = 0;
  for (int i = 0; i < 25; i++)
    if (kBark[i] < nyq) last_valid = i;
  g.nb = last_valid + 1;
  {
    std::vector<uint32_t> delim(g.nb), nbins(g.nb);
    for (int b = 0; b < g.nb; b++) {
      uint32_t bin = (uint32_t)((kBark[b] / ((float)sample_rate / (float)g.K)) + 0.5f);
      if (b == 0) { nbins[b] = bin; delim[b] = bin; }
      else if (b == g.nb - 1) { delim[b] = (uint32_t)g.K; nbins[b] = delim[b] - delim[b - 1]; }
      else { nbins[b] = bin - delim[b - 1]; delim[b] = bin; }
    }
    for (int b = 0; b < g.nb; b++) {
      uint32_t start = delim[b] - nbins[b], end = delim[b];
      if (end > (uint32_t)g.K) end = (uint32_t)g.K;        // defensive (tiny frames)
      if (start > end) start = end;
      g.band_start[b] = (int)start;
      g.band_start[b + 1] = (int)end;
      g.centre[b] = (float)start + ((float)(end - start) / 2.0f);   // masking_veto.c:103-110
    }
  }
  // band-sum tasks: slices of at most task_bins bins, bands in order
  {
    int task_bins = 128;
    for (;;) {
      int n = 0;
      for (int b = 0; b < g.nb; b++) {
        const int w = g.band_start[b + 1] - g.band_start[b];
        n += w > 0 ? (w + task_bins - 1) / task_bins : 1;
      }
      if (n <= kMaxBandTasks) break;
      task_bins *= 2;
    }
    int t = 0;
    for (int b = 0; b < g.nb; b++) {
      g.band_task0[b] = (short)t;
      const int w = g.band_start[b + 1] - g.band_start[b];
      const int n = w > 0 ? (w + task_bins - 1) / task_bins : 1;
      for (int s2 = 0; s2 < n; s2++) {
        g.task_band[t] = (short)b;
        g.task_k[t] = g.band_start[b] + s2 * task_bins;
        t++;
      }
    }
    g.band_task0[g.nb] = (short)t;
    g.task_k[t] = g.band_start[g.nb];
    g.ntask = t;
  }
  // temporal masking decays (masking_estimator.c:146-160)
  const float hop_time = (float)g.N / (4.0f * (float)sample_rate);
  for (int b = 0; b < g.nb; b++) {
    const float bark = fminf((float)b, 24.0f);
    const float weight = bark / 24.0f;
    const float tau = ((1.0f - weight) * 0.100f) + (weight * 0.025f);
    const float fwd = expf(-hop_time / tau);
    g.fwd_pow[b] = powf(fwd, kPowerLawExp);
  }
  g.bwd = expf(-hop_time / 0.010f);
  g.spl_ref = compute_spl_reference(sample_rate, g.N);

  // absolute threshold of hearing (absolute_hearing_thresholds.c:161-170, 141-159)
  std::vector<float> thr_spl(g.Kp, 0.f), thr_lin(g.Kp, 0.f);
  for (int k = 0; k < g.K; k++) {
    const float freq = fmaxf((float)k * ((float)sample_rate / (float)g.N), 20.f);
    const float fk = freq / 1000.f;
    thr_spl[k] = (3.64f * powf(fk, -0.8f)) - (6.5f * expf(-0.6f * powf(fk - 3.3f, 2.f))) +
                 (powf(10.f, -3.f) * powf(fk, 4.f));
    float lin = powf(10.f, (thr_spl[k] - g.spl_ref) / 10.f) - kSpectralEps;
    thr_lin[k] = lin < 0.f ? 0.f : lin;
  }
  std::vector<int> cband(g.Kp, 0), bob(g.Kp, 0);
  {
    int cur = 0;
    for (int k = 0; k < g.K; k++) {   // masking_veto.c:233-236
      while (cur < g.nb - 1 && (float)k > g.centre[cur + 1]) cur++;
      cband[k] = cur;
    }
    for (int b = 0; b < g.nb; b++)
      for (int k = g.band_start[b]; k < g.band_start[b + 1]; k++) bob[k] = b;
  }
  // masking_veto.c:238-250: the interpolation weight of a bin depends on the geometry only.
  // 0 where the reference takes the lower band alone (last band, coincident centres).
  std::vector<float> veto_t(g.Kp, 0.f);
  for (int k = 0; k < g.K; k++) {
    const int cb = cband[k];
    if (cb < g.nb - 1) {
      const float x0 = g.centre[cb], x1 = g.centre[cb + 1];
      if (x1 > x0) {
        const float t = ((float)k - x0) / (x1 - x0);
        veto_t[k] = fmaxf(0.0f, fminf(1.0f, t));
      }
    }
  }

  const bool plan_ok = build_fft_plan(G);
  g.win = upload(G, win);
  g.twn = upload(G, twn);
  g.abs_thr_spl = upload(G, thr_spl);
  g.abs_thr_lin = upload(G, thr_lin);
  g.cband = upload(G, cband);
  g.veto_t = upload(G, veto_t);
  g.band_of_bin = upload(G, bob);
  if (!plan_ok || !g.win || !g.tw || !g.twn || !g.twg || !g.abs_thr_spl || !g.abs_thr_lin || !g.cband ||
      !g.veto_t || !g.band_of_bin) {
    set_error("geometry upload", cudaGetLastError(), __FILE__, __LINE__);
    for (void* p : G->dev_allocs) dev_free(p);
    delete G;
    return nullptr;
  }
  g_geo_cache[key] = G;
  return G;
}

}  // namespace sb

namespace sb {

// Geometry for the FFT-only debug entry point: frame == N, no padding, unit window.
Geometry* make_fft_geometry(int fft_size) {
  if (fft_size < 4 || (fft_size & 1)) return nullptr;
  Geometry* G = new Geometry();
  GeoK& g = G->k;
  memset(&g, 0, sizeof(g));
  g.sr = 48000;
  g.frame = fft_size;
  g.N = fft_size;
  g.hop = fft_size;
  g.K = fft_size / 2 + 1;
  g.Kp = (g.K + 3) & ~3;
  g.M = fft_size / 2;
  g.cp = 0;
  g.frame_p = (g.frame + 3) & ~3;
  g.scale = 1.0f;
  std::vector<float> win(fft_size, 1.0f);
  std::vector<float2> twn(g.M + 1);
  for (int k = 0; k <= g.M; k++) {
    long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)g.N;
    twn[k] = make_float2((float)cosl(a), (float)sinl(a));
  }
  const bool plan_ok = build_fft_plan(G);
  g.win = upload(G, win);
  g.twn = upload(G, twn);
  if (!plan_ok || !g.win || !g.tw || !g.twn || !g.twg) {
    free_geometry(G);
    return nullptr;
  }
  return G;
}

void free_geometry(Geometry* G) {
  if (!G) return;
  for (void* p : G->dev_allocs) dev_free(p);
  delete G;
}

}  // namespace sb
