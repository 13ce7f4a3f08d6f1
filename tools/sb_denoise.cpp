// sb_denoise -- WAV in, WAV out through libspecbleach_b200.
//
// Plays the part of the reference's examples/denoiser_demo.c (libsndfile + a 512-sample
// block loop, mono only): learn the noise profile on the first --learn-seconds of the
// file (or restore one saved with --save-profile), then denoise the rest.  Differences:
// any channel count (one handle per channel, processed as a batch), 1D or 2D denoiser,
// PCM 16/24/32 and float32 WAV read and written in the input's own format, and the
// (de)interleaving / sample conversion happen on the GPU
// (specbleach_b200_process_interleaved).  Plain C ABI only.
//
//   sb_denoise [options] <noisy.wav> <denoised.wav>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "specbleach_2d_denoiser.h"
#include "specbleach_b200.h"
#include "specbleach_denoiser.h"

struct Wav {
  uint16_t format_tag = 0, channels = 0, bits = 0, block_align = 0;
  uint32_t sample_rate = 0;
  std::vector<unsigned char> header;   // everything up to and including the data chunk header
  std::vector<unsigned char> data;
  std::vector<unsigned char> trailer;  // chunks after the data chunk, copied through
};

static uint32_t rd32(const unsigned char* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

static bool read_wav(const char* path, Wav& w, std::string& err) {
  FILE* f = fopen(path, "rb");
  if (!f) { err = "cannot open input"; return false; }
  std::vector<unsigned char> all;
  unsigned char buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) all.insert(all.end(), buf, buf + n);
  fclose(f);
  if (all.size() < 12 || memcmp(all.data(), "RIFF", 4) != 0 || memcmp(all.data() + 8, "WAVE", 4) != 0) {
    err = "not a RIFF/WAVE file";
    return false;
  }
  size_t pos = 12;
  bool have_fmt = false;
  while (pos + 8 <= all.size()) {
    const uint32_t sz = rd32(&all[pos + 4]);
    if (memcmp(&all[pos], "fmt ", 4) == 0 && sz >= 16) {
      if (pos + 8 + 16 > all.size()) { err = "fmt chunk truncated"; return false; }
      const unsigned char* p = &all[pos + 8];
      w.format_tag = rd16(p);
      w.channels = rd16(p + 2);
      w.sample_rate = rd32(p + 4);
      w.block_align = rd16(p + 12);
      w.bits = rd16(p + 14);
      if (w.format_tag == 0xFFFE && sz >= 26) {                              // WAVE_FORMAT_EXTENSIBLE
        if (pos + 8 + 26 > all.size()) { err = "fmt chunk truncated"; return false; }
        w.format_tag = rd16(p + 24);
      }
      have_fmt = true;
    } else if (memcmp(&all[pos], "data", 4) == 0) {
      if (!have_fmt) { err = "data chunk before fmt chunk"; return false; }
      size_t avail = all.size() - (pos + 8);
      size_t len = sz <= avail ? sz : avail;
      w.header.assign(all.begin(), all.begin() + pos + 8);
      w.data.assign(all.begin() + pos + 8, all.begin() + pos + 8 + len);
      size_t after = pos + 8 + len + (len & 1);
      if (after < all.size()) w.trailer.assign(all.begin() + after, all.end());
      return true;
    }
    pos += 8 + (size_t)sz + (sz & 1);
  }
  err = "no data chunk";
  return false;
}

static bool write_wav(const char* path, const Wav& w, const std::vector<unsigned char>& data) {
  FILE* f = fopen(path, "wb");
  if (!f) return false;
  bool ok = fwrite(w.header.data(), 1, w.header.size(), f) == w.header.size() &&
            fwrite(data.data(), 1, data.size(), f) == data.size();
  if (ok && (data.size() & 1)) ok = fputc(0, f) != EOF;
  if (ok && !w.trailer.empty()) ok = fwrite(w.trailer.data(), 1, w.trailer.size(), f) == w.trailer.size();
  return (fclose(f) == 0) && ok;
}

static void usage(const char* prog) {
  fprintf(stderr,
          "Usage: %s [options] <noisy input.wav> <denoised output.wav>\n"
          "  --1d                   specbleach_denoiser (default: specbleach_2d_denoiser, NLM)\n"
          "  --frame-size <ms>      frame size (default 50)\n"
          "  --learn-seconds <s>    learn the noise profile on the first s seconds (default 1.0)\n"
          "  --learn-samples <n>    the same in samples (the demo's 8 blocks of 512: 4096)\n"
          "  --load-profile <file>  restore a saved profile instead of learning\n"
          "  --save-profile <file>  save the learned profile (channel 0) as the denoise pass used it\n"
          "  --reduction <dB>       (default 20)      --whitening <%%>   (default 50)\n"
          "  --smoothing <val>      1D: %%, 2D: NLM h (default 0 / 1)\n"
          "  --masking <0..1>       masking depth / NLM masking protection (default 0.5)\n"
          "  --strength <%%>         suppression strength (default 50)\n"
          "  --aggressiveness <-1..1> (default 1D: 1, 2D: 0)   --tonal <dB> (default 0)\n"
          "  --adaptive             adaptive noise estimation   --noise-method <0|1|2>\n"
          "  --residual             output the removed residual\n"
          "  --device <n>           CUDA device (default 0)\n",
          prog);
}

int main(int argc, char** argv) {
  bool two_d = true, adaptive = false, residual = false;
  float frame_ms = 50.f, learn_seconds = 1.f, reduction = 20.f, whitening = 50.f, smoothing = -1.f;
  float masking = 0.5f, strength = 50.f, aggr = 2.f, tonal = 0.f;
  int method = 0, device = 0;
  long learn_samples = -1;
  const char *load_profile = nullptr, *save_profile = nullptr;
  std::vector<const char*> pos;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto val = [&](float& dst) { if (i + 1 < argc) dst = (float)atof(argv[++i]); };
    if (a == "--1d") two_d = false;
    else if (a == "--frame-size") val(frame_ms);
    else if (a == "--learn-seconds") val(learn_seconds);
    else if (a == "--learn-samples" && i + 1 < argc) learn_samples = atol(argv[++i]);
    else if (a == "--reduction") val(reduction);
    else if (a == "--whitening") val(whitening);
    else if (a == "--smoothing") val(smoothing);
    else if (a == "--masking") val(masking);
    else if (a == "--strength") val(strength);
    else if (a == "--aggressiveness") val(aggr);
    else if (a == "--tonal") val(tonal);
    else if (a == "--adaptive") adaptive = true;
    else if (a == "--residual") residual = true;
    else if (a == "--noise-method" && i + 1 < argc) method = atoi(argv[++i]);
    else if (a == "--device" && i + 1 < argc) device = atoi(argv[++i]);
    else if (a == "--load-profile" && i + 1 < argc) load_profile = argv[++i];
    else if (a == "--save-profile" && i + 1 < argc) save_profile = argv[++i];
    else if (a == "--help" || a == "-h") { usage(argv[0]); return 0; }
    else if (a.rfind("--", 0) == 0) { fprintf(stderr, "unknown option %s\n", a.c_str()); usage(argv[0]); return 2; }
    else pos.push_back(argv[i]);
  }
  if (pos.size() != 2) { usage(argv[0]); return 2; }
  if (smoothing < 0.f) smoothing = two_d ? 1.f : 0.f;
  if (aggr > 1.5f) aggr = two_d ? 0.f : 1.f;

  Wav w;
  std::string err;
  if (!read_wav(pos[0], w, err)) { fprintf(stderr, "%s: %s\n", pos[0], err.c_str()); return 1; }
  int fmt;
  if (w.format_tag == 3 && w.bits == 32) fmt = SPECBLEACH_B200_PCM_F32;
  else if (w.format_tag == 1 && w.bits == 16) fmt = SPECBLEACH_B200_PCM_S16;
  else if (w.format_tag == 1 && w.bits == 24) fmt = SPECBLEACH_B200_PCM_S24;
  else if (w.format_tag == 1 && w.bits == 32) fmt = SPECBLEACH_B200_PCM_S32;
  else { fprintf(stderr, "unsupported WAV sample format (tag %u, %u bits)\n", w.format_tag, w.bits); return 1; }
  const uint32_t ch = w.channels;
  const size_t bps = w.bits / 8;
  if (ch == 0 || w.sample_rate == 0) { fprintf(stderr, "fmt chunk: no channels / sample rate\n"); return 1; }
  if (w.block_align != ch * bps) {
    fprintf(stderr, "fmt chunk: block align %u is not channels x bytes per sample (%u x %zu)\n",
            w.block_align, ch, bps);
    return 1;
  }
  if (w.data.size() / (bps * ch) > 0xffffffffull) { fprintf(stderr, "data chunk too long\n"); return 1; }
  const uint32_t frames = (uint32_t)(w.data.size() / (bps * ch));
  if (frames == 0) { fprintf(stderr, "empty audio\n"); return 1; }

  if (specbleach_b200_device_count() < 1 || !specbleach_b200_set_device(device)) {
    fprintf(stderr, "no CUDA device: %s\n", specbleach_b200_last_error());
    return 1;
  }
  std::vector<SpectralBleachHandle> hs(ch);
  for (uint32_t c = 0; c < ch; c++) {
    hs[c] = two_d ? specbleach_2d_initialize(w.sample_rate, frame_ms) : specbleach_initialize(w.sample_rate, frame_ms);
    if (!hs[c]) { fprintf(stderr, "initialize failed: %s\n", specbleach_b200_last_error()); return 1; }
  }
  auto load = [&](int learn) {
    bool ok = true;
    for (uint32_t c = 0; c < ch; c++) {
      if (two_d) {
        SpectralBleach2DDenoiserParameters p;
        memset(&p, 0, sizeof(p));
        p.learn_noise = learn; p.residual_listen = residual; p.reduction_amount = reduction;
        p.smoothing_factor = smoothing; p.whitening_factor = whitening; p.adaptive_noise = adaptive;
        p.noise_estimation_method = method; p.nlm_masking_protection = masking;
        p.suppression_strength = strength; p.aggressiveness = aggr; p.tonal_reduction = tonal;
        ok = specbleach_2d_load_parameters(hs[c], p) && ok;
      } else {
        SpectralBleachDenoiserParameters p;
        memset(&p, 0, sizeof(p));
        p.learn_noise = learn; p.residual_listen = residual; p.reduction_amount = reduction;
        p.smoothing_factor = smoothing; p.whitening_factor = whitening; p.adaptive_noise = adaptive;
        p.noise_estimation_method = method; p.masking_depth = masking;
        p.suppression_strength = strength; p.aggressiveness = aggr; p.tonal_reduction = tonal;
        ok = specbleach_load_parameters(hs[c], p) && ok;
      }
    }
    return ok;
  };

  std::vector<unsigned char> out(w.data.size());
  uint32_t learn_frames = 0;
  if (load_profile) {
    for (uint32_t c = 0; c < ch; c++)
      if (!specbleach_b200_profile_load(hs[c], load_profile)) {
        fprintf(stderr, "%s: %s\n", load_profile, specbleach_b200_last_error());
        return 1;
      }
  } else if (!adaptive || learn_seconds > 0.f) {
    learn_frames = learn_samples >= 0 ? (uint32_t)learn_samples
                                      : (uint32_t)(learn_seconds * (float)w.sample_rate);
    if (learn_frames > frames) learn_frames = frames;
  }
  if (learn_frames) {
    if (!load(1) || !specbleach_b200_process_interleaved(hs.data(), ch, learn_frames, fmt, w.data.data(), out.data())) {
      fprintf(stderr, "learn pass failed: %s\n", specbleach_b200_last_error());
      return 1;
    }
  }
  if (frames > learn_frames) {
    const size_t off = (size_t)learn_frames * ch * bps;
    if (!load(0) || !specbleach_b200_process_interleaved(hs.data(), ch, frames - learn_frames, fmt,
                                                         w.data.data() + off, out.data() + off)) {
      fprintf(stderr, "denoise pass failed: %s\n", specbleach_b200_last_error());
      return 1;
    }
  }
  // The file holds the profile as the denoise pass used it: the first denoised frame refines the
  // learned arrays (gap interpolation + smoothing, denoiser_profile_core.c:42-48), and a profile
  // restored with --load-profile is taken as is.  Saved before any frame was denoised (input not
  // longer than the learn segment) it would be the raw learned profile; the tool says so.
  if (save_profile && learn_frames) {
    if (frames <= learn_frames)
      fprintf(stderr, "note: %s holds the unrefined profile (nothing was denoised after learning)\n",
              save_profile);
    if (!specbleach_b200_profile_save(hs[0], save_profile))
      fprintf(stderr, "warning: could not save the profile: %s\n", specbleach_b200_last_error());
  }
  for (uint32_t c = 0; c < ch; c++) {
    if (two_d) specbleach_2d_free(hs[c]);
    else specbleach_free(hs[c]);
  }
  if (!write_wav(pos[1], w, out)) { fprintf(stderr, "cannot write %s\n", pos[1]); return 1; }
  fprintf(stderr, "%s: %u ch, %u Hz, %u frames (%.1f s), %s -> %s\n", two_d ? "2D" : "1D", ch,
          w.sample_rate, frames, (double)frames / w.sample_rate, pos[0], pos[1]);
  return 0;
}
