// ORACLE-ONLY (test infrastructure): thin C-ABI wrapper over the reference's own kaldi-native-fbank sources
// (/root/reference/ggml/examples/kaldi-native-fbank/csrc), compiled in place by oracle/Makefile into
// oracle/_ref/libknf_ref.so.  No reference source is copied; this file only calls the knf API
// (NumFrames feature-window.cc:76, ExtractWindow :121, FbankComputer::Compute feature-fbank.cc:73) with the options the
// reference hard-codes for this path (80 bins, 16 kHz, defaults otherwise; fairseq2.cpp:554-573).
#include <cstdint>
#include <vector>
#include "feature-fbank.h"
#include "feature-window.h"

extern "C" int knf_num_frames(int64_t num_samples) {
  knf::FrameExtractionOptions fo{};
  fo.samp_freq = 16000;
  return knf::NumFrames(num_samples, fo);
}

// wave: n samples (already multiplied by `scale` here), out: frames x 80 row-major. Returns #frames.
extern "C" int knf_fbank(const float* wave, int64_t n, float scale, float* out) {
  knf::MelBanksOptions mo{};
  mo.num_bins = 80;
  knf::FrameExtractionOptions fo{};
  fo.samp_freq = 16000;
  fo.dither = 0.0f;
  knf::FbankOptions opts{};
  opts.frame_opts = fo;
  opts.mel_opts = mo;
  std::vector<float> scaled(wave, wave + n);
  for (auto& v : scaled) v *= scale;
  int32_t nfr = knf::NumFrames(n, fo);
  knf::FbankComputer comp(opts);
  knf::FeatureWindowFunction win(comp.GetFrameOptions());
  std::vector<float> frame;
  for (int32_t f = 0; f < nfr; ++f) {
    frame.resize(0);
    knf::ExtractWindow(0, scaled.data(), n, f, fo, win, &frame);
    comp.Compute(0, 1.0f, &frame, out + (int64_t)f * 80);
  }
  return nfr;
}
