// sb_fbank: WaveformToFbank on device.
//   frame (400/160, snip_edges) -> x*2^15 -> remove DC -> pre-emphasis 0.97 -> Povey window -> zero-pad 512 ->
//   radix-2 FFT in shared memory -> |X|^2 (256 bins) -> 80 triangular mel filters -> log(max(.,eps))
//   -> per-utterance, per-bin standardisation over time (unbiased std) -> fp16.
// Arithmetic follows ggml/examples/kaldi-native-fbank/csrc (feature-window.cc:76-232, feature-fbank.cc:73-120,
// mel-computations.cc:107-257).  One warp per frame, four frames per CTA; the kernel is HBM-bound in principle
// (0.80 MB/utterance) and tiny next to the encoder.
#include <math.h>

#include <vector>

#include "common.cuh"

namespace sb {

constexpr int FR_LEN = 400, FR_SHIFT = 160, NFFT = 512, NBIN = 256, NMEL = 80;

__device__ float g_window[FR_LEN];
__device__ float2 g_twiddle[NFFT / 2];
__device__ float g_melw[NMEL * NBIN];
__device__ int g_mel_start[NMEL];
__device__ int g_mel_len[NMEL];
static bool g_tables_ready = false;

static double mel_scale(double f) { return 1127.0 * log(1.0 + f / 700.0); }

static int init_tables(cudaStream_t st) {
  if (g_tables_ready) return SB_OK;
  std::vector<float> win(FR_LEN);
  const double a = 2.0 * M_PI / (FR_LEN - 1);
  for (int i = 0; i < FR_LEN; ++i) win[i] = (float)pow(0.5 - 0.5 * cos(a * i), 0.85);
  std::vector<float2> tw(NFFT / 2);
  for (int i = 0; i < NFFT / 2; ++i) {
    double ang = -2.0 * M_PI * i / NFFT;
    tw[i] = make_float2((float)cos(ang), (float)sin(ang));
  }
  // mel-computations.cc:107-222: 80 bins, low 20 Hz, high = Nyquist, bins over n_fft/2 FFT bins
  std::vector<float> melw((size_t)NMEL * NBIN, 0.f);
  std::vector<int> ms(NMEL, 0), ml(NMEL, 0);
  const double sr = 16000.0, fft_bin_width = sr / NFFT;
  const double mlow = mel_scale(20.0), mhigh = mel_scale(0.5 * sr);
  const double delta = (mhigh - mlow) / (NMEL + 1);
  for (int b = 0; b < NMEL; ++b) {
    // knf computes in float; mirror its float arithmetic for the weights
    float left = (float)(mlow + b * delta), center = (float)(mlow + (b + 1) * delta), right = (float)(mlow + (b + 2) * delta);
    int first = -1, last = -1;
    for (int i = 0; i < NBIN; ++i) {
      float mel = (float)mel_scale(fft_bin_width * i);
      if (mel > left && mel < right) {
        float w = (mel <= center) ? (mel - left) / (center - left) : (right - mel) / (right - center);
        melw[(size_t)b * NBIN + i] = w;
        if (first < 0) first = i;
        last = i;
      }
    }
    ms[b] = first < 0 ? 0 : first;
    ml[b] = first < 0 ? 0 : last - first + 1;
  }
  SB_CUDA_OK(cudaMemcpyToSymbolAsync(g_window, win.data(), sizeof(float) * FR_LEN, 0, cudaMemcpyHostToDevice, st));
  SB_CUDA_OK(cudaMemcpyToSymbolAsync(g_twiddle, tw.data(), sizeof(float2) * NFFT / 2, 0, cudaMemcpyHostToDevice, st));
  SB_CUDA_OK(cudaMemcpyToSymbolAsync(g_melw, melw.data(), sizeof(float) * NMEL * NBIN, 0, cudaMemcpyHostToDevice, st));
  SB_CUDA_OK(cudaMemcpyToSymbolAsync(g_mel_start, ms.data(), sizeof(int) * NMEL, 0, cudaMemcpyHostToDevice, st));
  SB_CUDA_OK(cudaMemcpyToSymbolAsync(g_mel_len, ml.data(), sizeof(int) * NMEL, 0, cudaMemcpyHostToDevice, st));
  SB_CUDA_OK(cudaStreamSynchronize(st));  // host vectors go out of scope; one-time cost at first call
  g_tables_ready = true;
  return SB_OK;
}

__device__ __forceinline__ int num_frames_of(int n) { return n < FR_LEN ? 0 : 1 + (n - FR_LEN) / FR_SHIFT; }

__global__ void __launch_bounds__(128) fbank_frames_kernel(const float* __restrict__ wave, long long wave_ld,
                                                           const int* __restrict__ num_samples, float* __restrict__ work,
                                                           int frames_ld, int* __restrict__ frames_out) {
  __shared__ float2 xs[4][NFFT];
  __shared__ float pw[4][NBIN];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int f = blockIdx.x * 4 + warp;
  const int nfr = num_frames_of(num_samples[b]);
  if (blockIdx.x == 0 && threadIdx.x == 0 && frames_out != nullptr) frames_out[b] = nfr;
  if (f >= nfr) return;  // warp-uniform; no block-level sync is used below
  const float* src = wave + (long long)b * wave_ld + (long long)f * FR_SHIFT;
  float2* x = xs[warp];
  // load + scale, DC removal
  float v[13];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 13; ++i) {
    int idx = lane + 32 * i;
    v[i] = idx < FR_LEN ? src[idx] * 32768.0f : 0.f;
    sum += v[i];
  }
  const float mean = warp_sum(sum) / FR_LEN;
  // pre-emphasis needs the left neighbour: stage (x - mean) in shared memory
  float* tmp = reinterpret_cast<float*>(x);  // reuse as 1024 floats
#pragma unroll
  for (int i = 0; i < 13; ++i) {
    int idx = lane + 32 * i;
    if (idx < FR_LEN) tmp[idx] = v[i] - mean;
  }
  __syncwarp();
  float y[13];
#pragma unroll
  for (int i = 0; i < 13; ++i) {
    int idx = lane + 32 * i;
    if (idx < FR_LEN) {
      float cur = tmp[idx], prev = tmp[idx > 0 ? idx - 1 : 0];
      y[i] = (cur - 0.97f * prev) * g_window[idx];
    } else {
      y[i] = 0.f;
    }
  }
  __syncwarp();
  // bit-reversed scatter into the complex work array (zero-padded to 512)
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    int idx = lane + 32 * i;
    float val = (i < 13) ? y[i < 13 ? i : 12] : 0.f;
    if (idx >= FR_LEN) val = 0.f;
    int rev = __brev((unsigned)idx) >> (32 - 9);
    x[rev] = make_float2(val, 0.f);
  }
  __syncwarp();
  // 9 radix-2 stages, 256 butterflies each (8 per lane)
#pragma unroll 1
  for (int s = 1; s <= 9; ++s) {
    const int half = 1 << (s - 1), stride = NFFT >> s;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int j = lane + 32 * i;
      int pos = j & (half - 1);
      int k = ((j >> (s - 1)) << s) + pos;
      float2 w = g_twiddle[pos * stride];
      float2 u = x[k], t = x[k + half];
      float2 wt = make_float2(w.x * t.x - w.y * t.y, w.x * t.y + w.y * t.x);
      x[k] = make_float2(u.x + wt.x, u.y + wt.y);
      x[k + half] = make_float2(u.x - wt.x, u.y - wt.y);
    }
    __syncwarp();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int k = lane + 32 * i;
    float2 c = x[k];
    pw[warp][k] = c.x * c.x + c.y * c.y;
  }
  __syncwarp();
  float* dst = work + ((long long)b * frames_ld + f) * NMEL;
  for (int m = lane; m < NMEL; m += 32) {
    const int st = g_mel_start[m], len = g_mel_len[m];
    const float* w = g_melw + m * NBIN + st;
    float acc = 0.f;
    for (int i = 0; i < len; ++i) acc += w[i] * pw[warp][st + i];
    dst[m] = logf(fmaxf(acc, 1.1920928955078125e-07f));
  }
}

// per (utterance, bin): mean / unbiased std over frames; write fp16, zero the padding frames
__global__ void __launch_bounds__(256) fbank_standardize_kernel(const float* __restrict__ work, const int* __restrict__ num_samples,
                                                                elem_t* __restrict__ out, int frames_ld, int standardize) {
  const int b = blockIdx.x;
  const int nfr = num_frames_of(num_samples[b]);
  const float* src = work + (long long)b * frames_ld * NMEL;
  elem_t* dst = out + (long long)b * frames_ld * NMEL;
  __shared__ float s_mean[NMEL], s_rstd[NMEL];
  __shared__ double red[3][NMEL];  // 3 partial groups of 80 threads (240 of 256 threads active)
  const int bin = threadIdx.x % NMEL, grp = threadIdx.x / NMEL;
  if (standardize && nfr > 1) {
    double acc = 0.0;
    if (grp < 3)
      for (int f = grp; f < nfr; f += 3) acc += src[(long long)f * NMEL + bin];
    if (grp < 3) red[grp][bin] = acc;
    __syncthreads();
    if (threadIdx.x < NMEL) s_mean[threadIdx.x] = (float)((red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x]) / nfr);
    __syncthreads();
    acc = 0.0;
    if (grp < 3) {
      const float mu = s_mean[bin];
      for (int f = grp; f < nfr; f += 3) {
        float d = src[(long long)f * NMEL + bin] - mu;
        acc += (double)d * d;
      }
      red[grp][bin] = acc;
    }
    __syncthreads();
    if (threadIdx.x < NMEL) {
      double var = (red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x]) / (nfr - 1);
      s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var));
    }
    __syncthreads();
  } else {
    if (threadIdx.x < NMEL) { s_mean[threadIdx.x] = 0.f; s_rstd[threadIdx.x] = 1.f; }
    __syncthreads();
  }
  const long long total = (long long)frames_ld * NMEL;
  for (long long i = threadIdx.x; i < total; i += blockDim.x) {
    int f = (int)(i / NMEL), m = (int)(i - (long long)f * NMEL);
    float v = f < nfr ? (src[i] - s_mean[m]) * s_rstd[m] : 0.f;
    dst[i] = __float2half_rn(v);
  }
}

}  // namespace sb

extern "C" int sb_fbank(const float* wave, int64_t wave_ld, const int32_t* num_samples, int32_t batch, void* out,
                        int32_t out_frames_ld, float* work, int32_t* frames_out, int32_t standardize, sb_stream_t stream) {
  using namespace sb;
  SB_REQUIRE(wave && num_samples && out && work && batch > 0 && out_frames_ld > 0, SB_EINVAL, "sb_fbank: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = init_tables(st);
  if (rc) return rc;
  dim3 grid((out_frames_ld + 3) / 4, batch);
  fbank_frames_kernel<<<grid, 128, 0, st>>>(wave, wave_ld, num_samples, work, out_frames_ld, frames_out);
  SB_LAUNCH_OK();
  fbank_standardize_kernel<<<batch, 256, 0, st>>>(work, num_samples, (elem_t*)out, out_frames_ld, standardize);
  SB_LAUNCH_OK();
  return SB_OK;
}
