// EMMA monotonic-attention step probability (SeamlessStreaming):
//   PChooseLayer.forward (models/monotonic_decoder/p_choose.py:120-148):
//     q = EnergyProjection(seqs)  (4 x Linear+ReLU: sb_gemm),  k = EnergyProjection(AvgPool1d(keys, ratio, ceil))
//     p_choose[h,s,j] = sigmoid((q_h[s] . k_h[j] / sqrt(64) + energy_bias) / temperature)
// The two kernels here are the pooling and the energy/sigmoid; both are tiny and latency bound.
#include "common.cuh"

namespace sb {

// y[b][j][c] = mean_{t in [j*r, min((j+1)*r, T))} x[b][t][c]      (AvgPool1d(kernel=stride=r, ceil_mode=True))
__global__ void avgpool_time_kernel(const elem_t* __restrict__ x, elem_t* __restrict__ y, int T, int C, int r, int Tp) {
  const int j = blockIdx.x, b = blockIdx.y;
  const int t0 = j * r, t1 = min(t0 + r, T);
  const float inv = 1.f / (float)(t1 - t0);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int t = t0; t < t1; ++t) acc += __half2float(x[((long long)b * T + t) * C + c]);
    y[((long long)b * Tp + j) * C + c] = __float2half_rn(acc * inv);
  }
}

// one warp per (head, s, j): 64-dim dot product
__global__ void pchoose_kernel(const elem_t* __restrict__ q, const elem_t* __restrict__ k, float* __restrict__ p, int S, int Sp,
                               int heads, float bias, float inv_temp) {
  const int lane = threadIdx.x & 31;
  const long long idx = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long total = (long long)heads * S * Sp;
  if (idx >= total) return;
  const int j = (int)(idx % Sp), s = (int)((idx / Sp) % S), h = (int)(idx / ((long long)Sp * S));
  const int dim = heads * 64;
  const float2 a = __half22float2(reinterpret_cast<const __half2*>(q + (long long)s * dim + h * 64)[lane]);
  const float2 b = __half22float2(reinterpret_cast<const __half2*>(k + (long long)j * dim + h * 64)[lane]);
  const float e = warp_sum(a.x * b.x + a.y * b.y) * 0.125f + bias;
  if (lane == 0) p[idx] = 1.f / (1.f + __expf(-e * inv_temp));
}

}  // namespace sb

extern "C" int sb_avgpool_time(const void* x, void* y, int32_t batch, int32_t T, int32_t C, int32_t ratio, sb_stream_t stream) {
  using namespace sb;
  SB_REQUIRE(x && y && batch > 0 && T > 0 && C > 0 && ratio > 0, SB_EINVAL, "sb_avgpool_time: bad args");
  const int Tp = (T + ratio - 1) / ratio;
  avgpool_time_kernel<<<dim3(Tp, batch), 256, 0, (cudaStream_t)stream>>>((const elem_t*)x, (elem_t*)y, T, C, ratio, Tp);
  SB_LAUNCH_OK();
  return SB_OK;
}

extern "C" int sb_pchoose(const void* q_energy, const void* k_energy, float* p, int32_t S, int32_t Sp, int32_t heads,
                          float energy_bias, float temperature, sb_stream_t stream) {
  using namespace sb;
  SB_REQUIRE(q_energy && k_energy && p && S > 0 && Sp > 0 && heads > 0 && temperature > 0.f, SB_EINVAL, "sb_pchoose: bad args");
  const long long total = (long long)heads * S * Sp;
  pchoose_kernel<<<(unsigned)((total + 7) / 8), 256, 0, (cudaStream_t)stream>>>((const elem_t*)q_energy, (const elem_t*)k_energy, p,
                                                                              S, Sp, heads, energy_bias, 1.f / temperature);
  SB_LAUNCH_OK();
  return SB_OK;
}
