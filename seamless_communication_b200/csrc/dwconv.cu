// sb_dwconv_ln_silu: Conformer convolution module core for w2v-BERT 2.0 (conformer_shaw/builder.py:148-156):
//   causal depthwise Conv1d(kernel k, left pad k-1, no bias) -> LayerNorm(C) -> SiLU,   x,y: (B,T,C) fp16.
// (the pointwise convs and the GLU live in the sb_gemm epilogues around it.)
// One CTA = TT output frames x all C channels.  The (TT+k-1) x C input window is staged once in shared memory
// (each input element is reused k times from smem instead of HBM/L2), conv results stay in smem as fp32 for the
// LayerNorm, so HBM traffic is the algorithmic minimum: read x once (+ halo), write y once.
#include "common.cuh"

namespace sb {

constexpr int DW_TT = 8;
constexpr int DW_THREADS = 256;

__global__ void __launch_bounds__(DW_THREADS) dwconv_ln_silu_kernel(const elem_t* __restrict__ x, elem_t* __restrict__ y,
                                                                    const elem_t* __restrict__ w, const float* __restrict__ ln_w,
                                                                    const float* __restrict__ ln_b, int T, int C, int k) {
  extern __shared__ __align__(16) uint8_t smem_dw[];
  const int rows_in = DW_TT + k - 1;
  elem_t* sx = reinterpret_cast<elem_t*>(smem_dw);                 // [rows_in][C]
  elem_t* sw = sx + (size_t)rows_in * C;                           // [k][C]  (transposed weights)
  float* so = reinterpret_cast<float*>(sw + (size_t)k * C);        // [DW_TT][C]
  const int b = blockIdx.y, t0 = blockIdx.x * DW_TT;
  const elem_t* xb = x + (long long)b * T * C;
  // stage input rows t0-(k-1) .. t0+TT-1 (zeros before the sequence start / after its end)
  const int vec_per_row = C / 8;
  for (int i = threadIdx.x; i < rows_in * vec_per_row; i += DW_THREADS) {
    int r = i / vec_per_row, c = (i - r * vec_per_row) * 8;
    int t = t0 - (k - 1) + r;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (t >= 0 && t < T) u = *reinterpret_cast<const uint4*>(xb + (long long)t * C + c);
    *reinterpret_cast<uint4*>(sx + (size_t)r * C + c) = u;
  }
  for (int i = threadIdx.x; i < C * k; i += DW_THREADS) {  // w is [C][k] (Conv1d weight (C,1,k))
    int c = i / k, j = i - c * k;
    sw[(size_t)j * C + c] = w[i];
  }
  __syncthreads();
  // depthwise conv: each thread owns channel pairs
  for (int c = threadIdx.x * 2; c < C; c += DW_THREADS * 2) {
    float acc[DW_TT][2];
#pragma unroll
    for (int t = 0; t < DW_TT; ++t) acc[t][0] = acc[t][1] = 0.f;
    for (int j = 0; j < k; ++j) {
      float2 wv = __half22float2(*reinterpret_cast<const __half2*>(sw + (size_t)j * C + c));
#pragma unroll
      for (int t = 0; t < DW_TT; ++t) {
        float2 xv = __half22float2(*reinterpret_cast<const __half2*>(sx + (size_t)(t + j) * C + c));
        acc[t][0] += wv.x * xv.x;
        acc[t][1] += wv.y * xv.y;
      }
    }
#pragma unroll
    for (int t = 0; t < DW_TT; ++t) *reinterpret_cast<float2*>(so + (size_t)t * C + c) = make_float2(acc[t][0], acc[t][1]);
  }
  __syncthreads();
  // LayerNorm + SiLU: one warp per output frame
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t = warp; t < DW_TT; t += DW_THREADS / 32) {
    if (t0 + t >= T) continue;
    const float* row = so + (size_t)t * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += row[c];
    const float mean = warp_sum(s) / C;
    float sq = 0.f;
    for (int c = lane; c < C; c += 32) {
      float d = row[c] - mean;
      sq += d * d;
    }
    const float rstd = rsqrtf(warp_sum(sq) / C + 1e-5f);
    elem_t* yp = y + ((long long)b * T + t0 + t) * C;
    for (int c = lane * 2; c < C; c += 64) {
      float a0 = (row[c] - mean) * rstd * ln_w[c] + ln_b[c];
      float a1 = (row[c + 1] - mean) * rstd * ln_w[c + 1] + ln_b[c + 1];
      a0 = a0 / (1.f + __expf(-a0));
      a1 = a1 / (1.f + __expf(-a1));
      *reinterpret_cast<__half2*>(yp + c) = __floats2half2_rn(a0, a1);
    }
  }
}

// Fast path for a compile-time kernel size (w2v-BERT 2.0: k = 31).  Same tiling, but the per-channel taps live in
// registers (read straight from the (C, k) weight, no transposed smem copy whose column stores were 31-way bank
// conflicted) and the TT + K - 1 inputs of a channel pair are loaded from smem once and reused from registers by all
// TT outputs: 2 shared loads per 16 FMAs instead of 9.
template <int K>
__global__ void __launch_bounds__(DW_THREADS) dwconv_ln_silu_fixed_kernel(const elem_t* __restrict__ x, elem_t* __restrict__ y,
                                                                          const elem_t* __restrict__ w,
                                                                          const float* __restrict__ ln_w,
                                                                          const float* __restrict__ ln_b, int T, int C) {
  extern __shared__ __align__(16) uint8_t smem_dw[];
  constexpr int ROWS_IN = DW_TT + K - 1;
  elem_t* sx = reinterpret_cast<elem_t*>(smem_dw);                  // [ROWS_IN][C]
  float* so = reinterpret_cast<float*>(sx + (size_t)ROWS_IN * C);   // [DW_TT][C]
  const int b = blockIdx.y, t0 = blockIdx.x * DW_TT;
  const elem_t* xb = x + (long long)b * T * C;
  const int vec_per_row = C / 8;
  for (int i = threadIdx.x; i < ROWS_IN * vec_per_row; i += DW_THREADS) {
    int r = i / vec_per_row, c = (i - r * vec_per_row) * 8;
    int t = t0 - (K - 1) + r;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (t >= 0 && t < T) u = *reinterpret_cast<const uint4*>(xb + (long long)t * C + c);
    *reinterpret_cast<uint4*>(sx + (size_t)r * C + c) = u;
  }
  __syncthreads();
  for (int c = threadIdx.x * 2; c < C; c += DW_THREADS * 2) {
    float2 xr[ROWS_IN];
#pragma unroll
    for (int r = 0; r < ROWS_IN; ++r) xr[r] = __half22float2(*reinterpret_cast<const __half2*>(sx + (size_t)r * C + c));
    float2 acc[DW_TT];
#pragma unroll
    for (int t = 0; t < DW_TT; ++t) acc[t] = make_float2(0.f, 0.f);
    const elem_t* w0 = w + (size_t)c * K;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const float wa = __half2float(__ldg(w0 + j)), wb = __half2float(__ldg(w0 + K + j));
#pragma unroll
      for (int t = 0; t < DW_TT; ++t) {
        acc[t].x += wa * xr[t + j].x;
        acc[t].y += wb * xr[t + j].y;
      }
    }
#pragma unroll
    for (int t = 0; t < DW_TT; ++t) *reinterpret_cast<float2*>(so + (size_t)t * C + c) = acc[t];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t = warp; t < DW_TT; t += DW_THREADS / 32) {
    if (t0 + t >= T) continue;
    const float* row = so + (size_t)t * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += row[c];
    const float mean = warp_sum(s) / C;
    float sq = 0.f;
    for (int c = lane; c < C; c += 32) {
      float d = row[c] - mean;
      sq += d * d;
    }
    const float rstd = rsqrtf(warp_sum(sq) / C + 1e-5f);
    elem_t* yp = y + ((long long)b * T + t0 + t) * C;
    for (int c = lane * 2; c < C; c += 64) {
      float a0 = (row[c] - mean) * rstd * ln_w[c] + ln_b[c];
      float a1 = (row[c + 1] - mean) * rstd * ln_w[c + 1] + ln_b[c + 1];
      a0 = a0 / (1.f + __expf(-a0));
      a1 = a1 / (1.f + __expf(-a1));
      *reinterpret_cast<__half2*>(yp + c) = __floats2half2_rn(a0, a1);
    }
  }
}

}  // namespace sb

extern "C" int sb_dwconv_ln_silu(const void* x, void* y, const void* w, const float* ln_w, const float* ln_b,
                                 int32_t batch, int32_t T, int32_t C, int32_t k, sb_stream_t stream) {
  using namespace sb;
  SB_REQUIRE(x && y && w && ln_w && ln_b && batch > 0 && T > 0 && k > 0, SB_EINVAL, "sb_dwconv_ln_silu: bad args");
  SB_REQUIRE(C % 8 == 0, SB_ENOSUP, "sb_dwconv_ln_silu: C must be a multiple of 8");
  dim3 grid((T + DW_TT - 1) / DW_TT, batch);
  if (k == 31) {
    const size_t smem31 = (size_t)(DW_TT + 30) * C * 2 + (size_t)DW_TT * C * 4;
    if (smem31 <= 200 * 1024) {
      static size_t configured31 = 0;
      if (smem31 > configured31) {
        SB_CUDA_OK(cudaFuncSetAttribute(dwconv_ln_silu_fixed_kernel<31>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem31));
        configured31 = smem31;
      }
      dwconv_ln_silu_fixed_kernel<31><<<grid, DW_THREADS, smem31, (cudaStream_t)stream>>>(
          (const elem_t*)x, (elem_t*)y, (const elem_t*)w, ln_w, ln_b, T, C);
      SB_LAUNCH_OK();
      return SB_OK;
    }
  }
  size_t smem = (size_t)(DW_TT + k - 1) * C * 2 + (size_t)k * C * 2 + (size_t)DW_TT * C * 4;
  SB_REQUIRE(smem <= 200 * 1024, SB_ENOSUP, "sb_dwconv_ln_silu: C*k too large for shared memory (%zu B)", smem);
  static size_t configured = 0;
  if (smem > configured) {
    SB_CUDA_OK(cudaFuncSetAttribute(dwconv_ln_silu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dwconv_ln_silu_kernel<<<grid, DW_THREADS, smem, (cudaStream_t)stream>>>((const elem_t*)x, (elem_t*)y, (const elem_t*)w, ln_w,
                                                                         ln_b, T, C, k);
  SB_LAUNCH_OK();
  return SB_OK;
}
