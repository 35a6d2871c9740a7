// Development probe (not on the product path): measures TMA global->shared throughput of one CTA ring
// as a function of stage count, box rows and number of concurrent issuing threads.
#include "common.cuh"

namespace sb {
__device__ __forceinline__ uint32_t p_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool p_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(p_smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool p_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(p_smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ int g_wait_mode = 0;
__device__ __forceinline__ void p_wait(uint64_t* bar, uint32_t parity, int mode = 0) {
  uint32_t spins = 0;
  if (mode == 0) { while (!p_try(bar, parity)) { if (++spins > (1u << 26)) __trap(); } }
  else { while (!p_test(bar, parity)) { if (++spins > (1u << 28)) __trap(); } }
}

// grid: ctas; block: 64 threads (warp 0 = issuers, warp 1 lane 0 = consumer)
__global__ void __launch_bounds__(64) tma_probe_kernel(const __grid_constant__ CUtensorMap tm, int stages, int box_rows,
                                                       int issuers, int iters, int rows_total, long long* cycles_out, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = box_rows * 128;
  uint64_t* full = (uint64_t*)(smem + stages * stage_bytes);
  uint64_t* empty = full + stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(p_smem_u32(&full[s])), "r"(issuers) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(p_smem_u32(&empty[s])), "r"(1) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = clock64();
  const int sub_rows = box_rows / issuers;  // each issuer loads a sub-box of sub_rows rows (tensor map box = sub_rows)
  if (warp == 0 && lane < issuers) {
    int s = 0; uint32_t ph = 0;
    int row = (int)((long long)blockIdx.x * 7919 * box_rows % (rows_total - 2 * box_rows)) + lane * sub_rows;
    for (int it = 0; it < iters; ++it) {
      p_wait(&empty[s], ph ^ 1, mode);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(p_smem_u32(&full[s])), "r"(sub_rows * 128) : "memory");
      int col = (it & 7) * 64;
      row += box_rows; if (row >= rows_total - 2 * box_rows) row -= rows_total - 2 * box_rows;
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(p_smem_u32(smem + s * stage_bytes + lane * sub_rows * 128)), "l"(reinterpret_cast<uint64_t>(&tm)),
                     "r"(p_smem_u32(&full[s])), "r"(col), "r"(row) : "memory");
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      p_wait(&full[s], ph, mode);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(p_smem_u32(&empty[s])) : "memory");
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles_out[blockIdx.x] = clock64() - t0;
}
}  // namespace sb

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

extern "C" int sb_tma_probe(const void* base, int64_t rows_total, int64_t ld_elems, int ctas, int stages, int box_rows,
                            int issuers, int iters, long long* cycles_out, sb_stream_t stream, int mode) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  SB_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)ld_elems, (cuuint64_t)rows_total};
  cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)(box_rows / issuers)};
  cuuint32_t es[2] = {1, 1};
  CUresult r = ((EncFn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SB_REQUIRE(r == CUDA_SUCCESS, SB_ECUDA, "encode failed %d", (int)r);
  size_t smem = (size_t)stages * box_rows * 128 + 1024 + 256;
  SB_CUDA_OK(cudaFuncSetAttribute(sb::tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  sb::tma_probe_kernel<<<ctas, 64, smem, (cudaStream_t)stream>>>(tm, stages, box_rows, issuers, iters, (int)rows_total, cycles_out, mode);
  SB_LAUNCH_OK();
  return SB_OK;
}
