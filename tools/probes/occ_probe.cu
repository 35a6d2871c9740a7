// occupancy probe: which feature of a kernel makes cudaOccupancyMaxActiveBlocksPerMultiprocessor report 1 on sm_100a?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
#define ALLOC_IMM(N) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(N) : "memory"); \
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
#define DEALLOC_IMM(N) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "n"(N) : "memory");
__global__ void k_base(int* o) { if (o) o[threadIdx.x] = 1; }
__global__ void __launch_bounds__(256, 2) k_tmem256(int* o) { __shared__ uint32_t slot; if (threadIdx.x < 32) { ALLOC_IMM(256) } __syncthreads(); if (o) o[0] = slot; __syncthreads(); if (threadIdx.x < 32) { DEALLOC_IMM(256) } }
__global__ void __launch_bounds__(256, 2) k_tmem128(int* o) { __shared__ uint32_t slot; if (threadIdx.x < 32) { ALLOC_IMM(128) } __syncthreads(); if (o) o[0] = slot; __syncthreads(); if (threadIdx.x < 32) { DEALLOC_IMM(128) } }
__global__ void __launch_bounds__(256, 2) k_tmem_rt(int* o, uint32_t n) { __shared__ uint32_t slot; if (threadIdx.x < 32) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(n) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); } __syncthreads(); if (o) o[0] = slot; __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(n) : "memory"); }
__global__ void __launch_bounds__(256, 2) k_bar3(int* o) { asm volatile("bar.sync 1, 128;"); asm volatile("bar.sync 2, 256;"); __syncthreads(); if (o) o[0] = 1; }
__global__ void __launch_bounds__(256, 2) k_sleep(int* o) { __nanosleep(32); unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); if (o) o[0] = (int)t; }
__global__ void __launch_bounds__(256, 2) k_printf(int* o) { if (o && o[0] == 12345) { printf("x %d\n", o[1]); __trap(); } }
__global__ void __launch_bounds__(256, 2) k_atomic(unsigned* o) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(o) : "memory");
  asm volatile("fence.proxy.async;" ::: "memory"); __threadfence(); atomicAdd(o, v); }
__global__ void __launch_bounds__(256, 2) k_dynsmem(int* o) { extern __shared__ uint8_t sm[]; sm[threadIdx.x] = 1; __syncthreads(); if (o) o[0] = sm[5]; }
template <typename K> void report(const char* name, K k) {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k);
  printf("%-10s regs %3d:", name, fa.numRegs);
  for (int kb : {0, 48, 100, 109, 112}) { int o = -1; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k, 256, (size_t)kb * 1024); printf("  %dKB->%d", kb, o); }
  printf("\n");
}
int main() {
  int v; cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerMultiprocessor, 0); printf("smem/SM %d\n", v);
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, 0); printf("smem optin %d\n", v);
  cudaDeviceGetAttribute(&v, cudaDevAttrReservedSharedMemoryPerBlock, 0); printf("smem reserved/block %d\n", v);
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxRegistersPerMultiprocessor, 0); printf("regs/SM %d\n", v);
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxBlocksPerMultiprocessor, 0); printf("blocks/SM %d\n", v);
  report("base", k_base); report("tmem256", k_tmem256); report("tmem128", k_tmem128); report("tmem_rt", k_tmem_rt); report("bar3", k_bar3);
  report("sleep", k_sleep); report("printf", k_printf); report("atomic", k_atomic); report("dynsmem", k_dynsmem);
  return 0;
}
