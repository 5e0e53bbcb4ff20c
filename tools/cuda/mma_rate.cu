// Micro-benchmark (developer tool): cycles per tcgen05.mma kind::tf32 as a function of N, M and the source of A (shared
// memory descriptor or tensor memory), issued back to back by one thread into one accumulator.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I stf_b200/csrc tools/cuda/mma_rate.cu -o gpurun_out/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100.cuh"
using namespace stf::sm100;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  uint64_t d = 0;
  d |= (uint64_t)((a >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
               "r"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}

__global__ void __launch_bounds__(128, 1) rate_kernel(int M, int N, int ts, int iters, int distinct, long long *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  uint8_t *base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<float *>(base)[i] = 0.f;
  if (threadIdx.x == 0) mbar_init(&bar, 1), mbar_fence_init();
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x < 32) {
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_tf32(M, N);
    const uint64_t da = desc_sw128(smem_u32(base)), db = desc_sw128(smem_u32(base) + 32 * 1024);
    long long t0 = clock64();
    if (leader) {
      for (int i = 0; i < iters; ++i) {
        const uint32_t d = tmem + 256u + (distinct ? (uint32_t)(i & 1) * 0u : 0u);
        const uint64_t ao = (uint64_t)((i & 3) * 2) + (uint64_t)(distinct ? ((i >> 2) & 1) * 1024 : 0);
        if (ts) mma_ts(d, tmem + (uint32_t)((i & 7) * 8), db + (uint64_t)((i & 3) * 2), idesc, 1u);
        else umma_tf32(d, da + ao, db + (uint64_t)((i & 3) * 2), idesc, 1u);
      }
      umma_commit(&bar);
    }
    long long t1 = clock64();
    mbar_wait_spin(&bar, 0);
    long long t2 = clock64();
    if (leader) out[0] = t1 - t0, out[1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long *out;
  cudaMallocManaged(&out, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2000;
  for (int ts = 0; ts < 2; ++ts)
    for (int M : {128, 64})
      for (int N : {16, 32, 48, 64, 96, 128, 192, 256}) {
        if (M == 128 && N % 16) continue;
        for (int rep = 0; rep < 2; ++rep) {
          rate_kernel<<<1, 128, 100 * 1024>>>(M, N, ts, iters, 1, out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        printf("%s M=%3d N=%3d: issue %.1f cycles/MMA, complete %.1f cycles/MMA  (math floor N/2 = %d)\n", ts ? "TS" : "SS", M, N,
               (double)out[0] / iters, (double)out[1] / iters, N / 2);
      }
  return 0;
}
