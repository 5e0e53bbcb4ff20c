// Shared host/device helpers for libstf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/stf_b200.h"

namespace stf {

extern std::atomic<int64_t> g_launches;

inline int check_launch() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    (void)cudaGetLastError();  // clear the sticky-free launch error so later calls start clean
    return (int)e;
  }
  return STF_OK;
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

struct ScaleTable {
  float v[64];
  int levels;
  int monotone;  // 1: non-decreasing -> binary search is exact; 0: reference's linear count
  // Bucket LUT over the float bit pattern (monotone tables whose thresholds are at most two per bucket, e.g. the
  // reference's geometric 64-level table): key = (bits(sigma) >> 20) - key_min indexes `lut`, which holds the number
  // of thresholds below the bucket; at most two exact comparisons finish the count.  keys == 0: not usable.
  int key_min, keys;
  uint8_t lut[128];
};

// 16-byte streaming accesses: inputs are read once, outputs written once (no reuse in L1).
__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ int4 ldg_stream(const int4 *p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4 *p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void stg_stream(int4 *p, int4 v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

}  // namespace stf
