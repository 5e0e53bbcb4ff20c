// Epilogue arithmetic shared by the GEMM-engine kernels (conv_tcgen05.cu, mlp_tcgen05.cu): exact-erf GELU and the
// LayerNorm fold on PAIRS of accumulator columns with Blackwell's packed fp32 instructions (fma / mul / add .f32x2 ->
// FFMA2): the epilogues of the narrow-K layers (fc1 at C = 48 / 96, the fused MLP) are bound by instruction issue, not by
// memory, and a pair costs ~24 instructions instead of ~38.  Every operation is the scalar
// sequence's operation on each half with the same single rounding (IEEE fma / mul / add, the same MUFU approximations), so
// the two forms agree bit for bit.
#pragma once
#include <stdint.h>

namespace stf {
namespace epi {

__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t dup2(float x) { return pack2(x, x); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// Exact-erf GELU (nn.GELU default), Abramowitz-Stegun 7.1.26 (|abs error| <= 1.5e-7), branch-free:
//   ax = |x| / sqrt(2);  t = 1 / (1 + 0.3275911 ax);  erf|x| = 1 - (((((a5 t + a4) t) + a3) t + a2) t + a1) t exp(-ax^2)
//   gelu = 0.5 x (1 + sign(x) erf|x|)
__device__ __forceinline__ float gelu_erf(float x) {
  const float ax = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
  float p = fmaf(t, 1.061405429f, -1.453152027f);
  p = fmaf(t, p, 1.421413741f);
  p = fmaf(t, p, -0.284496736f);
  p = fmaf(t, p, 0.254829592f);
  const float e = __expf(-ax * ax);
  const float erf_abs = fmaf(-p * t, e, 1.0f);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}

// The same on two values (same bits per half: see the header comment).  -(p t) e + 1 is computed as p t e - 1 -- the same
// magnitude, and copysign only takes the magnitude.
__device__ __forceinline__ uint64_t gelu_erf2(uint64_t x) {
  const uint64_t ax = mul2(x & 0x7FFFFFFF7FFFFFFFull, dup2(0.70710678118654752440f));
  float u0, u1, t0, t1;
  unpack2(fma2(dup2(0.3275911f), ax, dup2(1.0f)), u0, u1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(u0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(u1));
  const uint64_t t = pack2(t0, t1);
  uint64_t p = fma2(t, dup2(1.061405429f), dup2(-1.453152027f));
  p = fma2(t, p, dup2(1.421413741f));
  p = fma2(t, p, dup2(-0.284496736f));
  p = fma2(t, p, dup2(0.254829592f));
  float z0, z1, e0, e1;
  unpack2(mul2(mul2(ax, ax), dup2(-1.4426950408889634f)), z0, z1);   // __expf(-ax * ax) = ex2.approx((ax * ax) * -log2(e))
  asm("ex2.approx.f32 %0, %1;" : "=f"(e0) : "f"(z0));
  asm("ex2.approx.f32 %0, %1;" : "=f"(e1) : "f"(z1));
  const uint64_t m = fma2(mul2(p, t), pack2(e0, e1), dup2(-1.0f));
  const uint64_t cs = (m & 0x7FFFFFFF7FFFFFFFull) | (x & 0x8000000080000000ull);
  return mul2(mul2(dup2(0.5f), x), add2(dup2(1.0f), cs));
}

}  // namespace epi
}  // namespace stf
