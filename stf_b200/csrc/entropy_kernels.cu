// Entropy-model element-wise kernels: single pass, HBM-bound,
// 128-bit coalesced streaming accesses, no atomics (deterministic, batch-invariant).
//
//   stf_build_indexes            GaussianConditional.build_indexes  entropy_models.py:661-666
//   stf_gaussian_compress_step   build_indexes + quantize("symbols") + dequantize, stf.py:717-719
//   stf_quantize_symbols         EntropyModel.quantize               entropy_models.py:126-150
//   stf_dequantize               EntropyModel.dequantize             entropy_models.py:158-165
//   stf_gaussian_likelihood      GaussianConditional.forward (eval)  entropy_models.py:645-659
//   stf_entropy_bottleneck       EntropyBottleneck.forward (eval)    entropy_models.py:446-489
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace stf {

std::atomic<int64_t> g_launches{0};

namespace {

constexpr int kThreads = 256;
constexpr int kVecPerThread = 4;  // 4 x float4 in flight per thread per stream

// sigma = torch.max(x, bound): NaN propagates (fmaxf would drop it).
__device__ __forceinline__ float lower_bound_f(float x, float bound) { return x < bound ? bound : x; }

// idx = #{ i < levels-1 : table[i] < sigma }  ==  (levels-1) - #{ i : sigma <= table[i] }.
// Monotone tables: 6-step branch-free binary search in shared memory; otherwise linear count.
template <bool kMonotone>
__device__ __forceinline__ int scale_index(float sigma, const float *__restrict__ t, int levels) {
  const int last = levels - 1;
  if (sigma != sigma) return last;
  if (kMonotone) {
    int lo = 0;
#pragma unroll
    for (int step = 32; step >= 1; step >>= 1) {
      int probe = lo + step;
      if (probe <= last && t[probe - 1] < sigma) lo = probe;
    }
    return lo;
  } else {
    int c = 0;
    for (int i = 0; i < last; ++i) c += (t[i] < sigma) ? 1 : 0;
    return c;
  }
}

// Same count through the bucket LUT (see ScaleTable): one conflict-free shared-memory byte lookup (the 128-entry LUT is
// 32 words = one per bank) + two exact comparisons against the thresholds of that bucket.  `tp` = thresholds
// t[0..levels-2] padded with +inf.  Bit-exact with the search above for every float input (tests/test_gpu_entropy.py).
__device__ __forceinline__ int scale_index_lut(float sigma, const float *__restrict__ tp, const uint8_t *__restrict__ lut,
                                               int key_min, int keys, int levels) {
  // The bit pattern as a SIGNED integer orders like the float for sigma > 0; negative floats (sign bit) land below
  // key_min and clamp to bucket 0.  NaNs of either sign are caught by the final select.
  int key = (__float_as_int(sigma) >> 20) - key_min;
  key = min(max(key, 0), keys - 1);
  const int base = lut[key];   // byte load: the 128-entry LUT spans 32 words = one per bank, conflict-free
  const int idx = base + (tp[base] < sigma ? 1 : 0) + (tp[base + 1] < sigma ? 1 : 0);
  return sigma != sigma ? levels - 1 : idx;
}

__device__ __forceinline__ int round_to_symbol(float v) { return __float2int_rn(v); }  // half-to-even

struct SliceGeom {
  int64_t inner;           // channels * plane, contiguous per batch element
  int64_t y_batch_stride;  // of the (possibly larger) source tensor
  int64_t out_batch_stride;
};

// ---------------------------------------------------------------------------------------------
// compress step: (y, mu, scale) -> (symbols, indexes, y_hat)
// ---------------------------------------------------------------------------------------------
// kIdxOnly: stf_build_indexes (scales -> indexes only): the y / means / symbols / y_hat paths compile out, which takes the
// kernel from 125 to ~40 registers, i.e. from 2 to 8 resident CTAs per SM for its single input stream (ncu: 0.66 of the HBM peak
// at 125 registers, latency-bound).
template <bool kVec, int kMode, bool kIdxOnly>  // kMode: 0 = linear count, 1 = binary search, 2 = bucket LUT
__global__ void __launch_bounds__(kThreads)
compress_step_kernel(const float *__restrict__ y, const float *__restrict__ scales,
                     const float *__restrict__ means, int32_t *__restrict__ symbols,
                     int32_t *__restrict__ indexes, float *__restrict__ y_hat, SliceGeom g,
                     float scale_bound, const __grid_constant__ ScaleTable table) {
  __shared__ float t[68];
  __shared__ __align__(16) uint8_t lut[128];
  if (kMode == 2) {  // thresholds t[0..levels-2] padded with +inf, LUT bytes packed four per word
    if (threadIdx.x < 68) t[threadIdx.x] = (int)threadIdx.x < table.levels - 1 ? table.v[threadIdx.x] : __int_as_float(0x7f800000);
    if (threadIdx.x >= 128) lut[threadIdx.x - 128] = table.lut[threadIdx.x - 128];
  } else if (threadIdx.x < 64) {
    t[threadIdx.x] = table.v[threadIdx.x < table.levels ? threadIdx.x : table.levels - 1];
  }
  __syncthreads();
  auto index_of = [&](float sigma) -> int {
    if (kMode == 2) return scale_index_lut(sigma, t, lut, table.key_min, table.keys, table.levels);
    return scale_index<kMode == 1>(sigma, t, table.levels);
  };
  const int b = blockIdx.y;
  const float *yb = (!kIdxOnly && y) ? y + (int64_t)b * g.y_batch_stride : nullptr;
  const float *sb = scales ? scales + (int64_t)b * g.inner : nullptr;
  const float *mb = (!kIdxOnly && means) ? means + (int64_t)b * g.inner : nullptr;
  int32_t *symb = (!kIdxOnly && symbols) ? symbols + (int64_t)b * g.out_batch_stride : nullptr;
  int32_t *idxb = indexes ? indexes + (int64_t)b * g.out_batch_stride : nullptr;
  float *yhb = (!kIdxOnly && y_hat) ? y_hat + (int64_t)b * g.inner : nullptr;

  if (kVec) {
    const int64_t nvec = g.inner >> 2;
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t v0 = (int64_t)blockIdx.x * kThreads + threadIdx.x; v0 < nvec; v0 += stride * kVecPerThread) {
      float4 yy[kVecPerThread], mm[kVecPerThread], ss[kVecPerThread];
#pragma unroll
      for (int u = 0; u < kVecPerThread; ++u) {
        int64_t v = v0 + u * stride;
        if (v < nvec) {
          if (!kIdxOnly) {
            if (yb) yy[u] = ldg_stream(reinterpret_cast<const float4 *>(yb) + v);
            mm[u] = mb ? ldg_stream(reinterpret_cast<const float4 *>(mb) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          if (sb) ss[u] = ldg_stream(reinterpret_cast<const float4 *>(sb) + v);
        }
      }
#pragma unroll
      for (int u = 0; u < kVecPerThread; ++u) {
        int64_t v = v0 + u * stride;
        if (v >= nvec) continue;
        if (!kIdxOnly && yb) {
          int4 q;
          q.x = round_to_symbol(yy[u].x - mm[u].x);
          q.y = round_to_symbol(yy[u].y - mm[u].y);
          q.z = round_to_symbol(yy[u].z - mm[u].z);
          q.w = round_to_symbol(yy[u].w - mm[u].w);
          if (symb) stg_stream(reinterpret_cast<int4 *>(symb) + v, q);
          if (yhb)
            stg_stream(reinterpret_cast<float4 *>(yhb) + v,
                       make_float4((float)q.x + mm[u].x, (float)q.y + mm[u].y, (float)q.z + mm[u].z,
                                   (float)q.w + mm[u].w));
        }
        if (sb && idxb) {
          int4 ix;
          ix.x = index_of(lower_bound_f(ss[u].x, scale_bound));
          ix.y = index_of(lower_bound_f(ss[u].y, scale_bound));
          ix.z = index_of(lower_bound_f(ss[u].z, scale_bound));
          ix.w = index_of(lower_bound_f(ss[u].w, scale_bound));
          stg_stream(reinterpret_cast<int4 *>(idxb) + v, ix);
        }
      }
    }
  } else {
    const int64_t stride = (int64_t)gridDim.x * kThreads;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < g.inner; i += stride) {
      float m = mb ? mb[i] : 0.f;
      if (yb) {
        int q = round_to_symbol(yb[i] - m);
        if (symb) symb[i] = q;
        if (yhb) yhb[i] = (float)q + m;
      }
      if (sb && idxb) idxb[i] = index_of(lower_bound_f(sb[i], scale_bound));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// dequantize: int32 symbols (+ strided batch) + means -> y_hat
// ---------------------------------------------------------------------------------------------
template <bool kVec>
__global__ void __launch_bounds__(kThreads)
dequantize_kernel(const int32_t *__restrict__ symbols, const float *__restrict__ means,
                  float *__restrict__ y_hat, SliceGeom g) {
  const int b = blockIdx.y;
  const int32_t *sb = symbols + (int64_t)b * g.y_batch_stride;
  const float *mb = means ? means + (int64_t)b * g.inner : nullptr;
  float *ob = y_hat + (int64_t)b * g.inner;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  if (kVec) {
    const int64_t nvec = g.inner >> 2;
    for (int64_t v = (int64_t)blockIdx.x * kThreads + threadIdx.x; v < nvec; v += stride) {
      int4 q = ldg_stream(reinterpret_cast<const int4 *>(sb) + v);
      float4 m = mb ? ldg_stream(reinterpret_cast<const float4 *>(mb) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
      stg_stream(reinterpret_cast<float4 *>(ob) + v,
                 make_float4((float)q.x + m.x, (float)q.y + m.y, (float)q.z + m.z, (float)q.w + m.w));
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < g.inner; i += stride)
      ob[i] = (float)sb[i] + (mb ? mb[i] : 0.f);
  }
}

// ---------------------------------------------------------------------------------------------
// Gaussian likelihood (eval) fused with ste_round
// ---------------------------------------------------------------------------------------------
// erfc with relative accuracy (needed in the tails: the likelihood is a difference of two of them, floored at
// 1e-9): erfc(z) = t exp(-z^2 + P(t)), t = 1/(1 + z/2), z >= 0 -- the Chebyshev fit of Numerical Recipes' erfcc
// (fractional error < 1.2e-7 everywhere; measured here in fp32: < 2e-6 for z <= 4.5, 4e-6 to z = 6.5).
// 1 MUFU.RCP + 1 MUFU.EX2 + 13 FMA-pipe instructions instead of erfcf's ~45 with a division: the kernel was
// instruction-issue bound at 0.54-0.64 of the HBM roofline with erfcf.  Tolerance bar: 1e-3 |ref| + 1e-9.
__device__ __forceinline__ float erfc_fast(float x) {
  const float z = fabsf(x);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.5f, z, 1.0f)));
  float p = fmaf(t, 0.17087277f, -0.82215223f);
  p = fmaf(t, p, 1.48851587f);
  p = fmaf(t, p, -1.13520398f);
  p = fmaf(t, p, 0.27886807f);
  p = fmaf(t, p, -0.18628806f);
  p = fmaf(t, p, 0.09678418f);
  p = fmaf(t, p, 0.37409196f);
  p = fmaf(t, p, 1.00002368f);
  p = fmaf(t, p, -1.26551223f);
  const float r = t * exp2f(fmaf(-z, z, p) * 1.4426950408889634f);
  return x >= 0.f ? r : 2.0f - r;   // (NaN compares false and is carried by r)
}

template <bool kSte>
__device__ __forceinline__ float gauss_lik(float y, float mu, float scale, float scale_bound,
                                           float lik_bound, float *y_hat_out) {
  const float kNegInvSqrt2 = -0.70710678118654752440f;  // float(-(2 ** -0.5))
  const float t = y - mu;
  const float r = rintf(t);
  float yh = r + mu;                                    // quantize(., "dequantize", means)
  // ste_round(t) + mu evaluates round(t) - t + t left to right (ops/ops.py:34, stf.py:626)
  *y_hat_out = kSte ? ((r - t) + t) + mu : yh;
  float v = fabsf(yh - mu);
  float s = lower_bound_f(scale, scale_bound);
  const float inv_s = 1.0f / s;   // one IEEE reciprocal for both arguments (the reference divides twice: <= 1.5 ulp apart)
  float upper = 0.5f * erfc_fast(kNegInvSqrt2 * ((0.5f - v) * inv_s));
  float lower = 0.5f * erfc_fast(kNegInvSqrt2 * ((-0.5f - v) * inv_s));
  return lower_bound_f(upper - lower, lik_bound);
}

template <bool kVec, bool kSte>
__global__ void __launch_bounds__(kThreads)
gaussian_likelihood_kernel(const float *__restrict__ y, const float *__restrict__ scales,
                           const float *__restrict__ means, float *__restrict__ y_hat,
                           float *__restrict__ lik, SliceGeom g, float scale_bound, float lik_bound) {
  const int b = blockIdx.y;
  const float *yb = y + (int64_t)b * g.y_batch_stride;
  const float *sb = scales + (int64_t)b * g.inner;
  const float *mb = means ? means + (int64_t)b * g.inner : nullptr;
  float *yhb = y_hat ? y_hat + (int64_t)b * g.inner : nullptr;
  float *lb = lik + (int64_t)b * g.inner;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  if (kVec) {
    const int64_t nvec = g.inner >> 2;
    constexpr int U = 2;
    for (int64_t v0 = (int64_t)blockIdx.x * kThreads + threadIdx.x; v0 < nvec; v0 += stride * U) {
      float4 yy[U], mm[U], ss[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        int64_t v = v0 + u * stride;
        if (v < nvec) {
          yy[u] = ldg_stream(reinterpret_cast<const float4 *>(yb) + v);
          ss[u] = ldg_stream(reinterpret_cast<const float4 *>(sb) + v);
          mm[u] = mb ? ldg_stream(reinterpret_cast<const float4 *>(mb) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        int64_t v = v0 + u * stride;
        if (v >= nvec) continue;
        float4 h, l;
        l.x = gauss_lik<kSte>(yy[u].x, mm[u].x, ss[u].x, scale_bound, lik_bound, &h.x);
        l.y = gauss_lik<kSte>(yy[u].y, mm[u].y, ss[u].y, scale_bound, lik_bound, &h.y);
        l.z = gauss_lik<kSte>(yy[u].z, mm[u].z, ss[u].z, scale_bound, lik_bound, &h.z);
        l.w = gauss_lik<kSte>(yy[u].w, mm[u].w, ss[u].w, scale_bound, lik_bound, &h.w);
        stg_stream(reinterpret_cast<float4 *>(lb) + v, l);
        if (yhb) stg_stream(reinterpret_cast<float4 *>(yhb) + v, h);
      }
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < g.inner; i += stride) {
      float h;
      lb[i] = gauss_lik<kSte>(yb[i], mb ? mb[i] : 0.f, sb[i], scale_bound, lik_bound, &h);
      if (yhb) yhb[i] = h;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Slice step on NHWC operands (the conv kernel's layout): mu / scale / y / y_hat are read and written pixel-major
// (element (b, p, c) at [(b * plane + p) * ld + c], a channel slice of a wider NHWC tensor is fine), while symbols, indexes
// and likelihoods leave in the reference's coding order (b, c, p) (stf.py:721-722) -- a 32 x 32 transpose through shared
// memory, so both sides are 128-byte coalesced.  Same arithmetic (device functions) as the NCHW kernels above: bit-identical
// symbols / indexes / y_hat for identical inputs.  One block = 32 pixels x (<= 32) channels of one image.
// ---------------------------------------------------------------------------------------------
struct SliceNhwc {
  const float *y;        int y_ld;     // slice of the latent (already offset to its first channel); nullptr in decode
  const float *scales;   int s_ld;
  const float *means;    int m_ld;
  const int32_t *sym_in; int64_t sym_in_bstride;   // decode: symbols of this slice in coding order
  int32_t *sym_out, *idx_out; int64_t out_bstride;
  float *y_hat;          int yh_ld;
  float *lik;            int64_t lik_bstride;      // forward: likelihoods in (b, c, p) order
  int C, plane;
  float scale_bound, lik_bound;
  int ste_round;
  int narrow;            // symbols_out is int16_t*, indexes_out uint8_t* (3 instead of 8 bytes per symbol over PCIe)
  int32_t *overflow;     // narrow: set to 1 when a symbol does not fit int16 or an index does not fit uint8
};

template <int kMode>  // index search as in compress_step_kernel
__global__ void __launch_bounds__(256)
slice_step_nhwc_kernel(const SliceNhwc a, const __grid_constant__ ScaleTable table) {
  __shared__ float t[68];
  __shared__ __align__(16) uint8_t lut[128];
  __shared__ int32_t tile_a[32][33];   // symbols (in or out)
  __shared__ int32_t tile_b[32][33];   // indexes, or likelihood bits
  if (kMode == 2) {
    if (threadIdx.x < 68) t[threadIdx.x] = (int)threadIdx.x < table.levels - 1 ? table.v[threadIdx.x] : __int_as_float(0x7f800000);
    if (threadIdx.x >= 128) lut[threadIdx.x - 128] = table.lut[threadIdx.x - 128];
  } else if (threadIdx.x < 64) {
    t[threadIdx.x] = table.v[threadIdx.x < table.levels ? threadIdx.x : table.levels - 1];
  }
  auto index_of = [&](float sigma) -> int {
    if (kMode == 2) return scale_index_lut(sigma, t, lut, table.key_min, table.keys, table.levels);
    return scale_index<kMode == 1>(sigma, t, table.levels);
  };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, p0 = blockIdx.x * 32;
  if (a.sym_in) {  // coding order -> tile: warp = channel, lane = pixel
    for (int c = warp; c < a.C; c += 8) {
      const int p = p0 + lane;
      if (p < a.plane) tile_a[c][lane] = a.sym_in[(int64_t)b * a.sym_in_bstride + (int64_t)c * a.plane + p];
    }
  }
  __syncthreads();
  for (int pp = warp; pp < 32; pp += 8) {  // warp = pixel, lane = channel: 128-byte rows of the NHWC tensors
    const int p = p0 + pp, c = lane;
    if (p >= a.plane || c >= a.C) continue;
    const int64_t pix = (int64_t)b * a.plane + p;
    const float m = a.means ? a.means[pix * a.m_ld + c] : 0.f;
    const float sc = a.scales ? a.scales[pix * a.s_ld + c] : 0.f;
    if (a.lik) {          // forward: likelihood + (ste-)rounded y_hat
      float yh;
      const float l = a.ste_round ? gauss_lik<true>(a.y[pix * a.y_ld + c], m, sc, a.scale_bound, a.lik_bound, &yh)
                                  : gauss_lik<false>(a.y[pix * a.y_ld + c], m, sc, a.scale_bound, a.lik_bound, &yh);
      tile_b[c][pp] = __float_as_int(l);
      if (a.y_hat) a.y_hat[pix * a.yh_ld + c] = yh;
      continue;
    }
    if (a.y) {            // encode: symbols + y_hat
      const int q = round_to_symbol(a.y[pix * a.y_ld + c] - m);
      tile_a[c][pp] = q;
      if (a.y_hat) a.y_hat[pix * a.yh_ld + c] = (float)q + m;
    } else if (a.sym_in) {  // decode: y_hat from the decoded symbols
      a.y_hat[pix * a.yh_ld + c] = (float)tile_a[c][pp] + m;
    }
    if (a.scales && a.idx_out) tile_b[c][pp] = index_of(lower_bound_f(sc, a.scale_bound));
  }
  __syncthreads();
  for (int c = warp; c < a.C; c += 8) {    // tile -> coding order
    const int p = p0 + lane;
    if (p >= a.plane) continue;
    const int64_t o = (int64_t)c * a.plane + p;
    if (a.lik) a.lik[(int64_t)b * a.lik_bstride + o] = __int_as_float(tile_b[c][lane]);
    if (!a.narrow) {
      if (a.sym_out) a.sym_out[(int64_t)b * a.out_bstride + o] = tile_a[c][lane];
      if (a.idx_out) a.idx_out[(int64_t)b * a.out_bstride + o] = tile_b[c][lane];
    } else {
      bool bad = false;
      if (a.sym_out) {
        const int32_t q = tile_a[c][lane];
        bad = q != (int32_t)(int16_t)q;
        reinterpret_cast<int16_t *>(a.sym_out)[(int64_t)b * a.out_bstride + o] = (int16_t)q;
      }
      if (a.idx_out) {
        const int32_t ix = tile_b[c][lane];
        bad = bad || (uint32_t)ix > 255u;
        reinterpret_cast<uint8_t *>(a.idx_out)[(int64_t)b * a.out_bstride + o] = (uint8_t)ix;
      }
      if (bad) *a.overflow = 1;   // (benign race: every writer stores the same value)
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Entropy bottleneck (factorised prior), eval mode.  One block = one (batch, channel) plane tile;
// the 58 pre-activated parameters (+ median) of the channel sit in shared memory.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float eb_logits(float x, const float *__restrict__ p) {
  // layer 0: 1 -> 3
  float h0 = fmaf(p[0], x, p[3]), h1 = fmaf(p[1], x, p[4]), h2 = fmaf(p[2], x, p[5]);
  h0 = fmaf(p[6], tanhf(h0), h0);
  h1 = fmaf(p[7], tanhf(h1), h1);
  h2 = fmaf(p[8], tanhf(h2), h2);
  p += 9;
#pragma unroll
  for (int l = 0; l < 3; ++l) {  // layers 1..3: 3 -> 3
    float g0 = p[0] * h0 + p[1] * h1 + p[2] * h2 + p[9];
    float g1 = p[3] * h0 + p[4] * h1 + p[5] * h2 + p[10];
    float g2 = p[6] * h0 + p[7] * h1 + p[8] * h2 + p[11];
    h0 = fmaf(p[12], tanhf(g0), g0);
    h1 = fmaf(p[13], tanhf(g1), g1);
    h2 = fmaf(p[14], tanhf(g2), g2);
    p += 15;
  }
  return p[0] * h0 + p[1] * h1 + p[2] * h2 + p[3];  // layer 4: 3 -> 1
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(128)
entropy_bottleneck_kernel(const float *__restrict__ z, const float *__restrict__ params,
                          float *__restrict__ z_hat, float *__restrict__ lik,
                          int32_t *__restrict__ symbols, int channels, int64_t plane, float lik_bound,
                          int ste_round) {
  __shared__ float p[STF_EB_PARAMS];
  const int c = blockIdx.y, b = blockIdx.z;
  if (threadIdx.x < STF_EB_PARAMS) p[threadIdx.x] = params[(int64_t)c * STF_EB_PARAMS + threadIdx.x];
  __syncthreads();
  const float median = p[STF_EB_MEDIAN_SLOT];
  const int64_t base = ((int64_t)b * channels + c) * plane;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (int64_t)gridDim.x * blockDim.x) {
    const float t = z[base + i] - median;
    float r = rintf(t);
    float v = r + median;
    if (symbols) symbols[base + i] = __float2int_rn(r);
    if (z_hat) z_hat[base + i] = ste_round ? ((r - t) + t) + median : v;  // ops/ops.py:34, stf.py:602-604
    if (lik) {
      float lo = eb_logits(v - 0.5f, p), hi = eb_logits(v + 0.5f, p);
      float sum = lo + hi;
      float sgn = sum > 0.f ? -1.f : (sum < 0.f ? 1.f : 0.f);  // -torch.sign(lo + hi)
      float l = fabsf(sigmoidf_(sgn * hi) - sigmoidf_(sgn * lo));
      lik[base + i] = lower_bound_f(l, lik_bound);
    }
  }
}

__global__ void __launch_bounds__(kThreads)
quantize_symbols_kernel(const float *__restrict__ x, const float *__restrict__ means,
                        int32_t *__restrict__ symbols, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride)
    symbols[i] = round_to_symbol(x[i] - (means ? means[i] : 0.f));
}

__global__ void __launch_bounds__(kThreads)
quantize_dequantize_kernel(const float *__restrict__ x, const float *__restrict__ means,
                           float *__restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
    const float m = means ? means[i] : 0.f;
    out[i] = rintf(x[i] - m) + m;
  }
}

int fill_table(ScaleTable *t, const float *table_host, int levels) {
  if (!table_host || levels < 1 || levels > 64) return STF_E_TABLE;
  t->levels = levels;
  t->monotone = 1;
  for (int i = 0; i < 64; ++i) t->v[i] = table_host[i < levels ? i : levels - 1];
  for (int i = 1; i < levels; ++i)
    if (!(table_host[i] >= table_host[i - 1])) t->monotone = 0;
  // bucket LUT (see ScaleTable): thresholds are t[0 .. levels-2]
  t->key_min = 0, t->keys = 0;
  for (uint8_t &b : t->lut) b = 0;
  const int nthr = levels - 1;
  auto bits_of = [](float f) { uint32_t u; memcpy(&u, &f, 4); return u; };
  auto float_of = [](uint32_t u) { float f; memcpy(&f, &u, 4); return f; };
  if (t->monotone && nthr >= 1 && table_host[0] > 0.f && table_host[nthr - 1] < 3.0e38f) {
    const int kmin = (int)(bits_of(table_host[0]) >> 20), kmax = (int)(bits_of(table_host[nthr - 1]) >> 20) + 1;
    const int keys = kmax - kmin + 1;
    if (keys <= 128) {
      bool ok = true;
      for (int k = 0; k < keys && ok; ++k) {
        const float lo = float_of((uint32_t)(kmin + k) << 20), hi = float_of((uint32_t)(kmin + k + 1) << 20);
        int below = 0, inside = 0;  // bucket 0 also takes everything below its lower edge, the last everything above
        for (int i = 0; i < nthr; ++i) {
          const bool is_below = k > 0 && table_host[i] < lo;
          const bool is_above = k < keys - 1 && !(table_host[i] < hi);
          below += is_below ? 1 : 0;
          inside += (!is_below && !is_above) ? 1 : 0;
        }
        ok = inside <= 2;
        t->lut[k] = (uint8_t)below;
      }
      if (ok) t->key_min = kmin, t->keys = keys;
    }
  }
  return STF_OK;
}

inline unsigned grid_x(int64_t work_items, int per_block, int batch) {
  int64_t blocks = (work_items + per_block - 1) / per_block;
  // enough CTAs to cover every SM several times, without one-vector-per-CTA grids on huge inputs
  int64_t cap = (int64_t)kNumSMs * 16 / (batch > 0 ? batch : 1);
  if (cap < 1) cap = 1;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

bool vec_ok(int64_t inner, int64_t s1, int64_t s2, const void *a, const void *b, const void *c,
            const void *d, const void *e, const void *f) {
  return (inner % 4 == 0) && (s1 % 4 == 0) && (s2 % 4 == 0) && aligned16(a) && aligned16(b) &&
         aligned16(c) && aligned16(d) && aligned16(e) && aligned16(f);
}

}  // namespace
}  // namespace stf

using namespace stf;

extern "C" const char *stf_version(void) { return "stf_b200 0.1.0 sm_100a"; }
extern "C" int64_t stf_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int stf_gaussian_compress_step(const float *y, int64_t y_batch_stride, const float *scales,
                                          const float *means, int32_t *symbols, int32_t *indexes,
                                          int64_t out_batch_stride, float *y_hat, int batch,
                                          int channels, int64_t plane, const float *table_host,
                                          int levels, float scale_bound, void *stream) {
  if (batch < 0 || channels < 0 || plane < 0) return STF_E_ARG;
  if (!y && !scales) return STF_E_ARG;
  if (scales && indexes == nullptr && y == nullptr) return STF_E_ARG;
  ScaleTable t;
  if (scales) {
    int rc = fill_table(&t, table_host, levels);
    if (rc) return rc;
  } else {
    t.levels = 1;
    t.monotone = 1;
    t.key_min = t.keys = 0;
    for (float &v : t.v) v = 0.f;
  }
  SliceGeom g{(int64_t)channels * plane, y_batch_stride, out_batch_stride};
  if (g.inner == 0 || batch == 0) return STF_OK;
  bool vec = vec_ok(g.inner, y_batch_stride, out_batch_stride, y, scales, means, symbols, indexes, y_hat);
  dim3 grid(grid_x(vec ? g.inner / 4 : g.inner, kThreads * (vec ? kVecPerThread : 1), batch), batch);
  cudaStream_t st = (cudaStream_t)stream;
  const bool idx_only = !y && !symbols && !y_hat;
#define LAUNCH(V, M)                                                                                                    \
  do {                                                                                                                  \
    if (idx_only)                                                                                                       \
      compress_step_kernel<V, M, true><<<grid, kThreads, 0, st>>>(y, scales, means, symbols, indexes, y_hat, g, scale_bound, t); \
    else                                                                                                                \
      compress_step_kernel<V, M, false><<<grid, kThreads, 0, st>>>(y, scales, means, symbols, indexes, y_hat, g, scale_bound, t); \
  } while (0)
  static const bool no_lut = getenv("STF_B200_NO_INDEX_LUT") != nullptr;  // (test hook: exercise the binary search)
  const int mode = !t.monotone ? 0 : (t.keys > 0 && !no_lut) ? 2 : 1;
  if (vec) {
    if (mode == 2) LAUNCH(true, 2); else if (mode == 1) LAUNCH(true, 1); else LAUNCH(true, 0);
  } else {
    if (mode == 2) LAUNCH(false, 2); else if (mode == 1) LAUNCH(false, 1); else LAUNCH(false, 0);
  }
#undef LAUNCH
  return check_launch();
}

extern "C" int stf_slice_step_nhwc(const stf_slice_args *a, void *stream) {
  if (!a || a->batch < 0 || a->channels <= 0 || a->channels > 32 || a->plane < 0) return STF_E_ARG;
  const bool forward = a->likelihood != nullptr;
  const bool encode = !forward && a->y != nullptr;
  const bool decode_fin = !forward && !encode && a->symbols_in != nullptr;
  if (forward && (!a->y || !a->scales)) return STF_E_ARG;
  if (encode && !a->symbols_out && !a->y_hat) return STF_E_ARG;
  if (decode_fin && !a->y_hat) return STF_E_ARG;
  if (!forward && !encode && !decode_fin && !(a->scales && a->indexes_out)) return STF_E_ARG;
  if (a->indexes_out && !a->scales) return STF_E_ARG;
  if (a->narrow && !a->overflow) return STF_E_ARG;
  ScaleTable t;
  if (a->scales && a->indexes_out) {
    int rc = fill_table(&t, a->table_host, a->levels);
    if (rc) return rc;
  } else {
    t.levels = 1, t.monotone = 1, t.key_min = t.keys = 0;
    for (float &v : t.v) v = 0.f;
  }
  if (a->batch == 0 || a->plane == 0) return STF_OK;
  SliceNhwc k{};
  k.y = a->y, k.y_ld = a->y_ld, k.scales = a->scales, k.s_ld = a->scales_ld, k.means = a->means, k.m_ld = a->means_ld;
  k.sym_in = decode_fin ? a->symbols_in : nullptr, k.sym_in_bstride = a->symbols_in_batch_stride;
  k.sym_out = encode ? a->symbols_out : nullptr, k.idx_out = forward ? nullptr : a->indexes_out, k.out_bstride = a->out_batch_stride;
  k.y_hat = a->y_hat, k.yh_ld = a->y_hat_ld, k.lik = a->likelihood, k.lik_bstride = a->likelihood_batch_stride;
  k.C = a->channels, k.plane = (int)a->plane, k.scale_bound = a->scale_bound, k.lik_bound = a->lik_bound, k.ste_round = a->ste_round;
  k.narrow = a->narrow, k.overflow = a->overflow;
  dim3 grid((unsigned)((a->plane + 31) / 32), (unsigned)a->batch);
  cudaStream_t st = (cudaStream_t)stream;
  static const bool no_lut = getenv("STF_B200_NO_INDEX_LUT") != nullptr;
  const int mode = !t.monotone ? 0 : (t.keys > 0 && !no_lut) ? 2 : 1;
  if (mode == 2) slice_step_nhwc_kernel<2><<<grid, 256, 0, st>>>(k, t);
  else if (mode == 1) slice_step_nhwc_kernel<1><<<grid, 256, 0, st>>>(k, t);
  else slice_step_nhwc_kernel<0><<<grid, 256, 0, st>>>(k, t);
  return check_launch();
}

extern "C" int stf_build_indexes(const float *scales, int32_t *indexes, int64_t n, const float *table_host,
                                 int levels, float scale_bound, void *stream) {
  if (!scales || !indexes || n < 0) return STF_E_ARG;
  // one "batch" of n elements; chunk so that `plane` stays below 2^31 per call
  return stf_gaussian_compress_step(nullptr, 0, scales, nullptr, nullptr, indexes, 0, nullptr, 1, 1, n,
                                    table_host, levels, scale_bound, stream);
}

extern "C" int stf_quantize_symbols(const float *x, const float *means, int32_t *symbols, int64_t n,
                                    void *stream) {
  if (!x || !symbols || n < 0) return STF_E_ARG;
  if (n == 0) return STF_OK;
  quantize_symbols_kernel<<<grid_x(n, kThreads, 1), kThreads, 0, (cudaStream_t)stream>>>(x, means, symbols, n);
  return check_launch();
}

extern "C" int stf_quantize_dequantize(const float *x, const float *means, float *out, int64_t n,
                                       void *stream) {
  if (!x || !out || n < 0) return STF_E_ARG;
  if (n == 0) return STF_OK;
  quantize_dequantize_kernel<<<grid_x(n, kThreads, 1), kThreads, 0, (cudaStream_t)stream>>>(x, means, out, n);
  return check_launch();
}

extern "C" int stf_dequantize(const int32_t *symbols, int64_t sym_batch_stride, const float *means,
                              float *y_hat, int batch, int channels, int64_t plane, void *stream) {
  if (!symbols || !y_hat || batch < 0 || channels < 0 || plane < 0) return STF_E_ARG;
  SliceGeom g{(int64_t)channels * plane, sym_batch_stride, 0};
  if (g.inner == 0 || batch == 0) return STF_OK;
  bool vec = vec_ok(g.inner, sym_batch_stride, 0, symbols, means, y_hat, nullptr, nullptr, nullptr);
  dim3 grid(grid_x(vec ? g.inner / 4 : g.inner, kThreads, batch), batch);
  if (vec)
    dequantize_kernel<true><<<grid, kThreads, 0, (cudaStream_t)stream>>>(symbols, means, y_hat, g);
  else
    dequantize_kernel<false><<<grid, kThreads, 0, (cudaStream_t)stream>>>(symbols, means, y_hat, g);
  return check_launch();
}

extern "C" int stf_gaussian_likelihood(const float *y, int64_t y_batch_stride, const float *scales,
                                       const float *means, float *y_hat, float *likelihood, int batch,
                                       int channels, int64_t plane, float scale_bound, float lik_bound,
                                       int ste_round, void *stream) {
  if (!y || !scales || !likelihood || batch < 0 || channels < 0 || plane < 0) return STF_E_ARG;
  SliceGeom g{(int64_t)channels * plane, y_batch_stride, 0};
  if (g.inner == 0 || batch == 0) return STF_OK;
  bool vec = vec_ok(g.inner, y_batch_stride, 0, y, scales, means, y_hat, likelihood, nullptr);
  dim3 grid(grid_x(vec ? g.inner / 4 : g.inner, kThreads * (vec ? 2 : 1), batch), batch);
#define LAUNCH(V, S)                                                              \
  gaussian_likelihood_kernel<V, S><<<grid, kThreads, 0, (cudaStream_t)stream>>>( \
      y, scales, means, y_hat, likelihood, g, scale_bound, lik_bound)
  if (vec) {
    if (ste_round) LAUNCH(true, true); else LAUNCH(true, false);
  } else {
    if (ste_round) LAUNCH(false, true); else LAUNCH(false, false);
  }
#undef LAUNCH
  return check_launch();
}

extern "C" int stf_entropy_bottleneck(const float *z, const float *params, float *z_hat, float *likelihood,
                                      int32_t *symbols, int batch, int channels, int64_t plane,
                                      float lik_bound, int ste_round, void *stream) {
  if (!z || !params || batch < 0 || channels < 0 || plane < 0) return STF_E_ARG;
  if (batch == 0 || channels == 0 || plane == 0) return STF_OK;
  if (channels > 65535 || batch > 65535) return STF_E_SHAPE;
  dim3 grid((unsigned)((plane + 127) / 128 > 64 ? 64 : (plane + 127) / 128), channels, batch);
  entropy_bottleneck_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(z, params, z_hat, likelihood, symbols,
                                                                   channels, plane, lik_bound, ste_round);
  return check_launch();
}
