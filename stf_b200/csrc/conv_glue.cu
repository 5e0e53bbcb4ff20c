// Element-wise glue around the cuDNN convolution stacks (the callers either side of
// the hot path).  ncu on one batch-32 compress+decompress showed 25 % of the GPU time in torch element-wise kernels:
// a broadcast bias add (strided, non-vectorised) + a GELU after every convolution, a LayerNorm launched with one CTA
// per 48-float row in PatchEmbed, and layout copies.  These two kernels replace them:
//
//   stf_bias_act        y = act(x + bias[c]) in place on an NHWC (channels_last) tensor   stf.py:510-548 (conv + GELU)
//   stf_layernorm_fwd   token-major LayerNorm over C <= 768 channels                        stf.py:375-379 (PatchEmbed.norm)
#include <math.h>

#include "common.cuh"

namespace stf {
namespace {

// torch's exact GELU (aten/src/ATen/native/cuda/ActivationGeluKernel.cu): x * 0.5 * (1 + erf(x * M_SQRT1_2))
__device__ __forceinline__ float gelu_exact(float x) { return x * 0.5f * (1.0f + erff(x * 0.70710678118654752440f)); }

template <int kAct>
__global__ void __launch_bounds__(256)
bias_act_kernel(float *__restrict__ x, const float *__restrict__ bias, int c4, int64_t n4) {
  // element e (float4 index) belongs to channels 4 * (e % c4) .. +3 of its pixel
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<float4 *>(x)[i];
    const float4 b = __ldg(reinterpret_cast<const float4 *>(bias) + (int)(i % c4));
    v.x += b.x, v.y += b.y, v.z += b.z, v.w += b.w;
    if (kAct == 1) v.x = gelu_exact(v.x), v.y = gelu_exact(v.y), v.z = gelu_exact(v.z), v.w = gelu_exact(v.w);
    reinterpret_cast<float4 *>(x)[i] = v;
  }
}

constexpr int kLnWarps = 8;
constexpr int kLnMaxPerLane = 24;

// PER = elements per lane (compile time: the row loops are fully unrolled with no dead iterations)
template <int PER>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_fwd_kernel(const float *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ beta,
                     float *__restrict__ y, int64_t M, int C, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int per = PER;
  const float inv_c = 1.0f / (float)C;
  for (int64_t row = (int64_t)blockIdx.x * kLnWarps + warp; row < M; row += (int64_t)gridDim.x * kLnWarps) {
    const float *xr = x + row * C;
    float v[PER];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = lane + 32 * i;
      v[i] = (i < per && c < C) ? xr[c] : 0.f;
      s += v[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * inv_c;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = lane + 32 * i;
      const float d = (i < per && c < C) ? v[i] - mean : 0.f;
      q = fmaf(d, d, q);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * inv_c + eps);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = lane + 32 * i;
      if (i < per && c < C) y[row * C + c] = fmaf((v[i] - mean) * rstd, __ldg(gamma + c), __ldg(beta + c));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// PatchEmbed (stf.py:350-381): Conv2d(in_chans -> E, kernel = stride = patch) + LayerNorm(E) on the NCHW image, emitted
// token-major (B * Wh * Ww, E) = the NHWC layout every later kernel works in.  K = in_chans * patch^2 is 12 for STF: far
// below a tensor-core tile, so this is an fp32 FFMA kernel, one thread per token, weights broadcast from shared memory, the
// token's E outputs and its LayerNorm statistics in registers, output staged through shared memory for coalesced stores.
// HBM-bound: 4 * (K + E) B per token.  Fixed summation order per token: batch-invariant.
// ---------------------------------------------------------------------------------------------
constexpr int kPeTokens = 64;     // tokens (= threads) per block, consecutive along the token row
constexpr int kPeMaxE = 96, kPeMaxK = 48;

template <int E>
__global__ void __launch_bounds__(kPeTokens)
patch_embed_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias,
                   const float *__restrict__ gamma, const float *__restrict__ beta, float *__restrict__ out, int B, int Cin,
                   int H, int W, int P, int Wh, int Ww, float eps) {
  __shared__ float ws[kPeMaxK * E];          // [k][e]: all threads read the same word (broadcast)
  __shared__ float stage[kPeTokens * (E + 1)];
  const int K = Cin * P * P;
  for (int i = threadIdx.x; i < K * E; i += kPeTokens) {
    const int k = i / E, e = i - k * E;
    ws[i] = w[e * K + k];                    // conv weight (E, Cin, P, P) contiguous: k = (c * P + dy) * P + dx
  }
  __syncthreads();
  const int64_t tokens = (int64_t)B * Wh * Ww;
  const int64_t tok0 = (int64_t)blockIdx.x * kPeTokens;
  const int64_t tok = tok0 + threadIdx.x;
  if (tok < tokens) {
    const int b = (int)(tok / ((int64_t)Wh * Ww));
    const int rem = (int)(tok - (int64_t)b * Wh * Ww);
    const int ty = rem / Ww, tx = rem - ty * Ww;
    float acc[E];
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = bias ? __ldg(bias + e) : 0.f;
    for (int c = 0; c < Cin; ++c)
      for (int dy = 0; dy < P; ++dy)
        for (int dx = 0; dx < P; ++dx) {
          const int yy = ty * P + dy, xx = tx * P + dx;   // zero padding to a multiple of the patch (stf.py:369-372)
          const float v = (yy < H && xx < W) ? __ldg(x + (((int64_t)b * Cin + c) * H + yy) * W + xx) : 0.f;
          const float *wk = ws + ((c * P + dy) * P + dx) * E;
#pragma unroll
          for (int e = 0; e < E; ++e) acc[e] = fmaf(v, wk[e], acc[e]);
        }
    if (gamma) {   // LayerNorm over the E channels (two-pass in registers)
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) s += acc[e];
      const float mean = s * (1.0f / E);
      float q = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const float d = acc[e] - mean;
        q = fmaf(d, d, q);
      }
      const float rstd = rsqrtf(q * (1.0f / E) + eps);
#pragma unroll
      for (int e = 0; e < E; ++e) acc[e] = fmaf((acc[e] - mean) * rstd, __ldg(gamma + e), __ldg(beta + e));
    }
#pragma unroll
    for (int e = 0; e < E; ++e) stage[threadIdx.x * (E + 1) + e] = acc[e];
  }
  __syncthreads();
  const int64_t left = tokens - tok0;
  const int n = (int)(left < kPeTokens ? left : kPeTokens) * E;
  for (int i = threadIdx.x; i < n; i += kPeTokens) {
    const int t = i / E, e = i - t * E;
    out[tok0 * E + i] = stage[t * (E + 1) + e];
  }
}

}  // namespace
}  // namespace stf

using namespace stf;

extern "C" int stf_patch_embed(const float *x, const float *weight, const float *bias, const float *ln_gamma,
                               const float *ln_beta, float *tokens, int batch, int in_chans, int H, int W, int patch,
                               int embed_dim, float ln_eps, void *stream) {
  if (!x || !weight || !tokens || batch < 0 || in_chans <= 0 || H <= 0 || W <= 0 || patch <= 0) return STF_E_ARG;
  if ((ln_gamma == nullptr) != (ln_beta == nullptr)) return STF_E_ARG;
  if (in_chans * patch * patch > kPeMaxK || (embed_dim != 48 && embed_dim != 96)) return STF_E_SHAPE;
  const int Wh = (H + patch - 1) / patch, Ww = (W + patch - 1) / patch;
  const int64_t tokens_n = (int64_t)batch * Wh * Ww;
  if (tokens_n == 0) return STF_OK;
  const unsigned blocks = (unsigned)((tokens_n + kPeTokens - 1) / kPeTokens);
  if (embed_dim == 48)
    patch_embed_kernel<48><<<blocks, kPeTokens, 0, (cudaStream_t)stream>>>(x, weight, bias, ln_gamma, ln_beta, tokens, batch,
                                                                          in_chans, H, W, patch, Wh, Ww, ln_eps);
  else
    patch_embed_kernel<96><<<blocks, kPeTokens, 0, (cudaStream_t)stream>>>(x, weight, bias, ln_gamma, ln_beta, tokens, batch,
                                                                          in_chans, H, W, patch, Wh, Ww, ln_eps);
  return check_launch();
}

extern "C" int stf_bias_act(float *x, const float *bias, int channels, int64_t n, int act, void *stream) {
  if (!x || !bias || channels <= 0 || n < 0 || (act != 0 && act != 1)) return STF_E_ARG;
  if (channels % 4 != 0 || n % channels != 0) return STF_E_SHAPE;
  if (!aligned16(x) || !aligned16(bias)) return STF_E_ALIGN;
  if (n == 0) return STF_OK;
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
  if (act == 1)
    bias_act_kernel<1><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, bias, channels / 4, n4);
  else
    bias_act_kernel<0><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, bias, channels / 4, n4);
  return check_launch();
}

extern "C" int stf_layernorm_fwd(const float *x, const float *gamma, const float *beta, float *y, int64_t M, int C,
                                 float eps, void *stream) {
  if (!x || !gamma || !beta || !y || M < 0 || C <= 0) return STF_E_ARG;
  if (C > 32 * kLnMaxPerLane) return STF_E_SHAPE;
  if (M == 0) return STF_OK;
  int64_t blocks = (M + kLnWarps - 1) / kLnWarps;
  if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
  const int per = (C + 31) / 32;
#define STF_LN(P) layernorm_fwd_kernel<P><<<(unsigned)blocks, kLnWarps * 32, 0, (cudaStream_t)stream>>>(x, gamma, beta, y, M, C, eps)
  if (per <= 2) STF_LN(2);
  else if (per <= 3) STF_LN(3);
  else if (per <= 6) STF_LN(6);
  else if (per <= 12) STF_LN(12);
  else STF_LN(24);
#undef STF_LN
  return check_launch();
}
