// Backward kernels of the STF rate-distortion training step (BASELINE config 5, SURVEY.md section 8 rows
// a1/a4/a5 backward, a11 "noise", a13 training forward, a15, a16):
//
//   stf_window_attention_bwd     d(softmax(q k^T + B + mask) v) for 16-token windows   stf.py:100-118 (autograd)
//   stf_layernorm_bwd            LayerNorm backward + residual-path add + dgamma/dbeta  stf.py:155,197 (autograd)
//   stf_gelu_bwd                 exact-erf GELU backward                                stf.py:35-36 (autograd)
//   stf_gaussian_likelihood_train{,_bwd}
//                                GaussianConditional.forward in training mode ("noise" quantisation) with the
//                                LowerBound custom gradient                             entropy_models.py:131-135,
//                                                                                       645-659; ops/bound_ops.py:21-27
//
// The GEMMs of the backward pass (dX = dY . W) run on the same tcgen05 kernel as the forward pass (stf_linear with
// the transposed weight packed); weight gradients (dW = dY^T . X) are plain library GEMMs (cuBLAS through torch).
// Reductions over rows / windows (dgamma, dbeta, d bias_table) are two-stage and atomic-free: every CTA writes its
// partial sums to its own slot, the host side adds the slots in a fixed order -> deterministic gradients.
#include <math.h>

#include "common.cuh"

namespace stf {
namespace {

constexpr float kMaskValue = -100.0f;  // stf.py:334

// ---------------------------------------------------------------------------------------------
// Window attention backward, 4x4 windows.  One thread per (window, head, token); a CTA owns `wpc` windows.
// Shared memory: the qkv tile, the dO tile, and per (window, head) the 16x16 P and dS matrices.
//   dP[n][m] = dO_n . v_m          dS[n][m] = P[n][m] (dP[n][m] - sum_j P[n][j] dP[n][j])
//   dq_n = q_scale * sum_m dS[n][m] k_m      dk_m = sum_n dS[n][m] q_n       dv_m = sum_n P[n][m] dO_n
//   dB[rel(n,m)][head] += dS[n][m]           (q in the tile is already scaled, as in the forward kernel)
// ---------------------------------------------------------------------------------------------
constexpr int kPS = 17;  // padded row stride of the P / dS matrices (column reads in the second phase)

template <int D>
__global__ void __launch_bounds__(384)
window_attention16_bwd_kernel(const float *__restrict__ qkv, const float *__restrict__ dout,
                              const float *__restrict__ bias_table, float *__restrict__ dqkv,
                              float *__restrict__ dbias_part, int64_t num_windows, int C, int heads, int shift,
                              int Hp, int Wp, int wpc, float q_scale) {
  constexpr int WS = 4, N = 16;
  extern __shared__ __align__(16) float sm[];
  const int ld = 3 * C;
  float *tile = sm;                                  // [wpc][16][3C]
  float *dO = tile + (size_t)wpc * N * ld;           // [wpc][16][C]
  float *Pm = dO + (size_t)wpc * N * C;              // [wpc*heads][16][kPS]
  float *dS = Pm + (size_t)wpc * heads * N * kPS;    // [wpc*heads][16][kPS]
  const int64_t win0 = (int64_t)blockIdx.x * wpc;
  const int nwin = (int)((num_windows - win0) < wpc ? (num_windows - win0) : wpc);

  {  // cooperative, coalesced loads of the two tiles
    const float4 *src = reinterpret_cast<const float4 *>(qkv + win0 * N * (int64_t)ld);
    float4 *dst = reinterpret_cast<float4 *>(tile);
    for (int i = threadIdx.x; i < nwin * N * ld / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    const float4 *src2 = reinterpret_cast<const float4 *>(dout + win0 * N * (int64_t)C);
    float4 *dst2 = reinterpret_cast<float4 *>(dO);
    for (int i = threadIdx.x; i < nwin * N * C / 4; i += blockDim.x) dst2[i] = __ldg(src2 + i);
  }
  __syncthreads();

  const int n = threadIdx.x % N;
  const int pair = threadIdx.x / N;  // (local window, head)
  const int wl = pair / heads, head = pair - wl * heads;
  const bool active = wl < nwin && pair < wpc * heads;
  const int hn = n / WS, wn = n % WS;
  if (active) {
    const int64_t win = win0 + wl;
    const float *base = tile + (size_t)wl * N * ld + head * D;
    const float *dob = dO + (size_t)wl * N * C + head * D;
    float q[D], go[D];
#pragma unroll
    for (int j = 0; j < D; ++j) q[j] = base[n * ld + j], go[j] = dob[n * C + j];
    int my_label = 0, wy = 0, wx = 0;
    if (shift > 0) {
      const int nWw = Wp / WS, nW = (Hp / WS) * nWw;
      const int wi = (int)(win % nW);
      wy = wi / nWw, wx = wi - wy * nWw;
      const int hs = wy * WS + hn, wsft = wx * WS + wn;
      my_label = 3 * (hs < Hp - WS ? 0 : (hs < Hp - shift ? 1 : 2)) + (wsft < Wp - WS ? 0 : (wsft < Wp - shift ? 1 : 2));
    }
    float s[N], dp[N];
    float smax = -INFINITY;
    const float *kbase = base + C, *vbase = base + 2 * C;
#pragma unroll
    for (int m = 0; m < N; ++m) {
      float acc = 0.f, g = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) {
        acc = fmaf(q[j], kbase[m * ld + j], acc);
        g = fmaf(go[j], vbase[m * ld + j], g);
      }
      const int hm = m / WS, wm = m % WS;
      const int rel = (hn - hm + WS - 1) * (2 * WS - 1) + (wn - wm + WS - 1);
      acc += __ldg(bias_table + rel * heads + head);
      if (shift > 0) {
        const int hs = wy * WS + hm, wsft = wx * WS + wm;
        const int lab = 3 * (hs < Hp - WS ? 0 : (hs < Hp - shift ? 1 : 2)) + (wsft < Wp - WS ? 0 : (wsft < Wp - shift ? 1 : 2));
        if (lab != my_label) acc += kMaskValue;
      }
      s[m] = acc;
      dp[m] = g;
      smax = fmaxf(smax, acc);
    }
    float denom = 0.f;
#pragma unroll
    for (int m = 0; m < N; ++m) {
      s[m] = expf(s[m] - smax);
      denom += s[m];
    }
    const float inv = 1.0f / denom;
    float dot = 0.f;
#pragma unroll
    for (int m = 0; m < N; ++m) {
      s[m] *= inv;  // P[n][m]
      dot = fmaf(s[m], dp[m], dot);
    }
    float dq[D];
#pragma unroll
    for (int j = 0; j < D; ++j) dq[j] = 0.f;
    float *prow = Pm + ((size_t)pair * N + n) * kPS, *drow = dS + ((size_t)pair * N + n) * kPS;
#pragma unroll
    for (int m = 0; m < N; ++m) {
      const float ds = s[m] * (dp[m] - dot);
      prow[m] = s[m];
      drow[m] = ds;
#pragma unroll
      for (int j = 0; j < D; ++j) dq[j] = fmaf(ds, kbase[m * ld + j], dq[j]);
    }
    float *dst = dqkv + ((win0 + wl) * N + n) * (int64_t)ld + head * D;
#pragma unroll
    for (int j = 0; j < D; j += 4)
      *reinterpret_cast<float4 *>(dst + j) = make_float4(q_scale * dq[j], q_scale * dq[j + 1], q_scale * dq[j + 2], q_scale * dq[j + 3]);
  }
  __syncthreads();
  if (active) {  // second phase: this thread is key / value token m = n
    const float *base = tile + (size_t)wl * N * ld + head * D;
    const float *dob = dO + (size_t)wl * N * C + head * D;
    const float *pcol = Pm + (size_t)pair * N * kPS + n, *dcol = dS + (size_t)pair * N * kPS + n;
    float dk[D], dv[D];
#pragma unroll
    for (int j = 0; j < D; ++j) dk[j] = 0.f, dv[j] = 0.f;
#pragma unroll
    for (int r = 0; r < N; ++r) {
      const float ds = dcol[r * kPS], p = pcol[r * kPS];
#pragma unroll
      for (int j = 0; j < D; ++j) {
        dk[j] = fmaf(ds, base[r * ld + j], dk[j]);
        dv[j] = fmaf(p, dob[r * C + j], dv[j]);
      }
    }
    float *dst = dqkv + ((win0 + wl) * N + n) * (int64_t)ld + head * D;
#pragma unroll
    for (int j = 0; j < D; j += 4) {
      *reinterpret_cast<float4 *>(dst + C + j) = make_float4(dk[j], dk[j + 1], dk[j + 2], dk[j + 3]);
      *reinterpret_cast<float4 *>(dst + 2 * C + j) = make_float4(dv[j], dv[j + 1], dv[j + 2], dv[j + 3]);
    }
  }
  // d bias_table partial of this CTA: entry (rel, head) sums dS over the CTA's windows and the (n, m) pairs with
  // that relative position, in a fixed order.
  constexpr int R = 2 * WS - 1;
  for (int e = threadIdx.x; e < R * R * heads; e += blockDim.x) {
    const int rel = e / heads, h = e - rel * heads;
    const int dh = rel / R - (WS - 1), dw = rel % R - (WS - 1);
    float acc = 0.f;
    for (int w = 0; w < nwin; ++w) {
      const float *d = dS + (size_t)(w * heads + h) * N * kPS;
      for (int hm = 0; hm < WS; ++hm) {
        const int hq = hm + dh;
        if (hq < 0 || hq >= WS) continue;
        for (int wm = 0; wm < WS; ++wm) {
          const int wq = wm + dw;
          if (wq < 0 || wq >= WS) continue;
          acc += d[(hq * WS + wq) * kPS + hm * WS + wm];
        }
      }
    }
    dbias_part[(int64_t)blockIdx.x * R * R * heads + e] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// Window attention backward for 8x8 windows (WACNN's first attention block: 64 tokens, 8 heads of 24 channels).
// One CTA per (window, head), one thread per token.  Shared memory: the head's q / k / v / dO slices and the 64x64 P and
// dS matrices.  Same formulas as the 16-token kernel; scores are staged in P between the softmax passes so that no
// 64-entry array lives in registers.  d bias_table partial of the CTA goes to slot (window, rel, head).
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(64)
window_attention64_bwd_kernel(const float *__restrict__ qkv, const float *__restrict__ dout,
                              const float *__restrict__ bias_table, float *__restrict__ dqkv,
                              float *__restrict__ dbias_part, int C, int heads, int shift, int Hp, int Wp, float q_scale) {
  constexpr int WS = 8, N = 64, R = 2 * WS - 1, PS = N + 1;
  extern __shared__ __align__(16) float sm[];
  float *qs = sm, *ks = qs + N * D, *vs = ks + N * D, *gs = vs + N * D;  // [64][D] each
  float *Pm = gs + N * D, *dS = Pm + N * PS;                              // [64][65] each
  const int64_t win = blockIdx.x / heads;
  const int head = blockIdx.x % heads;
  const int ld = 3 * C;
  const int n = threadIdx.x;
  {
    const float *row = qkv + (win * N + n) * (int64_t)ld + head * D;
    const float *grow = dout + (win * N + n) * (int64_t)C + head * D;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      qs[n * D + j] = row[j];
      ks[n * D + j] = row[C + j];
      vs[n * D + j] = row[2 * C + j];
      gs[n * D + j] = grow[j];
    }
  }
  __syncthreads();
  const int hn = n / WS, wn = n % WS;
  int my_label = 0, wy = 0, wx = 0;
  if (shift > 0) {
    const int nWw = Wp / WS, nW = (Hp / WS) * nWw;
    const int wi = (int)(win % nW);
    wy = wi / nWw, wx = wi - wy * nWw;
    const int hs = wy * WS + hn, wsft = wx * WS + wn;
    my_label = 3 * (hs < Hp - WS ? 0 : (hs < Hp - shift ? 1 : 2)) + (wsft < Wp - WS ? 0 : (wsft < Wp - shift ? 1 : 2));
  }
  float q[D], go[D];
#pragma unroll
  for (int j = 0; j < D; ++j) q[j] = qs[n * D + j], go[j] = gs[n * D + j];
  float *prow = Pm + n * PS, *drow = dS + n * PS;
  float smax = -INFINITY;
  for (int m = 0; m < N; ++m) {
    float acc = 0.f, g = 0.f;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      acc = fmaf(q[j], ks[m * D + j], acc);
      g = fmaf(go[j], vs[m * D + j], g);
    }
    const int hm = m / WS, wm = m % WS;
    acc += __ldg(bias_table + ((hn - hm + WS - 1) * R + (wn - wm + WS - 1)) * heads + head);
    if (shift > 0) {
      const int hs = wy * WS + hm, wsft = wx * WS + wm;
      const int lab = 3 * (hs < Hp - WS ? 0 : (hs < Hp - shift ? 1 : 2)) + (wsft < Wp - WS ? 0 : (wsft < Wp - shift ? 1 : 2));
      if (lab != my_label) acc += kMaskValue;
    }
    prow[m] = acc;   // raw score
    drow[m] = g;     // dP
    smax = fmaxf(smax, acc);
  }
  float denom = 0.f;
  for (int m = 0; m < N; ++m) {
    const float e = expf(prow[m] - smax);
    prow[m] = e;
    denom += e;
  }
  const float inv = 1.0f / denom;
  float dot = 0.f;
  for (int m = 0; m < N; ++m) {
    const float p = prow[m] * inv;
    prow[m] = p;
    dot = fmaf(p, drow[m], dot);
  }
  float dq[D];
#pragma unroll
  for (int j = 0; j < D; ++j) dq[j] = 0.f;
  for (int m = 0; m < N; ++m) {
    const float ds = prow[m] * (drow[m] - dot);
    drow[m] = ds;
#pragma unroll
    for (int j = 0; j < D; ++j) dq[j] = fmaf(ds, ks[m * D + j], dq[j]);
  }
  float *dst = dqkv + (win * N + n) * (int64_t)ld + head * D;
#pragma unroll
  for (int j = 0; j < D; ++j) dst[j] = q_scale * dq[j];
  __syncthreads();
  {  // second phase: this thread is key / value token m = n
    float dk[D], dv[D];
#pragma unroll
    for (int j = 0; j < D; ++j) dk[j] = 0.f, dv[j] = 0.f;
    for (int r = 0; r < N; ++r) {
      const float ds = dS[r * PS + n], p = Pm[r * PS + n];
#pragma unroll
      for (int j = 0; j < D; ++j) {
        dk[j] = fmaf(ds, qs[r * D + j], dk[j]);
        dv[j] = fmaf(p, gs[r * D + j], dv[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < D; ++j) dst[C + j] = dk[j], dst[2 * C + j] = dv[j];
  }
  for (int e = threadIdx.x; e < R * R; e += blockDim.x) {
    const int dh = e / R - (WS - 1), dw = e % R - (WS - 1);
    float acc = 0.f;
    for (int hm = 0; hm < WS; ++hm) {
      const int hq = hm + dh;
      if (hq < 0 || hq >= WS) continue;
      for (int wm = 0; wm < WS; ++wm) {
        const int wq = wm + dw;
        if (wq < 0 || wq >= WS) continue;
        acc += dS[(hq * WS + wq) * PS + hm * WS + wm];
      }
    }
    dbias_part[((int64_t)win * R * R + e) * heads + head] = acc;
  }
}

template <int D>
int launch_attn64_bwd(const float *qkv, const float *dout, const float *bias_table, float *dqkv, float *dbias_part,
                      int64_t num_windows, int C, int heads, int shift, int Hp, int Wp, float q_scale, cudaStream_t st) {
  const size_t smem = ((size_t)4 * 64 * D + (size_t)2 * 64 * 65) * 4;
  static std::atomic<int> attr_set{0};
  if (!attr_set.load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(window_attention64_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set.store(1, std::memory_order_release);
  }
  const int64_t blocks = num_windows * heads;
  if (blocks > 0x7fffffffLL) return STF_E_SHAPE;
  window_attention64_bwd_kernel<D><<<(unsigned)blocks, 64, smem, st>>>(qkv, dout, bias_table, dqkv, dbias_part, C, heads,
                                                                       shift, Hp, Wp, q_scale);
  return check_launch();
}

template <int D>
int launch_attn_bwd(const float *qkv, const float *dout, const float *bias_table, float *dqkv, float *dbias_part,
                    int64_t num_windows, int C, int heads, int shift, int Hp, int Wp, int wpc, float q_scale,
                    cudaStream_t st) {
  const size_t smem = ((size_t)wpc * 16 * 4 * C + (size_t)2 * wpc * heads * 16 * kPS) * 4;
  static std::atomic<int> attr_set{0};
  if (!attr_set.load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(window_attention16_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set.store(1, std::memory_order_release);
  }
  const int64_t blocks = (num_windows + wpc - 1) / wpc;
  const int threads = (wpc * heads * 16 + 31) / 32 * 32;
  window_attention16_bwd_kernel<D><<<(unsigned)blocks, threads, smem, st>>>(qkv, dout, bias_table, dqkv, dbias_part,
                                                                          num_windows, C, heads, shift, Hp, Wp, wpc, q_scale);
  return check_launch();
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward.  One warp per row, rows strided over the grid.
//   xhat = (x - mean) rstd;  d = g o gamma;  dx = rstd (d - mean_c(d) - xhat mean_c(d o xhat)) (+ res)
//   dgamma += g o xhat;  dbeta += g;   optionally xn = xhat o gamma + beta (the wgrad operand of the Linear behind)
// ---------------------------------------------------------------------------------------------
constexpr int kLnWarps = 8;
constexpr int kLnMaxPerLane = 24;  // C <= 768

// PER = elements per lane (compile time: the row loops are fully unrolled with no dead iterations)
template <int PER>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_bwd_kernel(const float *__restrict__ x, const float *__restrict__ g, const float *__restrict__ gamma,
                     const float *__restrict__ beta, const float *__restrict__ res, float *__restrict__ dx,
                     float *__restrict__ xn, float *__restrict__ part, int64_t M, int C, float eps) {
  extern __shared__ float red[];  // [kLnWarps][2][C]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int per = PER;
  constexpr int kLnMaxPerLane = PER;   // (shadows the global bound inside this instance)
  float gam[kLnMaxPerLane], bet[kLnMaxPerLane], dgam[kLnMaxPerLane], dbet[kLnMaxPerLane];
#pragma unroll
  for (int i = 0; i < kLnMaxPerLane; ++i) {
    const int c = lane + 32 * i;
    gam[i] = (i < per && c < C) ? gamma[c] : 0.f;
    bet[i] = (i < per && c < C && beta) ? beta[c] : 0.f;
    dgam[i] = 0.f, dbet[i] = 0.f;
  }
  const float inv_c = 1.0f / (float)C;
  for (int64_t row = (int64_t)blockIdx.x * kLnWarps + warp; row < M; row += (int64_t)gridDim.x * kLnWarps) {
    const float *xr = x + row * C, *gr = g + row * C;
    float xv[kLnMaxPerLane], gv[kLnMaxPerLane];
    float s1 = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) {
      const int c = lane + 32 * i;
      const bool ok = i < per && c < C;
      xv[i] = ok ? xr[c] : 0.f;
      gv[i] = ok ? gr[c] : 0.f;
      s1 += xv[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    const float mean = s1 * inv_c;
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) {
      const int c = lane + 32 * i;
      const float d = (i < per && c < C) ? xv[i] - mean : 0.f;
      s2 = fmaf(d, d, s2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    const float rstd = rsqrtf(s2 * inv_c + eps);
    float a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) {
      const float xh = (xv[i] - mean) * rstd;
      const float d = gv[i] * gam[i];
      a1 += d;
      a2 = fmaf(d, xh, a2);
      dgam[i] = fmaf(gv[i], xh, dgam[i]);
      dbet[i] += gv[i];
      xv[i] = xh;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    a1 *= inv_c, a2 *= inv_c;
#pragma unroll
    for (int i = 0; i < kLnMaxPerLane; ++i) {
      const int c = lane + 32 * i;
      if (i < per && c < C) {
        float v = rstd * (gv[i] * gam[i] - a1 - xv[i] * a2);
        if (res) v += res[row * C + c];
        dx[row * C + c] = v;
        if (xn) xn[row * C + c] = fmaf(xv[i], gam[i], bet[i]);
      }
    }
  }
  // CTA partial of dgamma / dbeta: warps are added in index order
#pragma unroll
  for (int i = 0; i < kLnMaxPerLane; ++i) {
    const int c = lane + 32 * i;
    if (i < per && c < C) red[(warp * 2) * C + c] = dgam[i], red[(warp * 2 + 1) * C + c] = dbet[i];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * C; e += blockDim.x) {
    float acc = 0.f;
    for (int w = 0; w < kLnWarps; ++w) acc += red[w * 2 * C + e];
    part[(int64_t)blockIdx.x * 2 * C + e] = acc;
  }
}

// GELU'(p) = Phi(p) + p phi(p)
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const float *__restrict__ p, const float *__restrict__ dh, float *__restrict__ dp, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 pv = ldg_stream(reinterpret_cast<const float4 *>(p) + i);
    const float4 gv = ldg_stream(reinterpret_cast<const float4 *>(dh) + i);
    auto f = [](float x, float g) {
      const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
      const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
      return g * fmaf(x, pdf, cdf);
    };
    stg_stream(reinterpret_cast<float4 *>(dp) + i, make_float4(f(pv.x, gv.x), f(pv.y, gv.y), f(pv.z, gv.z), f(pv.w, gv.w)));
  }
}

// ---------------------------------------------------------------------------------------------
// GaussianConditional, training mode.  v = |y + noise - mu|, sigma = max(scale, b),
//   lik = max(Phi((.5 - v)/sigma) - Phi((-.5 - v)/sigma), 1e-9)           entropy_models.py:626-659
// Backward with the LowerBound rule (ops/bound_ops.py:25-27): the gradient passes where x >= bound or grad < 0.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float std_cdf(float x) { return 0.5f * erfcf(-0.70710678118654752440f * x); }
__device__ __forceinline__ float std_pdf(float x) { return 0.39894228040143267794f * expf(-0.5f * x * x); }

__global__ void __launch_bounds__(256)
gaussian_train_fwd_kernel(const float *__restrict__ y, const float *__restrict__ scales, const float *__restrict__ means,
                          const float *__restrict__ noise, float *__restrict__ lik, int64_t n, float scale_bound,
                          float lik_bound) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float mu = means ? means[i] : 0.f;
    float out = y[i] + (noise ? noise[i] : 0.f);   // quantize("noise") ignores the means (entropy_models.py:131-135)
    const float v = fabsf(out - mu);
    const float sc = scales[i];
    const float s = sc < scale_bound ? scale_bound : sc;
    const float l = std_cdf((0.5f - v) / s) - std_cdf((-0.5f - v) / s);
    lik[i] = l < lik_bound ? lik_bound : l;
  }
}

__global__ void __launch_bounds__(256)
gaussian_train_bwd_kernel(const float *__restrict__ y, const float *__restrict__ scales, const float *__restrict__ means,
                          const float *__restrict__ noise, const float *__restrict__ dlik, float *__restrict__ dy,
                          float *__restrict__ dscale, float *__restrict__ dmean, int64_t n, float scale_bound,
                          float lik_bound) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float mu = means ? means[i] : 0.f;
    const float t = y[i] + (noise ? noise[i] : 0.f) - mu;
    const float v = fabsf(t);
    const float sc = scales[i];
    const bool clamped = sc < scale_bound;
    const float s = clamped ? scale_bound : sc;
    const float u = (0.5f - v) / s, lo = (-0.5f - v) / s;
    const float l = std_cdf(u) - std_cdf(lo);
    float g = dlik[i];
    if (!(l >= lik_bound || g < 0.f)) g = 0.f;        // LowerBound(likelihood, 1e-9)
    const float pu = std_pdf(u), pl = std_pdf(lo);
    const float dv = g * (pl - pu) / s;               // d lik / d v
    const float sgn = t > 0.f ? 1.f : (t < 0.f ? -1.f : 0.f);   // torch.abs backward: sign(0) = 0
    const float dt = dv * sgn;
    float ds = g * (pl * lo - pu * u) / s;            // d lik / d sigma
    if (!(sc >= scale_bound || ds < 0.f)) ds = 0.f;   // LowerBound(scale, 0.11)
    dy[i] = dt;
    if (dmean) dmean[i] = -dt;
    dscale[i] = ds;
  }
}

// Column sums of a row-major (M, C) matrix (bias gradients), two-stage and atomic-free: CTA b sums its row range into
// part[b][C].  Threads form (256 / C4) row lanes x C4 float4 columns; consecutive threads read consecutive 16-byte
// segments of a row.
__global__ void __launch_bounds__(256)
colsum_kernel(const float *__restrict__ a, float *__restrict__ part, int64_t M, int C) {
  extern __shared__ float4 acc_s[];  // [row lanes][C4]
  const int c4 = C >> 2;
  const int lanes = 256 / c4;
  const int col = threadIdx.x % c4, rl = threadIdx.x / c4;
  const int64_t rows_per_cta = (M + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta, r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rl < lanes)
    for (int64_t r = r0 + rl; r < r1; r += lanes) {
      const float4 v = ldg_stream(reinterpret_cast<const float4 *>(a + r * C) + col);
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
  if (rl < lanes) acc_s[rl * c4 + col] = acc;
  __syncthreads();
  for (int i = threadIdx.x; i < c4; i += blockDim.x) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int l = 0; l < lanes; ++l) {
      const float4 v = acc_s[l * c4 + i];
      t.x += v.x, t.y += v.y, t.z += v.z, t.w += v.w;
    }
    reinterpret_cast<float4 *>(part + (int64_t)blockIdx.x * C)[i] = t;
  }
}

inline unsigned ew_grid(int64_t n, int per_block) {
  int64_t b = (n + per_block - 1) / per_block;
  const int64_t cap = (int64_t)kNumSMs * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace
}  // namespace stf

using namespace stf;

extern "C" int stf_attention_bwd_ctas(int64_t num_windows, int C, int heads, int *wpc_out) {
  if (num_windows < 0 || C <= 0 || heads <= 0 || C % heads) return STF_E_ARG;
  int wpc = 384 / (heads * 16);
  const size_t per_win = ((size_t)16 * 4 * C + (size_t)2 * heads * 16 * kPS) * 4;
  const int by_smem = (int)((200 * 1024) / per_win);
  if (wpc > by_smem) wpc = by_smem;
  if (wpc < 1) return STF_E_SHAPE;
  if (wpc_out) *wpc_out = wpc;
  return (int)((num_windows + wpc - 1) / wpc);
}

extern "C" int stf_attention_bwd_slots(int64_t num_windows, int C, int heads, int ws) {
  if (ws == 8) return (num_windows < 0 || num_windows > 0x7fffffffLL) ? STF_E_ARG : (int)num_windows;
  if (ws != 4) return STF_E_SHAPE;
  return stf_attention_bwd_ctas(num_windows, C, heads, nullptr);
}

extern "C" int stf_window_attention_bwd(const float *qkv, const float *dout, const float *bias_table, float *dqkv,
                                        float *dbias_partials, int64_t num_windows, int C, int heads, int ws, int shift,
                                        int Hp, int Wp, float q_scale, void *stream) {
  if (!qkv || !dout || !bias_table || !dqkv || !dbias_partials || num_windows < 0 || C <= 0 || heads <= 0) return STF_E_ARG;
  if (num_windows == 0) return STF_OK;
  if ((ws != 4 && ws != 8) || C % heads != 0 || shift < 0 || shift >= ws) return STF_E_SHAPE;
  if (shift > 0 && (Hp <= 0 || Wp <= 0 || Hp % ws != 0 || Wp % ws != 0)) return STF_E_SHAPE;
  if (!aligned16(qkv) || !aligned16(dout) || !aligned16(dqkv)) return STF_E_ALIGN;
  if (ws == 8) {
    cudaStream_t st8 = (cudaStream_t)stream;
    switch (C / heads) {
      case 16: return launch_attn64_bwd<16>(qkv, dout, bias_table, dqkv, dbias_partials, num_windows, C, heads, shift, Hp, Wp, q_scale, st8);
      case 24: return launch_attn64_bwd<24>(qkv, dout, bias_table, dqkv, dbias_partials, num_windows, C, heads, shift, Hp, Wp, q_scale, st8);
      case 32: return launch_attn64_bwd<32>(qkv, dout, bias_table, dqkv, dbias_partials, num_windows, C, heads, shift, Hp, Wp, q_scale, st8);
    }
    return STF_E_SHAPE;
  }
  int wpc = 0;
  const int ctas = stf_attention_bwd_ctas(num_windows, C, heads, &wpc);
  if (ctas < 0) return ctas;
  const int d = C / heads;
  cudaStream_t st = (cudaStream_t)stream;
  switch (d) {
    case 16: return launch_attn_bwd<16>(qkv, dout, bias_table, dqkv, dbias_partials, num_windows, C, heads, shift, Hp, Wp, wpc, q_scale, st);
    case 24: return launch_attn_bwd<24>(qkv, dout, bias_table, dqkv, dbias_partials, num_windows, C, heads, shift, Hp, Wp, wpc, q_scale, st);
    case 32: return launch_attn_bwd<32>(qkv, dout, bias_table, dqkv, dbias_partials, num_windows, C, heads, shift, Hp, Wp, wpc, q_scale, st);
    case 40: return launch_attn_bwd<40>(qkv, dout, bias_table, dqkv, dbias_partials, num_windows, C, heads, shift, Hp, Wp, wpc, q_scale, st);
  }
  return STF_E_SHAPE;
}

extern "C" int stf_layernorm_bwd_ctas(int64_t M) {
  if (M < 0) return STF_E_ARG;
  int64_t b = (M + kLnWarps - 1) / kLnWarps;
  if (b > kNumSMs * 4) b = kNumSMs * 4;
  if (b < 1) b = 1;
  return (int)b;
}

extern "C" int stf_layernorm_bwd(const float *x, const float *g, const float *gamma, const float *beta, const float *res,
                                 float *dx, float *xn, float *partials, int64_t M, int C, float eps, void *stream) {
  if (!x || !g || !gamma || !dx || !partials || M < 0 || C <= 0) return STF_E_ARG;
  if (C > 32 * kLnMaxPerLane) return STF_E_SHAPE;
  if (M == 0) return STF_OK;
  const int ctas = stf_layernorm_bwd_ctas(M);
  const int per = (C + 31) / 32;
#define STF_LNB(P)                                                                                          \
  layernorm_bwd_kernel<P><<<ctas, kLnWarps * 32, (size_t)kLnWarps * 2 * C * 4, (cudaStream_t)stream>>>( \
      x, g, gamma, beta, res, dx, xn, partials, M, C, eps)
  if (per <= 2) STF_LNB(2);
  else if (per <= 3) STF_LNB(3);
  else if (per <= 6) STF_LNB(6);
  else if (per <= 12) STF_LNB(12);
  else STF_LNB(24);
#undef STF_LNB
  return check_launch();
}

extern "C" int stf_colsum_ctas(int64_t M) {
  if (M < 0) return STF_E_ARG;
  int64_t b = (M + 255) / 256;
  if (b > kNumSMs * 4) b = kNumSMs * 4;
  return (int)(b < 1 ? 1 : b);
}

extern "C" int stf_colsum(const float *a, float *partials, int64_t M, int C, void *stream) {
  if (!a || !partials || M < 0 || C <= 0) return STF_E_ARG;
  if (C % 4 != 0 || C > 1024 * 4) return STF_E_SHAPE;
  if (!aligned16(a) || !aligned16(partials)) return STF_E_ALIGN;
  const int ctas = stf_colsum_ctas(M);
  const int c4 = C / 4;
  if (c4 > 256) return STF_E_SHAPE;
  colsum_kernel<<<ctas, 256, (size_t)(256 / c4) * c4 * 16, (cudaStream_t)stream>>>(a, partials, M, C);
  return check_launch();
}

extern "C" int stf_gelu_bwd(const float *pre, const float *dh, float *dpre, int64_t n, void *stream) {
  if (!pre || !dh || !dpre || n < 0) return STF_E_ARG;
  if (n % 4 != 0) return STF_E_SHAPE;
  if (!aligned16(pre) || !aligned16(dh) || !aligned16(dpre)) return STF_E_ALIGN;
  if (n == 0) return STF_OK;
  gelu_bwd_kernel<<<ew_grid(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(pre, dh, dpre, n / 4);
  return check_launch();
}

extern "C" int stf_gaussian_likelihood_train(const float *y, const float *scales, const float *means, const float *noise,
                                             float *likelihood, int64_t n, float scale_bound, float lik_bound,
                                             void *stream) {
  if (!y || !scales || !likelihood || n < 0) return STF_E_ARG;
  if (n == 0) return STF_OK;
  gaussian_train_fwd_kernel<<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(y, scales, means, noise, likelihood, n,
                                                                              scale_bound, lik_bound);
  return check_launch();
}

extern "C" int stf_gaussian_likelihood_train_bwd(const float *y, const float *scales, const float *means,
                                                 const float *noise, const float *dlik, float *dy, float *dscale,
                                                 float *dmean, int64_t n, float scale_bound, float lik_bound,
                                                 void *stream) {
  if (!y || !scales || !dlik || !dy || !dscale || n < 0) return STF_E_ARG;
  if (n == 0) return STF_OK;
  gaussian_train_bwd_kernel<<<ew_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(y, scales, means, noise, dlik, dy, dscale,
                                                                              dmean, n, scale_bound, lik_bound);
  return check_launch();
}
