// Per-window multi-head attention core:  O = softmax(q k^T + B[rel] + mask) v   in fp32.
//
// Replaces reference compressai/models/stf.py:100-118 (same code in layers/win_attention.py:94-112):
// the (B_, nH, N, N) score tensor, the gathered (nH, N, N) bias, the (nW, N, N) mask tensor
// (stf.py:316-334) and the two batched matmuls are never materialised.  q arrives pre-scaled
// from the qkv GEMM epilogue.
//
// The tiles are 16x16x16 (STF) or 64x64x{24,40} (WACNN): far below a tcgen05 M=128 tile and
// only 2-14 % of a block's FLOPs, so this stage runs on the fp32 pipes:
// one thread owns one query row of one (window, head) pair; the 16 (64) threads of a pair read
// the same key / value rows, which the LSU serves as broadcasts out of L1.
//
// Index math contract: token n of window (wy, wx) sits at
// (h', w') = (wy*ws + n/ws, wx*ws + n%ws) of the shifted frame; region label of a coordinate c on
// an axis of padded length L is 0 if c < L-ws, 1 if c < L-shift, else 2; mask = -100 where the
// (row label, column label) pairs of query and key differ; rel_idx(n,m) =
// (hn-hm+ws-1)*(2ws-1) + (wn-wm+ws-1).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "sm100.cuh"

namespace stf {
namespace {

constexpr int kThreads = 128;
constexpr float kMaskValue = -100.0f;  // stf.py:334

// The output feeds the proj GEMM only, which reads TF32: store it already rounded (round-to-nearest on the bit
// pattern) so that stf_linear can take its x_is_tf32 fast path.
__device__ __forceinline__ float round_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

template <int WS, int D>
__global__ void __launch_bounds__(kThreads)
window_attention_kernel(const float *__restrict__ qkv, float *__restrict__ out,
                        const float *__restrict__ bias_table, const float *__restrict__ mask, int mask_windows,
                        int64_t num_pairs, int C, int heads, int shift, int Hp, int Wp, int tf32_out) {
  constexpr int N = WS * WS;
  constexpr int PAIRS = kThreads / N;
  const int n = threadIdx.x % N;
  const int64_t pair = (int64_t)blockIdx.x * PAIRS + threadIdx.x / N;
  if (pair >= num_pairs) return;
  const int64_t win = pair / heads;
  const int head = (int)(pair - win * heads);
  const int ld = 3 * C;
  const float *base = qkv + win * N * (int64_t)ld + head * D;

  float q[D];
#pragma unroll
  for (int j = 0; j < D; j += 4) {
    float4 v = __ldg(reinterpret_cast<const float4 *>(base + (int64_t)n * ld + j));
    q[j] = v.x, q[j + 1] = v.y, q[j + 2] = v.z, q[j + 3] = v.w;
  }

  // region labels (only windows in the last window row / column see a non-trivial mask)
  const int hn = n / WS, wn = n % WS;
  int my_label = 0, wy = 0, wx = 0;
  if (shift > 0) {
    const int nWw = Wp / WS, nW = (Hp / WS) * nWw;
    const int wi = (int)(win % nW);
    wy = wi / nWw, wx = wi - wy * nWw;
    const int hs = wy * WS + hn, wsft = wx * WS + wn;
    my_label = 3 * (hs < Hp - WS ? 0 : (hs < Hp - shift ? 1 : 2)) + (wsft < Wp - WS ? 0 : (wsft < Wp - shift ? 1 : 2));
  }

  float s[N];
  float smax = -INFINITY;
  const float *kbase = base + C;
#pragma unroll
  for (int m = 0; m < N; ++m) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < D; j += 4) {
      float4 kv = __ldg(reinterpret_cast<const float4 *>(kbase + (int64_t)m * ld + j));
      acc = fmaf(q[j], kv.x, acc);
      acc = fmaf(q[j + 1], kv.y, acc);
      acc = fmaf(q[j + 2], kv.z, acc);
      acc = fmaf(q[j + 3], kv.w, acc);
    }
    const int hm = m / WS, wm = m % WS;
    const int rel = (hn - hm + WS - 1) * (2 * WS - 1) + (wn - wm + WS - 1);
    acc += __ldg(bias_table + rel * heads + head);
    if (shift > 0) {
      const int hs = wy * WS + hm, wsft = wx * WS + wm;
      const int lab = 3 * (hs < Hp - WS ? 0 : (hs < Hp - shift ? 1 : 2)) + (wsft < Wp - WS ? 0 : (wsft < Wp - shift ? 1 : 2));
      if (lab != my_label) acc += kMaskValue;
    }
    if (mask) acc += __ldg(mask + ((int64_t)(win % mask_windows) * N + n) * N + m);  // explicit mask (stf.py:108-110)
    s[m] = acc;
    smax = fmaxf(smax, acc);
  }
  float denom = 0.f;
#pragma unroll
  for (int m = 0; m < N; ++m) {
    s[m] = expf(s[m] - smax);
    denom += s[m];
  }
  const float inv = 1.0f / denom;

  float o[D];
#pragma unroll
  for (int j = 0; j < D; ++j) o[j] = 0.f;
  const float *vbase = base + 2 * C;
#pragma unroll
  for (int m = 0; m < N; ++m) {
    const float p = s[m] * inv;
#pragma unroll
    for (int j = 0; j < D; j += 4) {
      float4 vv = __ldg(reinterpret_cast<const float4 *>(vbase + (int64_t)m * ld + j));
      o[j] = fmaf(p, vv.x, o[j]);
      o[j + 1] = fmaf(p, vv.y, o[j + 1]);
      o[j + 2] = fmaf(p, vv.z, o[j + 2]);
      o[j + 3] = fmaf(p, vv.w, o[j + 3]);
    }
  }
  float *dst = out + (win * N + n) * (int64_t)C + head * D;
#pragma unroll
  for (int j = 0; j < D; j += 4)
    *reinterpret_cast<float4 *>(dst + j) =
        tf32_out ? make_float4(round_tf32(o[j]), round_tf32(o[j + 1]), round_tf32(o[j + 2]), round_tf32(o[j + 3]))
                     : make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
}


// ---------------------------------------------------------------------------------------------
// 16-token windows (every STF stage, WACNN's d=40 block): shared-memory tiled variant.
// A CTA owns `wpc` consecutive windows.  Their qkv rows are one contiguous block of wpc*16*3C floats,
// fetched with a single bulk TMA copy; every thread (window, head, query row) then works out of shared
// memory (keys / values are broadcast reads), writes its output over its own q slot, and the rows leave
// with one bulk TMA store each.  HBM sees exactly one coalesced read of qkv and one write of the output.
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(384)
window_attention16_kernel(const float *__restrict__ qkv, float *__restrict__ out,
                          const float *__restrict__ bias_table, const float *__restrict__ mask, int mask_windows,
                          int64_t num_windows, int C, int heads, int shift, int Hp, int Wp, int wpc, int tf32_out) {
  constexpr int WS = 4, N = 16;
  extern __shared__ __align__(128) float tile[];  // [wpc][16][3C]
  __shared__ __align__(8) uint64_t bar;
  const int ld = 3 * C;
  const int64_t win0 = (int64_t)blockIdx.x * wpc;
  const int nwin = (int)((num_windows - win0) < wpc ? (num_windows - win0) : wpc);
  const uint32_t bytes = (uint32_t)(nwin * N * ld * 4);
  if (threadIdx.x == 0) {
    sm100::mbar_init(&bar, 1);
    sm100::mbar_fence_init();
    sm100::mbar_arrive_expect_tx(&bar, bytes);
    sm100::bulk_copy_g2s(tile, qkv + win0 * N * (int64_t)ld, bytes, &bar);
  }
  __syncthreads();  // barrier initialised before anyone waits on it
  sm100::mbar_wait(&bar, 0);

  const int n = threadIdx.x % N;
  const int pair = threadIdx.x / N;  // (local window, head)
  const int wl = pair / heads, head = pair - wl * heads;
  const bool active = wl < nwin;
  if (active) {
    const int64_t win = win0 + wl;
    float *base = tile + (size_t)wl * N * ld + head * D;
    float q[D];
#pragma unroll
    for (int j = 0; j < D; j += 4) {
      float4 v = *reinterpret_cast<const float4 *>(base + n * ld + j);
      q[j] = v.x, q[j + 1] = v.y, q[j + 2] = v.z, q[j + 3] = v.w;
    }
    const int hn = n / WS, wn = n % WS;
    int my_label = 0, wy = 0, wx = 0;
    if (shift > 0) {
      const int nWw = Wp / WS, nW = (Hp / WS) * nWw;
      const int wi = (int)(win % nW);
      wy = wi / nWw, wx = wi - wy * nWw;
      const int hs = wy * WS + hn, wsft = wx * WS + wn;
      my_label = 3 * (hs < Hp - WS ? 0 : (hs < Hp - shift ? 1 : 2)) + (wsft < Wp - WS ? 0 : (wsft < Wp - shift ? 1 : 2));
    }
    float s[N];
    float smax = -INFINITY;
    const float *kbase = base + C;
#pragma unroll
    for (int m = 0; m < N; ++m) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < D; j += 4) {
        const float4 kv = *reinterpret_cast<const float4 *>(kbase + m * ld + j);
        acc = fmaf(q[j], kv.x, acc);
        acc = fmaf(q[j + 1], kv.y, acc);
        acc = fmaf(q[j + 2], kv.z, acc);
        acc = fmaf(q[j + 3], kv.w, acc);
      }
      const int hm = m / WS, wm = m % WS;
      const int rel = (hn - hm + WS - 1) * (2 * WS - 1) + (wn - wm + WS - 1);
      acc += __ldg(bias_table + rel * heads + head);
      if (shift > 0) {
        const int hs = wy * WS + hm, wsft = wx * WS + wm;
        const int lab = 3 * (hs < Hp - WS ? 0 : (hs < Hp - shift ? 1 : 2)) + (wsft < Wp - WS ? 0 : (wsft < Wp - shift ? 1 : 2));
        if (lab != my_label) acc += kMaskValue;
      }
      if (mask) acc += __ldg(mask + ((int64_t)(win % mask_windows) * N + n) * N + m);
      s[m] = acc;
      smax = fmaxf(smax, acc);
    }
    float denom = 0.f;
#pragma unroll
    for (int m = 0; m < N; ++m) {
      s[m] = expf(s[m] - smax);
      denom += s[m];
    }
    const float inv = 1.0f / denom;
    float o[D];
#pragma unroll
    for (int j = 0; j < D; ++j) o[j] = 0.f;
    const float *vbase = base + 2 * C;
#pragma unroll
    for (int m = 0; m < N; ++m) {
      const float p = s[m] * inv;
#pragma unroll
      for (int j = 0; j < D; j += 4) {
        const float4 vv = *reinterpret_cast<const float4 *>(vbase + m * ld + j);
        o[j] = fmaf(p, vv.x, o[j]);
        o[j + 1] = fmaf(p, vv.y, o[j + 1]);
        o[j + 2] = fmaf(p, vv.z, o[j + 2]);
        o[j + 3] = fmaf(p, vv.w, o[j + 3]);
      }
    }
    // the q slot (row n, this head's columns) is read by this thread only: reuse it for the output
#pragma unroll
    for (int j = 0; j < D; j += 4)
      *reinterpret_cast<float4 *>(base + n * ld + j) =
          tf32_out ? make_float4(round_tf32(o[j]), round_tf32(o[j + 1]), round_tf32(o[j + 2]), round_tf32(o[j + 3]))
                     : make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
  }
  sm100::fence_proxy_async_smem();
  __syncthreads();
  // one bulk store per token row: first C floats of the staged row -> out row
  if ((int)threadIdx.x < nwin * N) {
    const int r = threadIdx.x;
    const uint32_t src = sm100::smem_u32(tile + (size_t)r * ld);
    float *dst = out + (win0 * N + r) * (int64_t)C;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"((uint32_t)(C * 4))
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// 64-token windows (WACNN's first attention block: 8x8 windows, 8 heads of 24 channels): one CTA per window.
// The window's 64 qkv rows are one contiguous block (64 * 3C floats = 147 KB at C = 192) fetched with a single bulk
// TMA copy; thread = (head, query row) works out of shared memory with an ONLINE softmax over chunks of 8 keys
// (running maximum / denominator, accumulator rescaled once per chunk) so that no 64-entry score array lives in
// registers; outputs overwrite the thread's own q slot and leave with one bulk store per token row.
// (The global-memory variant above ran at 7 % of the HBM roofline on the 2048x1408 WACNN configuration.)
// ---------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(512)
window_attention64_kernel(const float *__restrict__ qkv, float *__restrict__ out,
                          const float *__restrict__ bias_table, const float *__restrict__ mask, int mask_windows,
                          int64_t num_windows, int C, int heads, int shift, int Hp, int Wp, int tf32_out) {
  constexpr int WS = 8, N = 64, R = 2 * WS - 1;
  extern __shared__ __align__(128) float tile[];  // [64][3C] then the bias table [R*R][heads]
  __shared__ __align__(8) uint64_t bar;
  const int ld = 3 * C;
  const int64_t win = blockIdx.x;
  float *tbl = tile + (size_t)N * ld;
  const uint32_t bytes = (uint32_t)(N * ld * 4);
  if (threadIdx.x == 0) {
    sm100::mbar_init(&bar, 1);
    sm100::mbar_fence_init();
    sm100::mbar_arrive_expect_tx(&bar, bytes);
    sm100::bulk_copy_g2s(tile, qkv + win * N * (int64_t)ld, bytes, &bar);
  }
  for (int i = threadIdx.x; i < R * R * heads; i += blockDim.x) tbl[i] = __ldg(bias_table + i);
  __syncthreads();  // barrier initialised + bias table staged
  sm100::mbar_wait(&bar, 0);

  const int n = threadIdx.x % N;
  const int head = threadIdx.x / N;
  if (head < heads) {
    float *base = tile + head * D;
    float q[D], o[D];
#pragma unroll
    for (int j = 0; j < D; j += 4) {
      const float4 v = *reinterpret_cast<const float4 *>(base + n * ld + j);
      q[j] = v.x, q[j + 1] = v.y, q[j + 2] = v.z, q[j + 3] = v.w;
    }
#pragma unroll
    for (int j = 0; j < D; ++j) o[j] = 0.f;
    const int hn = n / WS, wn = n % WS;
    int my_label = 0, wy = 0, wx = 0;
    if (shift > 0) {
      const int nWw = Wp / WS, nW = (Hp / WS) * nWw;
      const int wi = (int)(win % nW);
      wy = wi / nWw, wx = wi - wy * nWw;
      const int hs = wy * WS + hn, wsft = wx * WS + wn;
      my_label = 3 * (hs < Hp - WS ? 0 : (hs < Hp - shift ? 1 : 2)) + (wsft < Wp - WS ? 0 : (wsft < Wp - shift ? 1 : 2));
    }
    const float *kbase = base + C, *vbase = base + 2 * C;
    float run_max = -INFINITY, denom = 0.f;
    for (int m0 = 0; m0 < N; m0 += 8) {
      float s[8];
      float cmax = -INFINITY;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int m = m0 + u;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < D; j += 4) {
          const float4 kv = *reinterpret_cast<const float4 *>(kbase + m * ld + j);
          acc = fmaf(q[j], kv.x, acc);
          acc = fmaf(q[j + 1], kv.y, acc);
          acc = fmaf(q[j + 2], kv.z, acc);
          acc = fmaf(q[j + 3], kv.w, acc);
        }
        const int hm = m / WS, wm = m % WS;
        acc += tbl[((hn - hm + WS - 1) * R + (wn - wm + WS - 1)) * heads + head];
        if (shift > 0) {
          const int hs = wy * WS + hm, wsft = wx * WS + wm;
          const int lab = 3 * (hs < Hp - WS ? 0 : (hs < Hp - shift ? 1 : 2)) + (wsft < Wp - WS ? 0 : (wsft < Wp - shift ? 1 : 2));
          if (lab != my_label) acc += kMaskValue;
        }
        if (mask) acc += __ldg(mask + ((int64_t)(win % mask_windows) * N + n) * N + m);
        s[u] = acc;
        cmax = fmaxf(cmax, acc);
      }
      const float new_max = fmaxf(run_max, cmax);
      const float scale = expf(run_max - new_max);   // exp(-inf) = 0 on the first chunk
      denom *= scale;
#pragma unroll
      for (int j = 0; j < D; ++j) o[j] *= scale;
      run_max = new_max;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float p = expf(s[u] - run_max);
        denom += p;
        const int m = m0 + u;
#pragma unroll
        for (int j = 0; j < D; j += 4) {
          const float4 vv = *reinterpret_cast<const float4 *>(vbase + m * ld + j);
          o[j] = fmaf(p, vv.x, o[j]);
          o[j + 1] = fmaf(p, vv.y, o[j + 1]);
          o[j + 2] = fmaf(p, vv.z, o[j + 2]);
          o[j + 3] = fmaf(p, vv.w, o[j + 3]);
        }
      }
    }
    const float inv = 1.0f / denom;
    // the q slot (row n, this head's columns) is read by this thread only: reuse it for the output
#pragma unroll
    for (int j = 0; j < D; j += 4) {
      const float4 r = make_float4(o[j] * inv, o[j + 1] * inv, o[j + 2] * inv, o[j + 3] * inv);
      *reinterpret_cast<float4 *>(base + n * ld + j) =
          tf32_out ? make_float4(round_tf32(r.x), round_tf32(r.y), round_tf32(r.z), round_tf32(r.w)) : r;
    }
  }
  sm100::fence_proxy_async_smem();
  __syncthreads();
  if ((int)threadIdx.x < N) {  // one bulk store per token row: first C floats of the staged row -> out row
    const int r = threadIdx.x;
    const uint32_t src = sm100::smem_u32(tile + (size_t)r * ld);
    float *dst = out + (win * N + r) * (int64_t)C;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"((uint32_t)(C * 4))
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

template <int D>
int launch64(const float *qkv, float *out, const float *bias_table, const float *mask, int mask_windows,
             int64_t num_windows, int C, int heads, int shift, int Hp, int Wp, int tf32_out, cudaStream_t st) {
  const size_t smem = (size_t)64 * 3 * C * 4 + (size_t)15 * 15 * heads * 4;
  if (heads * 64 > 512 || smem > 220 * 1024) return -100;   // not this variant: the caller falls back to the generic kernel
  static std::atomic<int> attr_set{0};
  if (!attr_set.load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(window_attention64_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set.store(1, std::memory_order_release);
  }
  if (num_windows > 0x7fffffffLL) return STF_E_SHAPE;
  window_attention64_kernel<D><<<(unsigned)num_windows, heads * 64, smem, st>>>(qkv, out, bias_table, mask, mask_windows,
                                                                               num_windows, C, heads, shift, Hp, Wp, tf32_out);
  return check_launch();
}

template <int D>
int launch16(const float *qkv, float *out, const float *bias_table, const float *mask, int mask_windows,
             int64_t num_windows, int C, int heads, int shift, int Hp, int Wp, int tf32_out, cudaStream_t st) {
  const int per_window = heads * 16;                 // threads per window
  const size_t win_bytes = (size_t)16 * 3 * C * 4;   // qkv tile of one window
  int wpc = 384 / per_window;
  static const int tile_kb = [] {
    const char *e = getenv("STF_B200_ATTN_TILE_KB");
    return e ? atoi(e) : 24;
  }();
  // 24 KB tiles (2-3 windows at C = 48): many small CTAs per SM overlap each other's load / compute / store phases better
  // than three 72 KB ones (measured at stage 0: 2.78 -> 3.24 TB/s)
  const int by_smem = (int)(((size_t)tile_kb * 1024) / win_bytes);
  if (wpc > by_smem) wpc = by_smem;
  if (wpc < 1) wpc = 1;
  if (per_window > 384 || win_bytes > 200 * 1024) return STF_E_SHAPE;
  const size_t smem = (size_t)wpc * win_bytes;
  static std::atomic<int> attr_set{0};
  if (!attr_set.load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(window_attention16_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set.store(1, std::memory_order_release);
  }
  const int64_t blocks = (num_windows + wpc - 1) / wpc;
  if (blocks > 0x7fffffffLL) return STF_E_SHAPE;
  const int threads = (wpc * per_window + 31) / 32 * 32;
  window_attention16_kernel<D><<<(unsigned)blocks, threads, smem, st>>>(qkv, out, bias_table, mask, mask_windows,
                                                                      num_windows, C, heads, shift, Hp, Wp, wpc, tf32_out);
  return check_launch();
}

template <int WS, int D>
int launch(const float *qkv, float *out, const float *bias_table, const float *mask, int mask_windows,
           int64_t num_windows, int C, int heads, int shift, int Hp, int Wp, int tf32_out, cudaStream_t st) {
  constexpr int PAIRS = kThreads / (WS * WS);
  const int64_t pairs = num_windows * heads;
  const int64_t blocks = (pairs + PAIRS - 1) / PAIRS;
  if (blocks > 0x7fffffffLL) return STF_E_SHAPE;
  window_attention_kernel<WS, D><<<(unsigned)blocks, kThreads, 0, st>>>(qkv, out, bias_table, mask, mask_windows, pairs, C, heads, shift, Hp, Wp, tf32_out);
  return check_launch();
}

}  // namespace
}  // namespace stf

using namespace stf;

extern "C" int stf_window_attention(const float *qkv, float *out, const float *bias_table, const float *mask,
                                    int mask_windows, int64_t num_windows, int C, int heads, int ws, int shift,
                                    int Hp, int Wp, int tf32_out, void *stream) {
  if (!qkv || !out || !bias_table || num_windows < 0 || C <= 0 || heads <= 0) return STF_E_ARG;
  if (num_windows == 0) return STF_OK;
  if (C % heads != 0 || shift < 0 || shift >= ws) return STF_E_SHAPE;
  if (shift > 0 && (Hp <= 0 || Wp <= 0 || Hp % ws != 0 || Wp % ws != 0)) return STF_E_SHAPE;
  if (mask && (mask_windows <= 0 || num_windows % mask_windows != 0)) return STF_E_SHAPE;
  if (!aligned16(qkv) || !aligned16(out)) return STF_E_ALIGN;
  const int d = C / heads;
  cudaStream_t st = (cudaStream_t)stream;
#define CASE(WS_, D_) \
  if (ws == WS_ && d == D_) return launch<WS_, D_>(qkv, out, bias_table, mask, mask_windows, num_windows, C, heads, shift, Hp, Wp, tf32_out, st)
#define CASE16(D_) \
  if (ws == 4 && d == D_) return launch16<D_>(qkv, out, bias_table, mask, mask_windows, num_windows, C, heads, shift, Hp, Wp, tf32_out, st)
  CASE16(16);
  CASE16(24);
  CASE16(32);
  CASE16(40);
#undef CASE16
#define CASE64(D_)                                                                                               \
  if (ws == 8 && d == D_) {                                                                                      \
    const int rc = launch64<D_>(qkv, out, bias_table, mask, mask_windows, num_windows, C, heads, shift, Hp, Wp, tf32_out, st); \
    if (rc != -100) return rc;                                                                                   \
  }
  CASE64(16);
  CASE64(24);
  CASE64(32);
  CASE64(40);
#undef CASE64
  CASE(8, 16);
  CASE(8, 24);
  CASE(8, 32);
  CASE(8, 40);
#undef CASE
  return STF_E_SHAPE;
}

// =============================================================================================
// Token-order window attention for 4x4 windows, head_dim 16 (every STF stage), contractions on tensor cores.
//
// qkv arrives in TOKEN order (B, H, W, 3C) straight out of the dense qkv GEMM (stf_conv2d, ksize 1, LayerNorm folded, q
// pre-scaled): window partition, cyclic shift, zero padding, window reverse and un-shift (stf.py:155-196) are this
// kernel's address arithmetic on per-token bulk (TMA) copies -- token (py, px) of window (wy, wx) is fetched from
// ((wy*4 + py + shift) mod Hp, (wx*4 + px + shift) mod Wp) and its output row goes back to the same place; pad tokens
// (h >= H or w >= W: zero AFTER norm1, stf.py:155-162) take the qkv bias row and their outputs are dropped.
// One warp per (window, head): S = q k^T and O = P v as mma.sync.m16n8k8 TF32 tensor-core tiles (hi / lo operand split =
// 3xTF32 in the fp32 precision mode), relative-position bias + analytic shifted-window mask added on the accumulator
// fragments, softmax with two shuffles per row reduction; P goes from the score fragments to the PV operand fragments in
// registers (keys relabelled 2t <-> t, 2t+1 <-> t+4 on both operands, no data movement).
// Shared-memory rows are padded to 3C + 4 floats: every fragment load is bank-conflict free.
// =============================================================================================
namespace stf {
namespace {

__device__ __forceinline__ uint32_t cvt_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// D += A . B with fp32-grade operands: hi.hi + lo.hi + hi.lo (the lo.lo term is below fp32 round-off); single pass otherwise
template <int kPrecise>
__device__ __forceinline__ void mma_split(float (&d)[4], const float (&a)[4], const float (&b)[2]) {
  uint32_t ah[4], bh[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) ah[i] = cvt_tf32(a[i]);
#pragma unroll
  for (int i = 0; i < 2; ++i) bh[i] = cvt_tf32(b[i]);
  mma_tf32(d, ah, bh);
  if (kPrecise) {
    uint32_t al[4], bl[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) al[i] = cvt_tf32(a[i] - __uint_as_float(ah[i]));
#pragma unroll
    for (int i = 0; i < 2; ++i) bl[i] = cvt_tf32(b[i] - __uint_as_float(bh[i]));
    mma_tf32(d, al, bh);
    mma_tf32(d, ah, bl);
  }
}

struct AttnTokParams {
  const float *qkv;       // (B, H, W, 3C)
  float *out;             // (B, H, W, C)
  const float *table;     // (49, heads)
  const float *pad_qkv;   // 3C: qkv row of a pad token (bias, q part scaled); may be null when there are no pad tokens
  int B, H, W, Hp, Wp, nWw, nW, C, heads, shift, wpc;
  int64_t num_windows;
};

template <int kPrecise>
__global__ void __launch_bounds__(192, 4) window_attention_tok_kernel(const AttnTokParams P) {
  constexpr int WS = 4, N = 16, D = 16;
  extern __shared__ __align__(128) float smem_f[];
  __shared__ __align__(8) uint64_t bar;
  const int C = P.C, heads = P.heads;
  const int ld = 3 * C + 4;                        // padded row pitch (floats)
  float *tile = smem_f;                            // [wpc * 16][ld]
  float *ostage = tile + (size_t)P.wpc * N * ld;   // [wpc * 16][C]
  float *tbl = ostage + (size_t)P.wpc * N * C;     // [49 * heads]
  const int64_t win0 = (int64_t)blockIdx.x * P.wpc;
  const int nwin = (int)((P.num_windows - win0) < P.wpc ? (P.num_windows - win0) : P.wpc);
  const int ntok = nwin * N;
  if (threadIdx.x == 0) {
    sm100::mbar_init(&bar, (uint32_t)ntok);
    sm100::mbar_fence_init();
  }
  for (int i = threadIdx.x; i < 49 * heads; i += blockDim.x) tbl[i] = __ldg(P.table + i);
  __syncthreads();

  // ---- gather: one thread per token issues that token's bulk copy (or fills a pad token with the bias row)
  int64_t my_tok = -1;   // global token index of this thread's token (valid tokens only)
  if ((int)threadIdx.x < ntok) {
    const int t = threadIdx.x, wl = t >> 4, n = t & 15;
    const int64_t win = win0 + wl;
    const int b = (int)(win / P.nW), wi = (int)(win - (int64_t)b * P.nW);
    const int wy = wi / P.nWw, wx = wi - wy * P.nWw;
    int y = wy * WS + (n >> 2) + P.shift, x = wx * WS + (n & 3) + P.shift;   // torch.roll(x, -shift): shifted[h'] = x[(h'+s) mod Hp]
    if (y >= P.Hp) y -= P.Hp;
    if (x >= P.Wp) x -= P.Wp;
    float *dst = tile + (size_t)t * ld;
    if (y < P.H && x < P.W) {
      my_tok = ((int64_t)b * P.H + y) * P.W + x;
      sm100::mbar_arrive_expect_tx(&bar, (uint32_t)(3 * C * 4));
      sm100::bulk_copy_g2s(dst, P.qkv + my_tok * 3 * C, (uint32_t)(3 * C * 4), &bar);
    } else {
      for (int j = 0; j < 3 * C; j += 4) *reinterpret_cast<float4 *>(dst + j) = __ldg(reinterpret_cast<const float4 *>(P.pad_qkv + j));
      sm100::mbar_arrive(&bar);
    }
  }
  sm100::mbar_wait(&bar, 0);
  __syncthreads();   // pad rows written with generic stores are visible to every warp

  // ---- one warp per (window, head)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = lane >> 2, tig = lane & 3;
  for (int pair = warp; pair < nwin * heads; pair += (int)(blockDim.x >> 5)) {
    const int wl = pair / heads, head = pair - wl * heads;
    const int64_t win = win0 + wl;
    const float *qb = tile + (size_t)wl * N * ld + head * D, *kb = qb + C, *vb = qb + 2 * C;
    // S = q k^T : rows gid / gid+8, key tiles nt = 0, 1
    float s[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) s[nt][i] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const float a[4] = {qb[gid * ld + 8 * ks + tig], qb[(gid + 8) * ld + 8 * ks + tig], qb[gid * ld + 8 * ks + tig + 4],
                          qb[(gid + 8) * ld + 8 * ks + tig + 4]};
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const float bb[2] = {kb[(8 * nt + gid) * ld + 8 * ks + tig], kb[(8 * nt + gid) * ld + 8 * ks + tig + 4]};
        mma_split<kPrecise>(s[nt], a, bb);
      }
    }
    // + relative-position bias + shifted-window mask (stf.py:100-110, 316-334); thread holds rows r0 = gid, r1 = gid + 8 and
    // keys m = 8 nt + 2 tig + {0, 1}
    int wy = 0, wx = 0;
    bool edge = false;   // only the last window row / column of the shifted frame mixes regions (every other label is 0)
    if (P.shift > 0) {
      const int wi = (int)(win % P.nW);
      wy = wi / P.nWw, wx = wi - wy * P.nWw;
      edge = wy == P.Hp / WS - 1 || wx == P.nWw - 1;
    }
    auto label = [&](int n) -> int {
      const int hs = wy * WS + (n >> 2), wsft = wx * WS + (n & 3);
      return 3 * (hs < P.Hp - WS ? 0 : (hs < P.Hp - P.shift ? 1 : 2)) + (wsft < P.Wp - WS ? 0 : (wsft < P.Wp - P.shift ? 1 : 2));
    };
    const int lab_r0 = edge ? label(gid) : 0, lab_r1 = edge ? label(gid + 8) : 0;
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int m = 8 * nt + 2 * tig + e;
        const int hm = m >> 2, wm = m & 3;
        const int rel0 = ((gid >> 2) - hm + WS - 1) * (2 * WS - 1) + ((gid & 3) - wm + WS - 1);
        const int rel1 = (((gid + 8) >> 2) - hm + WS - 1) * (2 * WS - 1) + (((gid + 8) & 3) - wm + WS - 1);
        float v0 = s[nt][e] + tbl[rel0 * heads + head], v1 = s[nt][2 + e] + tbl[rel1 * heads + head];
        if (edge) {
          const int lm = label(m);
          if (lm != lab_r0) v0 += kMaskValue;
          if (lm != lab_r1) v1 += kMaskValue;
        }
        s[nt][e] = v0, s[nt][2 + e] = v1;
        mx0 = fmaxf(mx0, v0), mx1 = fmaxf(mx1, v1);
      }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[nt][e] = expf(s[nt][e] - mx0), s[nt][2 + e] = expf(s[nt][2 + e] - mx1);
        sum0 += s[nt][e], sum1 += s[nt][2 + e];
      }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
    // O = P v : the score fragment (rows gid / gid+8, keys 2 tig, 2 tig + 1 of key tile kt) IS the A fragment of k-step kt
    // when MMA k index t stands for key 2t and t + 4 for key 2t + 1; the v fragment uses the same relabelling.
    float o[2][4];
#pragma unroll
    for (int nd = 0; nd < 2; ++nd)
#pragma unroll
      for (int i = 0; i < 4; ++i) o[nd][i] = 0.f;
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) {
      const float a[4] = {s[kt][0] * inv0, s[kt][2] * inv1, s[kt][1] * inv0, s[kt][3] * inv1};
#pragma unroll
      for (int nd = 0; nd < 2; ++nd) {
        const float bb[2] = {vb[(8 * kt + 2 * tig) * ld + 8 * nd + gid], vb[(8 * kt + 2 * tig + 1) * ld + 8 * nd + gid]};
        mma_split<kPrecise>(o[nd], a, bb);
      }
    }
    float *ob = ostage + (size_t)wl * N * C + head * D;
#pragma unroll
    for (int nd = 0; nd < 2; ++nd) {
      *reinterpret_cast<float2 *>(ob + gid * C + 8 * nd + 2 * tig) = make_float2(o[nd][0], o[nd][1]);
      *reinterpret_cast<float2 *>(ob + (gid + 8) * C + 8 * nd + 2 * tig) = make_float2(o[nd][2], o[nd][3]);
    }
  }
  sm100::fence_proxy_async_smem();
  __syncthreads();
  // ---- scatter: window_reverse + un-shift + crop = each valid token's row goes back to where its qkv row came from
  if (my_tok >= 0) {
    const uint32_t src = sm100::smem_u32(ostage + (size_t)threadIdx.x * C);
    float *dst = P.out + my_tok * C;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"((uint32_t)(C * 4))
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    // the staging rows must outlive the copies' READS only: the CTA leaves (and the next one starts its gather) while the
    // rows are still on their way to HBM; the writes are complete when the grid is
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

}  // namespace
}  // namespace stf

extern "C" int stf_window_attention_tokens(const float *qkv, float *out, const float *bias_table, const float *pad_qkv,
                                           int batch, int H, int W, int C, int heads, int ws, int shift, int precision,
                                           void *stream) {
  using namespace stf;
  if (!qkv || !out || !bias_table || batch < 0 || H <= 0 || W <= 0 || C <= 0 || heads <= 0) return STF_E_ARG;
  if (ws != 4 || C % heads != 0 || C / heads != 16 || shift < 0 || shift >= ws || C % 4) return STF_E_SHAPE;
  if (!aligned16(qkv) || !aligned16(out)) return STF_E_ALIGN;
  AttnTokParams P{};
  P.qkv = qkv, P.out = out, P.table = bias_table, P.pad_qkv = pad_qkv;
  P.B = batch, P.H = H, P.W = W, P.Hp = (H + 3) / 4 * 4, P.Wp = (W + 3) / 4 * 4;
  if ((P.Hp != H || P.Wp != W) && (!pad_qkv || !aligned16(pad_qkv))) return STF_E_ARG;
  P.nWw = P.Wp / 4, P.nW = (P.Hp / 4) * P.nWw, P.C = C, P.heads = heads, P.shift = shift;
  P.num_windows = (int64_t)batch * P.nW;
  if (P.num_windows == 0) return STF_OK;
  // Small CTAs (6 warps, 4 resident per SM): a CTA's gather latency is covered by the other three's compute.  4 / 2 / 1
  // windows per CTA at 3 / 6 / >= 12 heads, i.e. 12+ (window, head) pairs for its 6 warps.
  int wpc = 12 / heads;
  if (wpc < 1) wpc = 1;
  P.wpc = wpc;
  const size_t smem = ((size_t)wpc * 16 * (3 * C + 4) + (size_t)wpc * 16 * C + (size_t)49 * heads) * sizeof(float);
  if (smem > 110 * 1024) return STF_E_SHAPE;   // (C = 384: 93 KB for its single window)
  const int64_t blocks = (P.num_windows + wpc - 1) / wpc;
  if (blocks > 0x7fffffffLL) return STF_E_SHAPE;
  int threads = wpc * heads * 32;
  if (threads > 192) threads = 192;
  if (threads < wpc * 16) threads = (wpc * 16 + 31) / 32 * 32;
  const int precise = precision == STF_PREC_FP32;
  auto kern = precise ? window_attention_tok_kernel<1> : window_attention_tok_kernel<0>;
  static std::atomic<int> attr_set[2];
  if (!attr_set[precise].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set[precise].store(1, std::memory_order_release);
  }
  kern<<<(unsigned)blocks, threads, smem, (cudaStream_t)stream>>>(P);
  return check_launch();
}
