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
