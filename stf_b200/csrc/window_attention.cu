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

#include "common.cuh"

namespace stf {
namespace {

constexpr int kThreads = 128;
constexpr float kMaskValue = -100.0f;  // stf.py:334

template <int WS, int D>
__global__ void __launch_bounds__(kThreads)
window_attention_kernel(const float *__restrict__ qkv, float *__restrict__ out,
                        const float *__restrict__ bias_table, const float *__restrict__ mask, int mask_windows,
                        int64_t num_pairs, int C, int heads, int shift, int Hp, int Wp) {
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
  for (int j = 0; j < D; j += 4) *reinterpret_cast<float4 *>(dst + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
}

template <int WS, int D>
int launch(const float *qkv, float *out, const float *bias_table, const float *mask, int mask_windows,
           int64_t num_windows, int C, int heads, int shift, int Hp, int Wp, cudaStream_t st) {
  constexpr int PAIRS = kThreads / (WS * WS);
  const int64_t pairs = num_windows * heads;
  const int64_t blocks = (pairs + PAIRS - 1) / PAIRS;
  if (blocks > 0x7fffffffLL) return STF_E_SHAPE;
  window_attention_kernel<WS, D><<<(unsigned)blocks, kThreads, 0, st>>>(qkv, out, bias_table, mask, mask_windows, pairs, C, heads, shift, Hp, Wp);
  return check_launch();
}

}  // namespace
}  // namespace stf

using namespace stf;

extern "C" int stf_window_attention(const float *qkv, float *out, const float *bias_table, const float *mask,
                                    int mask_windows, int64_t num_windows, int C, int heads, int ws, int shift,
                                    int Hp, int Wp, void *stream) {
  if (!qkv || !out || !bias_table || num_windows < 0 || C <= 0 || heads <= 0) return STF_E_ARG;
  if (num_windows == 0) return STF_OK;
  if (C % heads != 0 || shift < 0 || shift >= ws) return STF_E_SHAPE;
  if (shift > 0 && (Hp <= 0 || Wp <= 0 || Hp % ws != 0 || Wp % ws != 0)) return STF_E_SHAPE;
  if (mask && (mask_windows <= 0 || num_windows % mask_windows != 0)) return STF_E_SHAPE;
  if (!aligned16(qkv) || !aligned16(out)) return STF_E_ALIGN;
  const int d = C / heads;
  cudaStream_t st = (cudaStream_t)stream;
#define CASE(WS_, D_) \
  if (ws == WS_ && d == D_) return launch<WS_, D_>(qkv, out, bias_table, mask, mask_windows, num_windows, C, heads, shift, Hp, Wp, st)
  CASE(4, 16);
  CASE(4, 24);
  CASE(4, 32);
  CASE(4, 40);
  CASE(8, 16);
  CASE(8, 24);
  CASE(8, 32);
  CASE(8, 40);
#undef CASE
  return STF_E_SHAPE;
}
