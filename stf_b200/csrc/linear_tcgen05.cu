// Fused linear layer on 5th-gen tensor cores:  Y = epilogue( prologue(X) . W^T ),  TF32 x TF32 -> FP32.
//
// Replaces (reference compressai/models/stf.py):
//   norm1 + F.pad + torch.roll + window_partition + qkv Linear + q*scale     :155-175, 97-100
//   proj Linear + window_reverse + torch.roll + crop + shortcut add            :119, 181-196
//   norm2 + fc1 + exact GELU                                                   :197, 35-36
//   fc2 + residual add                                                         :38, 197
//   PatchMerging (2x2 gather, LayerNorm(4C), Linear 4C->2C)                    :209-235
//   PatchSplit (LayerNorm, Linear C->2C, PixelShuffle(2) in token layout)      :251-260
//
// CTA = one 128-row M tile x one n_tile-column N tile.  6 warps, warp-specialised:
//   warps 0-3  A producers: gather rows (index math), LayerNorm in fp32, round to TF32, store into
//              the UMMA canonical K-major layout in shared memory; afterwards the same warps run
//              the epilogue (TMEM lane i <-> tile row i, so every thread owns one output row).
//   warp  4    TMEM allocator + the single thread that issues tcgen05.mma / tcgen05.commit.
//   warp  5    one thread streaming pre-packed weight tiles with 1-D bulk TMA copies.
// kStages-deep mbarrier ring (full: 4 producer warps + TMA tx bytes; empty: tcgen05.commit).
//
// Shared-memory operand layout (no swizzle): element (row r, col k) of a stage lives at byte
//   (k / 4) * rows * 16 + r * 16 + (k % 4) * 4        rows = 128 for A, n_tile for B
// i.e. [K/4][rows][4 floats]: 8 consecutive rows x 16 B form one 128-byte UMMA core matrix,
// SBO = 128 B between 8-row groups, LBO = rows*16 B between the K chunks.
#include <math.h>

#include "common.cuh"
#include "sm100.cuh"

namespace stf {
namespace {

using namespace sm100;

constexpr int kTileM = 128;
constexpr int kBlockK = 16;  // floats per pipeline stage (2 MMAs of K=8); every K here is a multiple of 16
constexpr int kStages = 5;
constexpr int kThreads = 192;
constexpr int kChunks = kBlockK / 4;  // 16-byte K chunks per stage
constexpr int kMaxNTile = 192;

struct LinearParams {
  stf_linear_args a;
  int n_tile;
  int k_blocks;
  int tmem_cols;
  uint32_t idesc;
  // window geometry
  int Hp, Wp, nWw, nW;  // padded size, windows per row, windows per image
};

struct RowSrc {  // where one A-tile row comes from
  const float *p;  // base pointer of the row (part 0 for MERGE); nullptr = all-zero row
  int merge_flags; // MERGE: bit0 = row 2i+1 valid, bit1 = col 2j+1 valid
};

__device__ __forceinline__ int window_row_to_token(const LinearParams &P, int g, bool *valid) {
  const stf_linear_args &a = P.a;
  const int ws = a.window, N = ws * ws;
  int b = g / (P.nW * N);
  int rem = g - b * (P.nW * N);
  int wi = rem / N, n = rem - wi * N;
  int wy = wi / P.nWw, wx = wi - wy * P.nWw;
  int hs = wy * ws + n / ws, wsft = wx * ws + n % ws;  // coordinates in the shifted frame
  int h = hs + a.shift, w = wsft + a.shift;            // torch.roll(x, -shift): shifted[h'] = x[(h'+s) mod Hp]
  if (h >= P.Hp) h -= P.Hp;
  if (w >= P.Wp) w -= P.Wp;
  *valid = (h < a.H) && (w < a.W);
  return (b * a.H + h) * a.W + w;
}

__device__ __forceinline__ RowSrc row_source(const LinearParams &P, int row) {
  const stf_linear_args &a = P.a;
  RowSrc r{nullptr, 0};
  if (row >= a.M) return r;
  if (a.rows == STF_ROWS_DENSE) {
    r.p = a.x + (int64_t)row * a.ldx;
  } else if (a.rows == STF_ROWS_WINDOW) {
    bool valid;
    int tok = window_row_to_token(P, row, &valid);
    if (valid) r.p = a.x + (int64_t)tok * a.ldx;
  } else {  // MERGE: output token (b, i, j) over ceil(H/2) x ceil(W/2)
    int H2 = (a.H + 1) >> 1, W2 = (a.W + 1) >> 1;
    int b = row / (H2 * W2);
    int rem = row - b * (H2 * W2);
    int i = rem / W2, j = rem - i * W2;
    r.p = a.x + ((int64_t)(b * a.H + 2 * i) * a.W + 2 * j) * a.ldx;
    r.merge_flags = ((2 * i + 1 < a.H) ? 1 : 0) | ((2 * j + 1 < a.W) ? 2 : 0);
  }
  return r;
}

// float4 of row `r` at K-offset k (k % 4 == 0); zeros where the reference pads.
__device__ __forceinline__ float4 load_chunk(const LinearParams &P, const RowSrc &r, int k) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!r.p) return z;
  if (P.a.rows != STF_ROWS_MERGE) return __ldg(reinterpret_cast<const float4 *>(r.p + k));
  const int C = P.a.K >> 2;
  int part = k / C, c = k - part * C;  // concat order x0,x1,x2,x3 = (0,0),(1,0),(0,1),(1,1)
  int di = part & 1, dj = part >> 1;
  if ((di && !(r.merge_flags & 1)) || (dj && !(r.merge_flags & 2))) return z;
  return __ldg(reinterpret_cast<const float4 *>(r.p + ((int64_t)di * P.a.W + dj) * P.a.ldx + c));
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__global__ void __launch_bounds__(kThreads)
linear_tf32_kernel(const __grid_constant__ LinearParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const stf_linear_args &a = P.a;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kTileM;
  const int nt = blockIdx.y;
  const int NT = P.n_tile;

  // ---- shared memory carve-up
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem);
  uint64_t *empty_bar = full_bar + kStages;
  uint64_t *accum_bar = empty_bar + kStages;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_bar + 1);
  const uint32_t a_stage_bytes = kChunks * kTileM * 16;
  const uint32_t b_stage_bytes = kChunks * NT * 16;
  uint8_t *a_smem = smem + 128;
  uint8_t *b_smem = a_smem + kStages * a_stage_bytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 4 + 1);  // 4 producer warps + the weight loader's expect_tx arrive
      mbar_init(&empty_bar[s], 1);     // one tcgen05.commit
    }
    mbar_init(accum_bar, 1);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // =========================== A producers ===========================
    // lane -> (row sub-index lane&7, K chunk lane>>3); 4 row groups of 8 per warp.
    const int sub = lane & 7, chunk = lane >> 3;
    RowSrc src[4];
    float mean[4], rstd[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) src[i] = row_source(P, m0 + warp * 32 + i * 8 + sub);

    const bool has_ln = a.ln_gamma != nullptr;
    if (has_ln) {
      // two-pass LayerNorm statistics; the 4 lanes sharing a row combine with shuffles
      const float inv_k = 1.0f / (float)a.K;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float s = 0.f;
        for (int k = chunk * 4; k < a.K; k += 16) {
          float4 v = load_chunk(P, src[i], k);
          s += (v.x + v.y) + (v.z + v.w);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        const float mu = s * inv_k;
        float q = 0.f;
        for (int k = chunk * 4; k < a.K; k += 16) {
          float4 v = load_chunk(P, src[i], k);
          float dx = v.x - mu, dy = v.y - mu, dz = v.z - mu, dw = v.w - mu;
          q += (dx * dx + dy * dy) + (dz * dz + dw * dw);
        }
        q += __shfl_xor_sync(0xffffffffu, q, 8);
        q += __shfl_xor_sync(0xffffffffu, q, 16);
        mean[i] = mu;
        rstd[i] = rsqrtf(q * inv_k + a.ln_eps);
      }
    }
    // WINDOW pad rows are zero AFTER the LayerNorm (stf.py:155-162); MERGE pads before it.
    for (int kb = 0; kb < P.k_blocks; ++kb) {
      const int s = kb % kStages;
      const uint32_t it = (uint32_t)(kb / kStages);
      mbar_wait(&empty_bar[s], (it & 1u) ^ 1u);
      const int k = kb * kBlockK + chunk * 4;
      float4 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = load_chunk(P, src[i], k);
      if (has_ln) {
        const float4 g = __ldg(reinterpret_cast<const float4 *>(a.ln_gamma + k));
        const float4 bt = __ldg(reinterpret_cast<const float4 *>(a.ln_beta + k));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (src[i].p) {
            v[i].x = (v[i].x - mean[i]) * rstd[i] * g.x + bt.x;
            v[i].y = (v[i].y - mean[i]) * rstd[i] * g.y + bt.y;
            v[i].z = (v[i].z - mean[i]) * rstd[i] * g.z + bt.z;
            v[i].w = (v[i].w - mean[i]) * rstd[i] * g.w + bt.w;
          }
        }
      }
      uint8_t *stage = a_smem + s * a_stage_bytes + chunk * (kTileM * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 t = make_float4(to_tf32(v[i].x), to_tf32(v[i].y), to_tf32(v[i].z), to_tf32(v[i].w));
        *reinterpret_cast<float4 *>(stage + (warp * 32 + i * 8 + sub) * 16) = t;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[s]);
    }

    // =========================== epilogue ===========================
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int row = m0 + warp * 32 + lane;  // TMEM lane == tile row
    float *dst = nullptr;
    const float *res = nullptr;
    int out_tok00 = 0;  // PIXEL_SHUFFLE: output token (2h, 2w)
    if (row < a.M) {
      if (a.epilogue == STF_EPI_WINDOW_RESIDUAL) {
        bool valid;
        int tok = window_row_to_token(P, row, &valid);
        if (valid) {
          dst = a.y + (int64_t)tok * a.ldy;
          res = a.residual + (int64_t)tok * a.ldy;
        }
      } else if (a.epilogue == STF_EPI_PIXEL_SHUFFLE) {
        int b = row / (a.H * a.W);
        int rem = row - b * (a.H * a.W);
        int h = rem / a.W, w = rem - h * a.W;
        out_tok00 = (b * 2 * a.H + 2 * h) * (2 * a.W) + 2 * w;
        dst = a.y;
      } else {
        dst = a.y + (int64_t)row * a.ldy;
        if (a.epilogue == STF_EPI_RESIDUAL) res = a.residual + (int64_t)row * a.ldy;
      }
    }
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < NT; c0 += 16) {
      float acc[16];
      tmem_ld16(lane_base + (uint32_t)c0, acc);  // warp-collective: executed by all lanes
      if (!dst) continue;
      const int n0 = nt * NT + c0;
      if (a.bias) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 bv = __ldg(reinterpret_cast<const float4 *>(a.bias + n0 + j));
          acc[j] += bv.x, acc[j + 1] += bv.y, acc[j + 2] += bv.z, acc[j + 3] += bv.w;
        }
      }
      if (a.epilogue == STF_EPI_QKV) {
        if (n0 < a.q_cols) {  // q_cols is a multiple of 16 (head_dim % 8 == 0, C % 16 == 0)
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] *= a.q_scale;
        }
      } else if (a.epilogue == STF_EPI_GELU) {
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = gelu_erf(acc[j]);
      } else if (res) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 rv = __ldg(reinterpret_cast<const float4 *>(res + n0 + j));
          acc[j] = rv.x + acc[j], acc[j + 1] = rv.y + acc[j + 1], acc[j + 2] = rv.z + acc[j + 2],
          acc[j + 3] = rv.w + acc[j + 3];
        }
      }
      if (a.epilogue == STF_EPI_PIXEL_SHUFFLE) {
        // feature f = 4*c + 2*i + j -> token (2h+i, 2w+j), channel c
        const int cbase = n0 >> 2;
#pragma unroll
        for (int ij = 0; ij < 4; ++ij) {
          int tok = out_tok00 + (ij >> 1) * (2 * a.W) + (ij & 1);
          float4 o = make_float4(acc[ij], acc[4 + ij], acc[8 + ij], acc[12 + ij]);
          *reinterpret_cast<float4 *>(dst + (int64_t)tok * a.ldy + cbase) = o;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4 *>(dst + n0 + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      // K-major, no swizzle: LBO = distance between the two 16-byte K chunks of one MMA,
      // SBO = distance between consecutive 8-row core matrices (verified on B200 at bring-up).
      const uint32_t a_lbo = (uint32_t)(kTileM * 16), a_sbo = 128u;
      const uint32_t b_lbo = (uint32_t)(NT * 16), b_sbo = 128u;
      for (int kb = 0; kb < P.k_blocks; ++kb) {
        const int s = kb % kStages;
        const uint32_t it = (uint32_t)(kb / kStages);
        mbar_wait(&full_bar[s], it & 1u);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(a_smem + s * a_stage_bytes);
        const uint32_t b_addr = smem_u32(b_smem + s * b_stage_bytes);
#pragma unroll
        for (int ks = 0; ks < kBlockK / 8; ++ks) {  // one MMA consumes K = 8 tf32 = 2 chunks
          uint64_t da = umma_smem_desc(a_addr + ks * 2 * (kTileM * 16), a_lbo, a_sbo);
          uint64_t db = umma_smem_desc(b_addr + ks * 2 * (NT * 16), b_lbo, b_sbo);
          umma_tf32(tmem_base, da, db, P.idesc, (kb | ks) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the stage once these MMAs have read it
      }
      umma_commit(accum_bar);  // accumulator complete -> epilogue
    }
    __syncwarp();
  } else {
    // =========================== weight loader (bulk TMA) ===========================
    if (lane == 0) {
      const float *wt = a.w_packed + (size_t)nt * (size_t)(a.K >> 2) * NT * 4;
      for (int kb = 0; kb < P.k_blocks; ++kb) {
        const int s = kb % kStages;
        const uint32_t it = (uint32_t)(kb / kStages);
        mbar_wait(&empty_bar[s], (it & 1u) ^ 1u);
        mbar_arrive_expect_tx(&full_bar[s], b_stage_bytes);
        bulk_copy_g2s(b_smem + s * b_stage_bytes, wt + (size_t)kb * kChunks * NT * 4, b_stage_bytes, &full_bar[s]);
      }
    }
    __syncwarp();
  }

  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
  }
}

__global__ void __launch_bounds__(256)
pack_weight_kernel(const float *__restrict__ w, float *__restrict__ packed, int N, int K, int NT) {
  // one thread per output float4: packed[nt][kc][n_in][0..3] = tf32(W[nt*NT + n_in][4*kc .. 4*kc+3])
  const int64_t total = (int64_t)N * (K >> 2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int n_in = (int)(i % NT);
    int64_t r = i / NT;
    int kc = (int)(r % (K >> 2));
    int t = (int)(r / (K >> 2));
    float4 v = __ldg(reinterpret_cast<const float4 *>(w + (int64_t)(t * NT + n_in) * K + kc * 4));
    reinterpret_cast<float4 *>(packed)[i] =
        make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
  }
}

size_t linear_smem_bytes(int n_tile) {
  return 128 + (size_t)kStages * (kChunks * kTileM * 16 + kChunks * n_tile * 16);
}

int launch_linear(const stf_linear_args *args, void *stream) {
  if (!args) return STF_E_ARG;
  const stf_linear_args &a = *args;
  if (!a.x || !a.w_packed || !a.y || a.M < 0 || a.N <= 0 || a.K <= 0) return STF_E_ARG;
  if (a.M == 0) return STF_OK;
  if (a.K % kBlockK != 0 || a.N % 16 != 0) return STF_E_SHAPE;
  if (a.ldx % 4 != 0 || a.ldy % 4 != 0) return STF_E_SHAPE;
  if (!aligned16(a.x) || !aligned16(a.w_packed) || !aligned16(a.y) || !aligned16(a.bias) ||
      !aligned16(a.residual) || !aligned16(a.ln_gamma) || !aligned16(a.ln_beta))
    return STF_E_ALIGN;
  if ((a.ln_gamma == nullptr) != (a.ln_beta == nullptr)) return STF_E_ARG;
  if ((a.epilogue == STF_EPI_RESIDUAL || a.epilogue == STF_EPI_WINDOW_RESIDUAL) && !a.residual) return STF_E_ARG;
  if (a.epilogue < STF_EPI_STORE || a.epilogue > STF_EPI_PIXEL_SHUFFLE) return STF_E_ARG;
  if (a.rows < STF_ROWS_DENSE || a.rows > STF_ROWS_MERGE) return STF_E_ARG;
  if (a.epilogue == STF_EPI_QKV && (a.q_cols % 16 != 0)) return STF_E_SHAPE;

  LinearParams P;
  P.a = a;
  P.n_tile = stf_linear_n_tile(a.N);
  if (P.n_tile <= 0) return STF_E_SHAPE;
  P.k_blocks = a.K / kBlockK;
  P.tmem_cols = 32;
  while (P.tmem_cols < P.n_tile) P.tmem_cols <<= 1;
  P.idesc = umma_idesc_tf32(kTileM, P.n_tile);
  P.Hp = P.Wp = P.nWw = P.nW = 0;
  const bool windowed = a.rows == STF_ROWS_WINDOW || a.epilogue == STF_EPI_WINDOW_RESIDUAL;
  if (windowed) {
    if (a.window <= 0 || a.batch <= 0 || a.H <= 0 || a.W <= 0 || a.shift < 0 || a.shift >= a.window) return STF_E_SHAPE;
    P.Hp = (a.H + a.window - 1) / a.window * a.window;
    P.Wp = (a.W + a.window - 1) / a.window * a.window;
    P.nWw = P.Wp / a.window;
    P.nW = (P.Hp / a.window) * P.nWw;
    if ((int64_t)a.M != (int64_t)a.batch * P.Hp * P.Wp) return STF_E_SHAPE;
  }
  if (a.rows == STF_ROWS_MERGE) {
    if (a.batch <= 0 || a.H <= 0 || a.W <= 0 || a.K % 4 != 0 || (a.K / 4) % 4 != 0) return STF_E_SHAPE;
    if ((int64_t)a.M != (int64_t)a.batch * ((a.H + 1) / 2) * ((a.W + 1) / 2)) return STF_E_SHAPE;
  }
  if (a.epilogue == STF_EPI_PIXEL_SHUFFLE) {
    if (a.batch <= 0 || a.H <= 0 || a.W <= 0 || (int64_t)a.M != (int64_t)a.batch * a.H * a.W) return STF_E_SHAPE;
    if (a.N % 16 != 0 || a.ldy < a.N / 4) return STF_E_SHAPE;
  }
  const size_t smem = linear_smem_bytes(P.n_tile);
  static std::atomic<int> attr_set{0};
  if (!attr_set.load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(linear_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)linear_smem_bytes(kMaxNTile));
    if (e != cudaSuccess) return (int)e;
    attr_set.store(1, std::memory_order_release);
  }
  dim3 grid((a.M + kTileM - 1) / kTileM, a.N / P.n_tile);
  linear_tf32_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(P);
  return check_launch();
}

}  // namespace
}  // namespace stf

using namespace stf;

extern "C" int stf_linear_n_tile(int N) {
  if (N <= 0 || N % 16 != 0) return STF_E_SHAPE;
  for (int nt = kMaxNTile; nt >= 16; nt -= 16)
    if (N % nt == 0) return nt;
  return STF_E_SHAPE;
}

extern "C" int stf_pack_linear_weight(const float *weight, float *packed, int N, int K, void *stream) {
  if (!weight || !packed || N <= 0 || K <= 0) return STF_E_ARG;
  if (K % kBlockK != 0) return STF_E_SHAPE;
  int nt = stf_linear_n_tile(N);
  if (nt <= 0) return STF_E_SHAPE;
  if (!aligned16(weight) || !aligned16(packed)) return STF_E_ALIGN;
  int64_t total = (int64_t)N * (K / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  pack_weight_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(weight, packed, N, K, nt);
  return check_launch();
}

extern "C" int stf_linear(const stf_linear_args *args, void *stream) { return launch_linear(args, stream); }
