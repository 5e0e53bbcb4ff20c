// Implicit-GEMM 2-D convolution on 5th-gen tensor cores, NHWC, TMA in / TMA out:
//
//   Y[b, oy, ox, n] = act( sum_{ky,kx,c} X[b, oy*s + ky - p, ox*s + kx - p, c] * W[n, c, ky, kx] + bias[n] )
//
// with X the channel concatenation of up to three NHWC tensors (the slice loop's torch.cat([latent, y_hat_0..]) never
// materialises), p = k/2 zero padding, s in {1, 2}, and an optional PixelShuffle(2) folded into the store.
//
// Replaces (reference memory4963/STF): the five-layer cc_mean / cc_scale / lrp stacks of the slice loop
// (compressai/models/stf.py:510-548, called at :613-633, :706-729, :757-779; cnn.py:89-127), the hyperprior h_a / h_mean_s /
// h_scale_s (stf.py:472-509) incl. subpel_conv3x3 = conv + PixelShuffle (layers/layers.py:47-51) and the 5x5 end_conv
// (stf.py:466); i.e. nn.Conv2d + bias + nn.GELU (+ torch.cat in front, + nn.PixelShuffle behind).
//
// Why our own kernel and not cuDNN: (1) the decoder must rebuild the encoder's indexes bit for bit (stf.py:767), so mu and
// scale must not depend on how many images share a launch -- here every output element is one fixed-order K loop
// (tap-major, then source, then channel; k-steps of 8) whatever the batch, tile shape or N tiling, so strings are
// batch-invariant by construction; (2) the concat, the bias + GELU and the pixel shuffle cost nothing extra.
//
// GEMM view: M = output pixels (a 128-pixel tile is a TW x TH box of one image), N = output channels, K = taps x channels.
//   A operand  k-block (tap, 32-channel block of one source): ONE 4-D tensor-map TMA load of the box
//              [32 ch][TW*s][TH*s][1 image] at (c0, ox0*s + kx - p, oy0*s + ky - p, b) -- the halo, the zero padding and
//              the ragged right/bottom edge are TMA out-of-bounds zero fill, stride 2 is the map's element stride --
//              landing as 128 rows x 128 B in the SWIZZLE_128B K-major layout tcgen05.mma reads.
//   B operand  pre-packed weights [N][taps * Cpad] (hi plane, and a lo plane for the 3xTF32 mode), 2-D tensor map, box [32][n_tile].
//   D          fp32 accumulators in TMEM, double buffered (epilogue of tile i under the K loop of tile i+1).
//   epilogue   tcgen05.ld -> + bias -> exact-erf GELU -> swizzled staging tile -> 4-D tensor-map TMA store (edge clipping and
//              the pixel-shuffle scatter are the store map's geometry).
// Warp roles (512 threads, one persistent CTA per SM): warp 0 TMA producer, warp 1 MMA issuer (+ TMEM alloc), warps 4-11 epilogue
// (warp & 3 = TMEM lane quadrant, two warps per quadrant on alternating column chunks), warps 12-15 A pass (3xTF32: A stage ->
// truncated hi in place + lo plane; LayerNorm: row statistics).
#include <cuda.h>  // CUtensorMap + enums only; cuTensorMapEncodeTiled is resolved at run time (no link-time libcuda dependency)
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "epilogue_math.cuh"
#include "sm100.cuh"

namespace stf {
namespace {

using namespace sm100;

constexpr int kTileM = 128;
constexpr int kBlockK = 32;                       // floats per k-block = one 128-byte swizzle row
constexpr uint32_t kAStageBytes = kTileM * 128;   // 16 KB
constexpr int kMaxStages = 8;
// Halo mode (3x3, stride 1): the A operand of all nine taps is ONE TMA load per 32-channel block -- the (TH + 2) x 16-pixel
// halo of an 8-wide x 16-tall output tile; tap (ky, kx) is the same shared-memory tile read through a descriptor whose start
// is shifted by (ky * 16 + kx) pixels (8-row core-matrix groups = the 8 pixels of one output row, group stride = one halo
// row = 2048 B).  L2 -> shared-memory traffic of the A operand drops 4x (it bounds the kernel: ~53 B/clk/SM).
constexpr int kHaloW = 16, kHaloTW = 8, kHaloTH = 16;
constexpr uint32_t kHaloBytes = (kHaloTH + 2) * kHaloW * 128;   // 36 KB per 32-channel block
constexpr int kEpiWarps = 8;   // two per TMEM lane quadrant (alternating column chunks): one warp per scheduler cannot hide its
                               // own ALU / LDS / MUFU latencies (measured: the GELU epilogue of a 192-column tile took 38 k cycles)
constexpr int kThreads = 512;
constexpr int kProducerWarp = 0, kMmaWarp = 1, kFirstEpiWarp = 4, kFirstSplitWarp = kFirstEpiWarp + kEpiWarps;
constexpr int kMaxSrc = 3;
constexpr uint32_t kSlabBytes = 32 * 128;         // one epilogue warp's 32 rows x 32 floats (or 32 x 16 in 64-byte rows)
constexpr int kMaxSlabs = 4;                      // staging slabs per epilogue warp (2..4: bulk stores in flight per warp + 1)
constexpr int kStatSlots = 12;                    // (mean, rstd) of the 128 rows of a tile, ring over the tiles in flight: the
                                                  // A pass leads the epilogue by at most kMaxStages k-blocks + 2 tiles

struct ConvParams {
  alignas(64) CUtensorMap a_map[kMaxSrc];
  alignas(64) CUtensorMap b_map[2];   // hi, lo
  alignas(64) CUtensorMap y_map[4];   // plain: [0]; pixel shuffle: one per (i, j) sub-pixel
  const float *bias;                  // N floats in packed column order
  const float *residual;              // act 2: NHWC tensor of the output's geometry (may alias the output), pixel stride res_ld
  int res_ld, Ho, Wo;
  int n_src;
  int src_kb[kMaxSrc];                // 32-channel k-blocks per tap of each source
  int kb_per_tap, k_blocks, ksize, pad, stride;
  int tail_ksteps;                    // k-steps of 8 with real data in the LAST k-block (4 unless a Linear layer's K % 32 != 0)
  int N, n_tile, n_tiles, n_pad;      // n_pad = n_tiles * n_tile rounded up to 32: length of the shared-memory epilogue vectors
  int TW, TH, tiles_x, tiles_y, tiles_per_img, m_tiles, total_tiles;
  int act;                            // 0 none, 1 exact-erf GELU, 2 residual + 0.5 * tanh(.)  (the LRP tail, stf.py:631-633),
                                      // 3 residual + (.)  (fc2 / proj + shortcut, stf.py:196-197)
  int ln_k;                           // has_ln: number of real input features (K before padding to 32)
  int has_ln;                         // LayerNorm over the K inputs of every row folded through the GEMM (see pack)
  float ln_eps;
  const float *svec;                  // has_ln: s[N] = sum_k (gamma o W)[n][k]
  int shuffle_cout;                   // 0: plain store; else channels after PixelShuffle(2) (N = 4 * shuffle_cout)
  int cw;                             // store chunk width in channels: 32 (SWIZZLE_128B staging) or 16 (SWIZZLE_64B)
  int stages;                         // operand ring (halo mode: halo tiles)
  int b_stages;                       // halo mode: weight-tile ring
  int slabs;                          // staging slabs per epilogue warp
  int debug;                          // bring-up (env STF_B200_CONV_DEBUG): 1 no stores, 2 no staging writes, 4 no epilogue math, 8 no A loads
  uint32_t idesc;
  int tmem_cols, acc_stride;
  uint32_t stage_bytes, b_plane_bytes;  // per pipeline stage: A hi (+ A lo) + B hi (+ B lo)
};

// ---------------------------------------------------------------------------- PTX: tensor-map TMA
__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_dst),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, uint32_t smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
               "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// Shared-memory matrix descriptor, K-major operand in the SWIZZLE_128B layout (rows of 128 B, 8-row swizzle atoms of
// 1024 B): start address >> 4 in [0,14), LBO (unused for swizzled K-major) = 1 in [16,30), SBO = 1024 B >> 4 in [32,46),
// descriptor version 1 in [46,48), layout type 2 (SWIZZLE_128B) in [61,64).  Stage bases are 1024-byte aligned (base
// offset field 0); a k-step of 8 tf32 inside the 128-byte row advances the start address by 32 B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread; the second half is skipped (warp-uniformly) when
// `second` is 0.  Loads and tcgen05.wait::ld sit in ONE asm statement, so no use of the outputs can be scheduled above the wait.
// Same, for an operand whose 8-row groups are `sbo` bytes apart (and an explicit base_offset field [49,52), kept 0: see the
// halo-mode MMA loop).
__device__ __forceinline__ uint64_t umma_desc_sw128_ex(uint32_t smem_addr, uint32_t sbo, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t second, uint32_t (&r)[32]) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %33, 0;\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%32];\n\t"
      "@p tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%34];\n\t"
      "tcgen05.wait::ld.sync.aligned;\n\t}"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
        "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
        "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr), "r"(second), "r"(taddr + 16u)
      : "memory");
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// Epilogue math of one chunk (32 accumulator columns of one row), act fixed at compile time: the 32 element chains are
// independent and branch-free, bias / LayerNorm vectors come as 128-bit broadcast loads.  Columns past the tile or past N carry
// finite garbage: they land in staging columns the TMA store clips.
template <int kLn, int kAct>
__device__ __forceinline__ void epi_math(const uint32_t (&r)[32], float (&v)[32], const float *__restrict__ tv,
                                         const float *__restrict__ sv, float mean, float rstd,
                                         const float *__restrict__ res_row, int res_cols) {
  using namespace epi;
  if (kAct == 1) {   // GELU (every conv stack layer but the last, fc1): pairs of columns on the packed fp32 pipe
    const uint64_t nmean2 = dup2(-mean), rstd2 = dup2(rstd);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 t4 = *reinterpret_cast<const float4 *>(tv + j);
      float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kLn) s4 = *reinterpret_cast<const float4 *>(sv + j);
#pragma unroll
      for (int q = 0; q < 4; q += 2) {
        const uint64_t acc = pack2(__uint_as_float(r[j + q]), __uint_as_float(r[j + q + 1]));
        const uint64_t t2 = q ? pack2(t4.z, t4.w) : pack2(t4.x, t4.y);
        uint64_t a;
        if (kLn) a = fma2(rstd2, fma2(nmean2, q ? pack2(s4.z, s4.w) : pack2(s4.x, s4.y), acc), t2);   // rstd (acc - mean s) + t
        else a = add2(acc, t2);
        unpack2(gelu_erf2(a), v[j + q], v[j + q + 1]);
      }
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 t4 = *reinterpret_cast<const float4 *>(tv + j);
    float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), r4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kLn) s4 = *reinterpret_cast<const float4 *>(sv + j);
    if (kAct >= 2 && res_row && j < res_cols) r4 = *reinterpret_cast<const float4 *>(res_row + j);
    const float tt[4] = {t4.x, t4.y, t4.z, t4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w}, rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float a = __uint_as_float(r[j + q]);
      a = kLn ? fmaf(rstd, a - mean * ss[q], tt[q]) : a + tt[q];
      if (kAct == 2) a = rr[q] + 0.5f * tanhf(a);
      else if (kAct == 3) a = rr[q] + a;
      v[j + q] = a;
    }
  }
}

// ---------------------------------------------------------------------------- the kernel
template <int kPrecise, int kLn, int kHalo>
__global__ void __launch_bounds__(kThreads, 1) conv_tf32_kernel(const __grid_constant__ ConvParams P) {
  constexpr bool kAPass = kPrecise || kLn;   // warps 8-11 touch every landed A stage (hi / lo split and / or row statistics)
  extern __shared__ uint8_t smem_raw[];
  // operand stages need 1024-byte alignment (swizzle atoms); the launch asks for 1 KB of slack
  const uint32_t raw_u32 = smem_u32(smem_raw);
  uint8_t *smem = smem_raw + ((1024u - (raw_u32 & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t S = (uint32_t)P.stages;
  const uint32_t a_lo_off = kHalo ? kHaloBytes : kAStageBytes;             // precise: A lo plane behind A hi
  const uint32_t b_off = kPrecise ? 2 * kAStageBytes : kAStageBytes;       // B hi plane (tap mode: behind the A planes of the stage)
  const uint32_t b_lo_off = b_off + P.b_plane_bytes;
  const uint32_t SB = (uint32_t)P.b_stages;
  const uint32_t b_stage_bytes = (uint32_t)(kPrecise ? 2 : 1) * P.b_plane_bytes;
  uint8_t *ring = smem;
  uint8_t *b_ring = ring + (size_t)S * P.stage_bytes;     // halo mode: the weight tiles have their own ring
  uint8_t *staging = b_ring + (kHalo ? (size_t)SB * b_stage_bytes : 0);   // [8 warps][slabs] x 4 KB, 1024-aligned
  float *bias_s = reinterpret_cast<float *>(staging + (size_t)kEpiWarps * P.slabs * kSlabBytes);   // t[Npad] (bias, or beta.W^T + bias)
  float *svec_s = bias_s + P.n_pad;                                         // s[Npad] (LayerNorm fold); Npad covers every chunk, zero past N
  float2 *stats = reinterpret_cast<float2 *>(svec_s + (kLn ? P.n_pad : 0));   // [kStatSlots][128]
  uint64_t *bars = reinterpret_cast<uint64_t *>(stats + (kLn ? kStatSlots * kTileM : 0));
  uint64_t *full = bars, *empty = full + kMaxStages, *split = empty + kMaxStages, *acc_full = split + kMaxStages,
           *acc_empty = acc_full + 2, *full_b = acc_empty + 2, *empty_b = full_b + kMaxStages;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(empty_b + kMaxStages);

  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);    // the producer's arrive.expect_tx (+ TMA transaction bytes)
      mbar_init(&empty[s], 1);   // one tcgen05.commit
      mbar_init(&split[s], 4);   // one arrival per splitter warp
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], kEpiWarps);  // one arrival per epilogue warp
    }
    if (kHalo)
      for (uint32_t s = 0; s < SB; ++s) {
        mbar_init(&full_b[s], 1);
        mbar_init(&empty_b[s], 1);
      }
    mbar_fence_init();
    for (int s = 0; s < P.n_src; ++s) tma_prefetch_desc(&P.a_map[s]);
    tma_prefetch_desc(&P.b_map[0]);
    if (kPrecise) tma_prefetch_desc(&P.b_map[1]);
    tma_prefetch_desc(&P.y_map[0]);
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kProducerWarp) {
    // =========================== TMA producer ===========================
    // The whole warp runs the loop on warp-uniform values and ONE elected lane issues: ptxas then keeps tensor-map
    // pointers, coordinates and barrier addresses in uniform registers.  (A lane-0-only loop made it wrap every TMA / MMA
    // instruction in an ELECT + R2UR.BROADCAST "waterfall" loop: ~600 cycles of issue overhead per k-block, measured.)
    {
      const bool leader = elect_one();
      uint32_t st = 0, ph = 1;  // waiting on parity 1 of a fresh barrier returns immediately
      uint32_t hb_st = 0, hb_ph = 1;
      const uint32_t tx_bytes = kAStageBytes + (uint32_t)(kPrecise ? 2 : 1) * P.b_plane_bytes;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        const int mt = tile / P.n_tiles, nt = tile - mt * P.n_tiles;
        const int b = mt / P.tiles_per_img, rem = mt - b * P.tiles_per_img;
        const int ty = rem / P.tiles_x, tx = rem - ty * P.tiles_x;
        const int x0 = tx * P.TW * P.stride - P.pad, y0 = ty * P.TH * P.stride - P.pad;
        const int n0 = nt * P.n_tile;
        int kb = 0;
        if (kHalo) {
          // channel-block major: one halo tile per 32-channel block, then the nine weight tiles that read it
          int cbg = 0;   // channel block index over the concatenated sources
          for (int s = 0; s < P.n_src; ++s) {
            for (int cb = 0; cb < P.src_kb[s]; ++cb, ++cbg) {
              mbar_wait(&empty[st], ph);
              if (leader) mbar_arrive_expect_tx(&full[st], kHaloBytes);
              if (!leader) {
              } else if (P.debug & 8) asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[st])), "r"(kHaloBytes) : "memory");
              else tma_load_4d(smem_u32(ring) + st * P.stage_bytes, &P.a_map[s], &full[st], cb * kBlockK, x0, y0, b);
              if (++st == S) st = 0, ph ^= 1u;
              for (int tap = 0; tap < 9; ++tap) {
                mbar_wait(&empty_b[hb_st], hb_ph);
                const uint32_t bb = smem_u32(b_ring) + hb_st * b_stage_bytes;
                if (leader) mbar_arrive_expect_tx(&full_b[hb_st], b_stage_bytes);
                if (!leader) {
                } else if (P.debug & 64) {
                  asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full_b[hb_st])), "r"(b_stage_bytes) : "memory");
                } else {
                  tma_load_2d(bb, &P.b_map[0], &full_b[hb_st], (tap * P.kb_per_tap + cbg) * kBlockK, n0);
                  if (kPrecise) tma_load_2d(bb + P.b_plane_bytes, &P.b_map[1], &full_b[hb_st], (tap * P.kb_per_tap + cbg) * kBlockK, n0);
                }
                if (++hb_st == SB) hb_st = 0, hb_ph ^= 1u;
              }
            }
          }
          continue;
        }
        for (int tap = 0; tap < P.ksize * P.ksize; ++tap) {
          const int ky = tap / P.ksize, kx = tap - ky * P.ksize;
          for (int s = 0; s < P.n_src; ++s) {
            for (int cb = 0; cb < P.src_kb[s]; ++cb, ++kb) {
              mbar_wait(&empty[st], ph);
              const uint32_t base = smem_u32(ring) + st * P.stage_bytes;
              if (leader) mbar_arrive_expect_tx(&full[st], tx_bytes);
              if (!leader) {
              } else if (P.debug & 8) asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[st])), "r"(kAStageBytes) : "memory");
              else tma_load_4d(base, &P.a_map[s], &full[st], cb * kBlockK, x0 + kx, y0 + ky, b);
              if (!leader) {
              } else if (P.debug & 64) {
                asm volatile("mbarrier.complete_tx.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[st])), "r"(tx_bytes - kAStageBytes) : "memory");
              } else {
                tma_load_2d(base + b_off, &P.b_map[0], &full[st], kb * kBlockK, n0);
                if (kPrecise) tma_load_2d(base + b_lo_off, &P.b_map[1], &full[st], kb * kBlockK, n0);
              }
              if (++st == S) st = 0, ph ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // =========================== MMA issuer ===========================
    // Whole warp, warp-uniform loop; one elected lane issues tcgen05.mma / commit (descriptors stay in uniform registers).
    {
      const bool leader = elect_one();
      uint32_t st = 0, ph = 0, hb_st = 0, hb_ph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&acc_empty[buf], (((uint32_t)it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * P.acc_stride);
        if (kHalo) {
          for (int cbg = 0; cbg < P.kb_per_tap; ++cbg) {
            mbar_wait(kAPass ? &split[st] : &full[st], ph);
            const uint32_t hbase = smem_u32(ring) + st * P.stage_bytes;
#pragma unroll   // (fully unrolled: ky / kx are constants, so the nine descriptors stay in the uniform datapath)
            for (int tap = 0; tap < 9; ++tap) {
              const int ky = tap / 3, kx = tap - ky * 3;
              mbar_wait(&full_b[hb_st], hb_ph);
              tc_fence_after();
              // rows of the tap's A operand: output row g = halo row g + ky, pixels kx .. kx + 7 -> group stride = one halo row
              const uint32_t a_addr = hbase + (uint32_t)((ky * kHaloW + kx) * 128);
              // base_offset stays 0: measured on B200, the tensor core derives the swizzle phase from the absolute
              // shared-memory address bits [7,10) -- the same rule the TMA write used -- so a start that is shifted by
              // kx rows inside the 1024-byte atom reads back consistently (base_offset = kx gives wrong results).
              const uint64_t da = umma_desc_sw128_ex(a_addr, kHaloW * 128, 0u);
              const uint64_t dal = umma_desc_sw128_ex(a_addr + a_lo_off, kHaloW * 128, 0u);
              const uint64_t db = umma_desc_sw128(smem_u32(b_ring) + hb_st * b_stage_bytes);
#pragma unroll
              for (int ks = 0; ks < kBlockK / 8; ++ks) {
                const uint64_t dak = da + (uint64_t)(ks * 2), dbk = db + (uint64_t)(ks * 2);
                if (leader) umma_tf32(d_tmem, dak, dbk, P.idesc, (cbg | tap | ks) ? 1u : 0u);
                if (kPrecise && leader) {
                  umma_tf32(d_tmem, dal + (uint64_t)(ks * 2), dbk, P.idesc, 1u);
                  umma_tf32(d_tmem, dak, dbk + (uint64_t)(P.b_plane_bytes >> 4), P.idesc, 1u);
                }
              }
              if (leader) umma_commit(&empty_b[hb_st]);
              if (++hb_st == SB) hb_st = 0, hb_ph ^= 1u;
            }
            if (leader) umma_commit(&empty[st]);   // the halo tile is free when its nine taps have been read
            if (++st == S) st = 0, ph ^= 1u;
          }
          if (leader) umma_commit(&acc_full[buf]);
          __syncwarp();
          continue;
        }
        for (int kb = 0; kb < P.k_blocks; ++kb) {
          mbar_wait(kAPass ? &split[st] : &full[st], ph);   // (the A pass has waited for the stage's TMA bytes)
          tc_fence_after();
          const uint32_t base = smem_u32(ring) + st * P.stage_bytes;
          const uint64_t da = umma_desc_sw128(base), db = umma_desc_sw128(base + b_off);
#pragma unroll
          for (int ks = 0; ks < kBlockK / 8; ++ks) {  // one MMA consumes K = 8 tf32 = 32 B of every row
            // Linear layers whose K is not a multiple of 32 (C = 48: 16 real columns in the last k-block): the k-steps that
            // only hold TMA zero fill add +0 to every accumulator -- skipped (6 instead of 8 k-steps at C = 48)
            if (kb == P.k_blocks - 1 && ks >= P.tail_ksteps) break;
            const uint64_t dak = da + (uint64_t)(ks * 2), dbk = db + (uint64_t)(ks * 2);
            if (leader) umma_tf32(d_tmem, dak, dbk, P.idesc, (kb | ks) ? 1u : 0u);
            if (kPrecise && leader) {  // 3xTF32: hi.hi + lo.hi + hi.lo (lo.lo is below fp32 round-off)
              umma_tf32(d_tmem, dak + (uint64_t)(a_lo_off >> 4), dbk, P.idesc, 1u);
              umma_tf32(d_tmem, dak, dbk + (uint64_t)(P.b_plane_bytes >> 4), P.idesc, 1u);
            }
          }
          if (leader) umma_commit(&empty[st]);  // frees the stage when the MMAs above have read it
          if (++st == S) st = 0, ph ^= 1u;
        }
        if (leader) umma_commit(&acc_full[buf]);
        __syncwarp();
      }
    }
  } else if (warp >= kFirstEpiWarp && warp < kFirstEpiWarp + kEpiWarps) {
    // =========================== epilogue ===========================
    // The epilogue vectors are loaded here, by the warps that use them, behind the CTA-wide start barrier: the producer's
    // first TMA loads no longer wait for these global loads (~0.7 us on the critical path of every launch; the slice loop
    // launches ~1000 short kernels per batch).
    for (int i = threadIdx.x - kFirstEpiWarp * 32; i < P.n_pad; i += kEpiWarps * 32) {
      bias_s[i] = (P.bias && i < P.N) ? __ldg(P.bias + i) : 0.f;
      if (kLn) svec_s[i] = i < P.N ? __ldg(P.svec + i) : 0.f;
    }
    named_bar_sync(1, kEpiWarps * 32);
    const int quad = warp & 3;
    const int row = quad * 32 + lane;  // TMEM lane == tile row == pixel (row / TW, row % TW) of the tile box
    // Each epilogue warp stores its own 32 rows: a private ring of `slabs` staging slabs and its own bulk-store groups
    // (lane 0), so `slabs - 1` TMA stores per warp stay in flight and no CTA-wide barrier sits in the chunk loop.  The
    // warp's rows are the sub-box (min(TW, 32) x 32 / min(TW, 32)) of the tile at (row0 % TW, row0 / TW).
    const int cw = P.cw;
    const uint32_t row_bytes = (uint32_t)cw * 4u;
    const uint32_t swz = cw == 32 ? (uint32_t)(row & 7) : (uint32_t)((row >> 1) & 3);
    const bool epi_leader = elect_one();   // the lane that issues this warp's TMA stores (and owns its bulk groups)
    const int ew = warp - kFirstEpiWarp, half = ew >> 2;   // the two warps of a quadrant take even / odd chunks
    const uint32_t slab0 = smem_u32(staging) + (uint32_t)(ew * P.slabs) * kSlabBytes;
    const int sub_x = (quad * 32) % P.TW, sub_y = (quad * 32) / P.TW;
    int it = 0;
    uint32_t slab = 0;  // running slab index of this warp
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const int mt = tile / P.n_tiles, nt = tile - mt * P.n_tiles;
      const int b = mt / P.tiles_per_img, rem = mt - b * P.tiles_per_img;
      const int ty = rem / P.tiles_x, tx = rem - ty * P.tiles_x;
      const int ox0 = tx * P.TW, oy0 = ty * P.TH;
      const float *res_pix = nullptr;
      if (P.act >= 2) {  // y_hat_slice + 0.5 * tanh(lrp) / shortcut + (.): the residual is the pixel's own channels
        const int oy = oy0 + row / P.TW, ox = ox0 + row % P.TW;
        if (oy < P.Ho && ox < P.Wo) res_pix = P.residual + ((int64_t)(b * P.Ho + oy) * P.Wo + ox) * P.res_ld;
      }
      mbar_wait_relaxed(&acc_full[buf], ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
      float mean = 0.f, rstd = 1.f;
      if (kLn) {   // written by the A pass before it published the tile's last k-block
        const float2 st2 = stats[(it % kStatSlots) * kTileM + row];
        mean = st2.x, rstd = st2.y;
      }
      const uint32_t t_acc = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * P.acc_stride);
      const int n_chunks = (P.n_tile + cw - 1) / cw;
      if (half >= n_chunks) {  // no chunk for this warp in this tile: it still has to release the accumulator
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
      }
      for (int c = half; c < n_chunks; c += 2) {
        const int col0 = c * cw;            // column inside the tile
        const int n0 = nt * P.n_tile + col0;  // packed output column
        uint32_t r[32];
        tmem_ld32(t_acc + (uint32_t)col0, (cw == 32 && col0 + 16 < P.n_tile) ? 1u : 0u, r);
        if (c + 2 >= n_chunks) {  // last TMEM read of this warp for this tile: hand the accumulator back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        float v[32];
        const float *res_row = res_pix ? res_pix + n0 : nullptr;
        const int res_cols = P.N - n0;      // residual columns that exist (guards the row end of the last chunk)
        if (P.debug & 4) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        } else {
          switch (P.act) {
            case 1: epi_math<kLn, 1>(r, v, bias_s + n0, svec_s + n0, mean, rstd, nullptr, 0); break;
            case 2: epi_math<kLn, 2>(r, v, bias_s + n0, svec_s + n0, mean, rstd, res_row, res_cols); break;
            case 3: epi_math<kLn, 3>(r, v, bias_s + n0, svec_s + n0, mean, rstd, res_row, res_cols); break;
            default: epi_math<kLn, 0>(r, v, bias_s + n0, svec_s + n0, mean, rstd, nullptr, 0); break;
          }
        }
        if (epi_leader) {  // the store that last read this slab (`slabs` chunks of this warp ago) has finished reading it
          if (P.slabs == 1) bulk_wait_read<0>(); else if (P.slabs == 2) bulk_wait_read<1>();
          else if (P.slabs == 3) bulk_wait_read<2>(); else bulk_wait_read<3>();
        }
        __syncwarp();
        const uint32_t src = slab0 + slab * kSlabBytes;
        const uint32_t stg = src + (uint32_t)lane * row_bytes;
        if (P.debug & 2) {
          if (v[0] == 1.2345e-30f) sts128(stg, make_float4(v[1], v[2], v[3], v[4]));   // keep v alive
        } else if (cw == 32) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts128(stg + (((uint32_t)j ^ swz) << 4), make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128(stg + (((uint32_t)j ^ swz) << 4), make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (epi_leader) {
          if (n0 >= P.N || (P.debug & 1)) {
            // chunk entirely past the last output channel (overhanging last column tile): nothing to store
          } else if (P.shuffle_cout) {
            const int g = n0 / P.shuffle_cout, c0 = n0 - g * P.shuffle_cout;  // sub-pixel (i, j) = (g >> 1, g & 1)
            tma_store_4d(&P.y_map[g], src, c0, ox0 + sub_x, oy0 + sub_y, b);
          } else {
            tma_store_4d(&P.y_map[0], src, n0, ox0 + sub_x, oy0 + sub_y, b);
          }
          bulk_commit();   // (an empty group when nothing was stored: keeps the slab <-> group bookkeeping uniform)
        }
        if (++slab == (uint32_t)P.slabs) slab = 0;
      }
    }
    // the staging slabs must outlive the stores' READS only; the writes are complete (and visible) when the grid has finished
    if (epi_leader) bulk_wait_read<0>();
  } else if (kAPass && warp >= kFirstSplitWarp) {
    // =========================== A pass: hi / lo split (3xTF32) and LayerNorm row statistics ===========================
    // Element-wise on the landed A stage.  3xTF32: the raw fp32 stage IS the hi operand (the tensor core reads the upper 19
    // bits of each word: hi = trunc_tf32(x), checked on B200 against an explicit rewrite); lo = x - hi (exact in fp32) goes
    // into the lo plane, of which the tensor core again reads the upper 11 significant bits: x = hi + lo to 2^-22 relative.  LayerNorm: shifted one-pass
    // sum / sum of squares of every row while it streams through (thread t owns the 16-byte position t % 8 of rows
    // t / 8 + 16 i; the swizzle only permutes chunks inside a row), reduced over the row's 8 lanes after the last k-block.
    const int st_thread = threadIdx.x - kFirstSplitWarp * 32;  // 0..127
    const int grp_lane0 = lane & ~7;                           // first lane of this thread's 8-lane row group
    const int first_pos = (st_thread >> 3) & 7;                // swizzled position of the row's logical chunk 0 (row & 7)
    uint32_t st = 0, ph = 0;
    int it = 0;
    float shift0[8], sum[8], sq[8];
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x, ++it) {
      if (kHalo) {  // 3xTF32: split each landed halo tile once (it serves all nine taps)
        for (int cbg = 0; cbg < P.kb_per_tap; ++cbg) {
          mbar_wait(&full[st], ph);
          const uint32_t base = smem_u32(ring) + st * P.stage_bytes + (uint32_t)st_thread * 16u;
#pragma unroll 6
          for (int i = 0; i < (int)(kHaloBytes / 2048); ++i) {
            const uint32_t a = base + (uint32_t)i * 2048u;
            const float4 x = lds128(a);
            const float4 hi = make_float4(trunc_tf32(x.x), trunc_tf32(x.y), trunc_tf32(x.z), trunc_tf32(x.w));
            sts128(a + a_lo_off, make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&split[st]);
          if (++st == S) st = 0, ph ^= 1u;
        }
        continue;
      }
      for (int kb = 0; kb < P.k_blocks; ++kb) {
        mbar_wait(&full[st], ph);
        const uint32_t base = smem_u32(ring) + st * P.stage_bytes + (uint32_t)st_thread * 16u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t a = base + (uint32_t)i * 2048u;
          const float4 x = lds128(a);
          if (kLn) {
            if (kb == 0) {  // shift by the row's first element: keeps the one-pass variance well conditioned
              shift0[i] = __shfl_sync(0xffffffffu, x.x, grp_lane0 + first_pos);
              sum[i] = 0.f, sq[i] = 0.f;
            }
            const float dx = x.x - shift0[i], dy = x.y - shift0[i], dz = x.z - shift0[i], dw = x.w - shift0[i];
            sum[i] += (dx + dy) + (dz + dw);
            sq[i] += (dx * dx + dy * dy) + (dz * dz + dw * dw);
          }
          if (kPrecise) {
            const float4 hi = make_float4(trunc_tf32(x.x), trunc_tf32(x.y), trunc_tf32(x.z), trunc_tf32(x.w));
            if (P.debug & 16) sts128(a, hi);   // (bring-up: rewrite hi explicitly.  Measured on B200: results are identical either
                                              // way -- tcgen05.mma kind::tf32 reads the upper 19 bits of an fp32 word, i.e. truncates)
            sts128(a + a_lo_off, make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w));
          }
        }
        if (kLn && kb == P.k_blocks - 1) {
          // Channels past K in the last k-block are TMA zero fill: each contributed (0 - shift)^n to the sums; remove them.
          const float inv_k = 1.0f / (float)P.ln_k;
          const float pad = (float)(P.k_blocks * kBlockK - P.ln_k);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float s1 = sum[i], s2 = sq[i];
            s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
            s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
            s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 4);
            s2 += __shfl_xor_sync(0xffffffffu, s2, 4);
            s1 += pad * shift0[i];
            s2 -= pad * shift0[i] * shift0[i];
            if ((lane & 7) == 0) {
              const float md = s1 * inv_k;
              const float var = fmaxf(s2 * inv_k - md * md, 0.f);
              stats[(it % kStatSlots) * kTileM + (st_thread >> 3) + 16 * i] = make_float2(shift0[i] + md, rsqrtf(var + P.ln_eps));
            }
          }
        }
        if (kPrecise) fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&split[st]);
        if (++st == S) st = 0, ph ^= 1u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
  }
}

// ---------------------------------------------------------------------------- weight packing
// out: [planes][N][Kp]  t[N]  s[N];  Kp = taps * Cp, Cp = sum_s ceil32(C_s).  Column n' of the packed matrix is output
// channel n: plain n' = n; pixel shuffle n' = g * Cout + c  <->  n = 4 c + g  (PixelShuffle(2): f = 4c + 2i + j, g = 2i + j).
// With a LayerNorm in front (Linear layers, ksize 1): the image holds gamma o W, and
//   LN(x) . W^T + bias = rstd * ( x . (gamma o W)^T - mean * s ) + t,   s = sum_k (gamma o W)[n][k],  t = beta . W^T + bias
// so the GEMM runs on the raw rows and the epilogue applies the row statistics.  Rows n' < scale_cols (the q third of a qkv
// Linear) are multiplied by row_scale (q * d^-1/2, stf.py:99; exact for d = 16) in W, s and t alike.
struct PackParams {
  const float *w;      // (N, Ctot, k, k) contiguous
  const float *bias;   // N or nullptr
  const float *gamma, *beta;   // Ctot each, or nullptr
  float *out;
  int N, Ctot, taps, Cp, Kp, n_src;
  int src_c[kMaxSrc], src_cp[kMaxSrc];
  int shuffle_cout, planes;
  int scale_cols;
  float row_scale;
};

__device__ __forceinline__ int pack_src_channel(const PackParams &P, int q) {
  int cbase = 0;
  for (int s = 0; s < P.n_src; ++s) {
    if (q < P.src_cp[s]) return q < P.src_c[s] ? cbase + q : -1;
    q -= P.src_cp[s];
    cbase += P.src_c[s];
  }
  return -1;
}

__global__ void pack_conv_kernel(const PackParams P) {
  const int64_t total = (int64_t)P.N * P.Kp;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int np = (int)(i / P.Kp), kp = (int)(i - (int64_t)np * P.Kp);
    const int n = P.shuffle_cout ? 4 * (np % P.shuffle_cout) + np / P.shuffle_cout : np;
    const int tap = kp / P.Cp;
    const int c = pack_src_channel(P, kp - tap * P.Cp);
    float v = c >= 0 ? P.w[((int64_t)n * P.Ctot + c) * P.taps + tap] : 0.f;
    if (c >= 0 && P.gamma) v *= P.gamma[c];
    if (np < P.scale_cols) v *= P.row_scale;
    const float hi = to_tf32(v);
    P.out[i] = hi;
    if (P.planes == 2) P.out[total + i] = to_tf32(v - hi);
  }
}

// t[n'] and s[n'] (one block per output column, fixed-order tree reduction: deterministic); runs after pack_conv_kernel.
__global__ void __launch_bounds__(128) pack_vectors_kernel(const PackParams P) {
  __shared__ float red_s[128], red_t[128];
  const int np = blockIdx.x;
  const int n = P.shuffle_cout ? 4 * (np % P.shuffle_cout) + np / P.shuffle_cout : np;
  const int64_t total = (int64_t)P.N * P.Kp;
  float ps = 0.f, pt = 0.f;
  if (P.gamma) {
    for (int kp = threadIdx.x; kp < P.Kp; kp += 128) {
      ps += P.out[(int64_t)np * P.Kp + kp];
      if (P.planes == 2) ps += P.out[total + (int64_t)np * P.Kp + kp];
    }
    for (int c = threadIdx.x; c < P.Ctot; c += 128) pt += P.beta[c] * P.w[((int64_t)n * P.Ctot + c) * P.taps];   // (taps == 1)
  }
  red_s[threadIdx.x] = ps, red_t[threadIdx.x] = pt;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red_s[threadIdx.x] += red_s[threadIdx.x + o], red_t[threadIdx.x] += red_t[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float t = red_t[0] + (P.bias ? P.bias[n] : 0.f);
    if (np < P.scale_cols) t *= P.row_scale;
    P.out[(int64_t)P.planes * total + np] = t;
    P.out[(int64_t)P.planes * total + P.N + np] = red_s[0];
  }
}

// ---------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    (void)cudaGetLastError();
  });
  return fn;
}

int ceil32(int c) { return (c + 31) & ~31; }

struct Geometry {
  int Ho, Wo, TW, TH, tiles_x, tiles_y;
};

Geometry geometry(int H, int W, int ksize, int stride, bool halo = false) {
  Geometry g;
  if (halo) {   // 3x3 stride 1: 8-wide x 16-tall tiles (one 8-pixel core-matrix group per output row)
    g.Ho = H, g.Wo = W, g.TW = kHaloTW, g.TH = kHaloTH;
    g.tiles_x = (W + kHaloTW - 1) / kHaloTW, g.tiles_y = (H + kHaloTH - 1) / kHaloTH;
    return g;
  }
  const int pad = ksize / 2;
  g.Ho = (H + 2 * pad - ksize) / stride + 1;
  g.Wo = (W + 2 * pad - ksize) / stride + 1;
  if (g.Ho == 1 && g.Wo > 16) {  // token-major rows (a Linear layer = 1x1 "convolution" over an H = 1 image): 128 x 1 tiles
    g.TW = kTileM, g.TH = 1;
    g.tiles_x = (g.Wo + kTileM - 1) / kTileM, g.tiles_y = 1;
    return g;
  }
  int best = 1 << 30;
  g.TW = 16, g.TH = 8;
  const int cand[3] = {16, 32, 8};  // preference order on ties
  for (int i = 0; i < 3; ++i) {
    const int tw = cand[i], th = kTileM / tw;
    const int n = ((g.Wo + tw - 1) / tw) * ((g.Ho + th - 1) / th);
    if (n < best) best = n, g.TW = tw, g.TH = th;
  }
  g.tiles_x = (g.Wo + g.TW - 1) / g.TW;
  g.tiles_y = (g.Ho + g.TH - 1) / g.TH;
  return g;
}

// Column tile.  One tile when N fits (<= 256 columns; 128 in the 3xTF32 mode, whose stages carry a lo plane of both
// operands); several tiles are multiples of the 32-column store chunk (the last one may overhang N: TMA zero-fills the
// weight rows past N and clips the store).  Which width: the one with the lowest estimated time -- waves x k-blocks x
// max(operand bytes per k-block / the ~53 B/clk an SM pulls from L2 (measured), MMA cycles) -- so that a small batch
// (12 pixel tiles for one 768x512 image) spreads over the SMs as narrow column tiles, while a large batch takes the widest
// tile (least A re-reads).  The choice never changes a result bit: every output element is the same fixed-order K loop.
int conv_n_tile(int N, int precise, int gran, int m_tiles, int k_blocks) {
  if (N <= 0 || N % 16) return STF_E_SHAPE;
  if (gran < 32) gran = 32;
  int best = -1;
  double best_t = 0;
  for (int nt = gran; nt <= 256 + gran - 1; nt += gran) {
    int w = nt;
    if (w >= N) w = N;          // single tile: N itself (any multiple of 16)
    if (w > 256) break;
    // the operand ring must hold >= 2 stages (>= 3 for K loops long enough to need the latency hiding) next to the
    // staging slabs and the epilogue vectors (3xTF32 stages carry a lo plane of both operands)
    const size_t stage = (size_t)(precise ? 2 : 1) * (16384 + 128 * (size_t)w);
    const size_t avail = 227 * 1024 - (size_t)kEpiWarps * kSlabBytes - 24 * 1024;
    const int stages = (int)(avail / stage);
    if (stages < 2) {
      if (w == N) break;
      continue;
    }
    const int n_tiles = (N + w - 1) / w;
    const long long tiles = (long long)m_tiles * n_tiles;
    const long long waves = (tiles + kNumSMs - 1) / kNumSMs;
    const double bytes = (precise ? 2.0 : 1.0) * (16384.0 + 128.0 * w);   // L2 -> shared memory per k-block
    const double mma = (precise ? 3.0 : 1.0) * 2.0 * w;                    // tensor-pipe cycles per k-block
    // shared-memory port (128 B/clk): TMA fill + operand reads of every MMA (4 KB of A + 32 B per column, 4 k-steps,
    // x3 in the 3xTF32 mode) + the A pass (read the stage, write hi and lo)
    const double port = (bytes + (precise ? 12.0 : 4.0) * (4096.0 + 32.0 * w) + (precise ? 49152.0 : 0.0)) / 128.0;
    double per_kb = bytes / 53.0 > mma ? bytes / 53.0 : mma;
    if (port > per_kb) per_kb = port;
    if (stages == 2) per_kb *= 1.15;    // two stages hide less of the L2 latency
    const double t = (double)waves * ((double)k_blocks * per_kb + 12.0 * w + 1500.0);
    if (best < 0 || t < best_t * 0.999 || (t <= best_t * 1.001 && w > best)) best = w, best_t = t;
    if (w == N) break;
  }
  return best > 0 ? best : STF_E_SHAPE;
}

int packed_geometry(const stf_conv_args *a, int *Cp, int *Kp) {
  if (a->n_src < 1 || a->n_src > kMaxSrc) return STF_E_ARG;
  int cp = 0;
  for (int s = 0; s < a->n_src; ++s) {
    if (a->src_channels[s] <= 0 || a->src_channels[s] % 4) return STF_E_SHAPE;
    cp += ceil32(a->src_channels[s]);
  }
  *Cp = cp;
  *Kp = cp * a->ksize * a->ksize;
  return STF_OK;
}

}  // namespace

extern "C" int64_t stf_packed_conv_floats(const stf_conv_args *a) {
  int Cp, Kp;
  if (!a || packed_geometry(a, &Cp, &Kp) != STF_OK) return STF_E_ARG;
  return (int64_t)(a->precision == STF_PREC_FP32 ? 2 : 1) * a->N * Kp + 2 * (int64_t)a->N;
}

extern "C" int stf_pack_conv(const stf_conv_args *a, const float *weight, const float *bias, const float *ln_gamma,
                             const float *ln_beta, int scale_cols, float row_scale, float *packed, void *stream) {
  if (!a || !weight || !packed) return STF_E_ARG;
  if (!aligned16(packed)) return STF_E_ALIGN;
  if ((ln_gamma == nullptr) != (ln_beta == nullptr)) return STF_E_ARG;
  if (ln_gamma && (a->ksize != 1 || a->n_src != 1 || !a->has_ln)) return STF_E_ARG;
  if (!ln_gamma && a->has_ln) return STF_E_ARG;
  PackParams P{};
  int rc = packed_geometry(a, &P.Cp, &P.Kp);
  if (rc != STF_OK) return rc;
  if (a->pixel_shuffle != 0 && (a->pixel_shuffle != 2 || a->N % 4)) return STF_E_SHAPE;
  P.w = weight, P.bias = bias, P.gamma = ln_gamma, P.beta = ln_beta, P.out = packed, P.N = a->N;
  P.taps = a->ksize * a->ksize, P.n_src = a->n_src;
  P.Ctot = 0;
  for (int s = 0; s < a->n_src; ++s) P.src_c[s] = a->src_channels[s], P.src_cp[s] = ceil32(a->src_channels[s]), P.Ctot += a->src_channels[s];
  P.shuffle_cout = a->pixel_shuffle ? a->N / 4 : 0;
  P.planes = a->precision == STF_PREC_FP32 ? 2 : 1;
  P.scale_cols = scale_cols, P.row_scale = row_scale;
  const int64_t total = (int64_t)P.N * P.Kp;
  const int blocks = (int)((total + 255) / 256 < 2048 ? (total + 255) / 256 : 2048);
  pack_conv_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(P);
  rc = check_launch();
  if (rc != STF_OK) return rc;
  pack_vectors_kernel<<<P.N, 128, 0, (cudaStream_t)stream>>>(P);
  return check_launch();
}

extern "C" int stf_conv2d(const stf_conv_args *a, void *stream) {
  if (!a || !a->w_packed || !a->y) return STF_E_ARG;
  if (a->batch <= 0 || a->H <= 0 || a->W <= 0 || a->N <= 0) return STF_E_SHAPE;
  if (a->ksize != 1 && a->ksize != 3 && a->ksize != 5) return STF_E_SHAPE;
  if (a->stride != 1 && a->stride != 2) return STF_E_SHAPE;
  if (a->pixel_shuffle != 0 && (a->pixel_shuffle != 2 || a->N % 4 || (a->N / 4) % 16)) return STF_E_SHAPE;
  if (a->ldy % 4 || !aligned16(a->y) || !aligned16(a->w_packed)) return STF_E_ALIGN;
  const int precise = a->precision == STF_PREC_FP32 ? 1 : 0;
  ConvParams P{};
  int Cp, Kp;
  int rc = packed_geometry(a, &Cp, &Kp);
  if (rc != STF_OK) return rc;
  const int shuffle_cout = a->pixel_shuffle ? a->N / 4 : 0;
  const int cw = (!shuffle_cout || shuffle_cout % 32 == 0) ? 32 : 16;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return STF_E_ARG;
  // halo mode: every 3x3 stride-1 convolution (STF_B200_CONV_HALO=0 keeps the tap-by-tap loads: A/B measurements)
  static const bool halo_on = !(getenv("STF_B200_CONV_HALO") && atoi(getenv("STF_B200_CONV_HALO")) == 0);
  // (single-pass mode only: in the 3xTF32 parity mode two halo tiles with their lo planes leave no room for the weight ring)
  const bool halo = halo_on && a->ksize == 3 && a->stride == 1 && !a->has_ln && !precise;
  const Geometry g = geometry(a->H, a->W, a->ksize, a->stride, halo);
  const int s = a->stride;
  const int n_tile = conv_n_tile(a->N, precise, cw, g.tiles_x * g.tiles_y * a->batch, (Cp / kBlockK) * a->ksize * a->ksize);
  if (n_tile < 0) return n_tile;
  if (shuffle_cout && n_tile % cw) return STF_E_SHAPE;

  P.n_src = a->n_src;
  P.kb_per_tap = 0;
  for (int i = 0; i < a->n_src; ++i) {
    if (!a->src[i] || !aligned16(a->src[i]) || a->src_ld[i] % 4 || a->src_ld[i] < a->src_channels[i]) return STF_E_ALIGN;
    P.src_kb[i] = ceil32(a->src_channels[i]) / kBlockK;
    P.kb_per_tap += P.src_kb[i];
    const cuuint64_t gdim[4] = {(cuuint64_t)a->src_channels[i], (cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->batch};
    const cuuint64_t gstr[3] = {(cuuint64_t)a->src_ld[i] * 4, (cuuint64_t)a->W * a->src_ld[i] * 4,
                                (cuuint64_t)a->H * a->W * a->src_ld[i] * 4};
    const cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)(halo ? kHaloW : g.TW * s),
                               (cuuint32_t)(halo ? kHaloTH + 2 : g.TH * s), 1};
    const cuuint32_t estr[4] = {1, (cuuint32_t)s, (cuuint32_t)s, 1};
    if (enc(&P.a_map[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(a->src[i]), gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return STF_E_SHAPE;
  }
  P.ksize = a->ksize, P.pad = a->ksize / 2, P.stride = s;
  P.k_blocks = P.kb_per_tap * a->ksize * a->ksize;
  P.tail_ksteps = 4;
  if (a->ksize == 1 && a->n_src == 1 && a->src_channels[0] % kBlockK) P.tail_ksteps = (a->src_channels[0] % kBlockK + 7) / 8;
  for (int p = 0; p < 1 + precise; ++p) {
    const cuuint64_t gdim[2] = {(cuuint64_t)Kp, (cuuint64_t)a->N};
    const cuuint64_t gstr[1] = {(cuuint64_t)Kp * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)n_tile};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&P.b_map[p], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(a->w_packed) + (size_t)p * a->N * Kp, gdim,
            gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return STF_E_SHAPE;
  }
  P.bias = a->w_packed + (size_t)(1 + precise) * a->N * Kp;
  P.svec = P.bias + a->N;
  P.has_ln = a->has_ln ? 1 : 0, P.ln_eps = a->ln_eps, P.ln_k = a->src_channels[0];
  if (P.has_ln && (a->ksize != 1 || a->n_src != 1)) return STF_E_ARG;
  P.shuffle_cout = shuffle_cout;
  P.cw = cw;
  const int n_maps = P.shuffle_cout ? 4 : 1;
  for (int m = 0; m < n_maps; ++m) {
    cuuint64_t gdim[4], gstr[3];
    float *base = a->y;
    if (P.shuffle_cout) {
      const int i = m >> 1, j = m & 1;
      const int64_t Wo2 = 2 * (int64_t)g.Wo, Ho2 = 2 * (int64_t)g.Ho;
      base += ((int64_t)i * Wo2 + j) * a->ldy;
      gdim[0] = (cuuint64_t)P.shuffle_cout, gdim[1] = (cuuint64_t)g.Wo, gdim[2] = (cuuint64_t)g.Ho, gdim[3] = (cuuint64_t)a->batch;
      gstr[0] = (cuuint64_t)2 * a->ldy * 4, gstr[1] = (cuuint64_t)2 * Wo2 * a->ldy * 4, gstr[2] = (cuuint64_t)Ho2 * Wo2 * a->ldy * 4;
    } else {
      gdim[0] = (cuuint64_t)a->N, gdim[1] = (cuuint64_t)g.Wo, gdim[2] = (cuuint64_t)g.Ho, gdim[3] = (cuuint64_t)a->batch;
      gstr[0] = (cuuint64_t)a->ldy * 4, gstr[1] = (cuuint64_t)g.Wo * a->ldy * 4, gstr[2] = (cuuint64_t)g.Ho * g.Wo * a->ldy * 4;
    }
    const int bw = g.TW < 32 ? g.TW : 32;   // one epilogue warp's 32 rows of the tile
    const cuuint32_t box[4] = {(cuuint32_t)P.cw, (cuuint32_t)bw, (cuuint32_t)(32 / bw), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (enc(&P.y_map[m], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            P.cw == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return STF_E_SHAPE;
  }
  P.N = a->N, P.n_tile = n_tile, P.n_tiles = (a->N + n_tile - 1) / n_tile;
  P.TW = g.TW, P.TH = g.TH, P.tiles_x = g.tiles_x, P.tiles_y = g.tiles_y;
  P.tiles_per_img = g.tiles_x * g.tiles_y;
  P.m_tiles = P.tiles_per_img * a->batch;
  P.total_tiles = P.m_tiles * P.n_tiles;
  P.act = a->act;
  if (a->act < 0 || a->act > 3) return STF_E_ARG;
  if (a->act >= 2) {
    if (!a->residual || a->pixel_shuffle || a->res_ld % 4 || !aligned16(a->residual)) return STF_E_ARG;
    P.residual = a->residual, P.res_ld = a->res_ld;
  }
  P.Ho = g.Ho, P.Wo = g.Wo;
  P.idesc = umma_idesc_tf32(kTileM, n_tile);
  P.acc_stride = n_tile;
  int cols = 32;
  while (cols < 2 * n_tile) cols <<= 1;
  P.tmem_cols = cols;
  P.b_plane_bytes = (uint32_t)n_tile * 128u;
  P.stage_bytes = halo ? (uint32_t)(1 + precise) * kHaloBytes : (uint32_t)(1 + precise) * (kAStageBytes + P.b_plane_bytes);
  // shared memory: operand ring first (as many stages as fit next to the minimum of 2 staging slabs per epilogue warp; the
  // ring runs across tile boundaries, so short K loops prefetch the next tiles), what is left goes to more staging slabs
  P.n_pad = ((P.n_tiles * n_tile + 31) & ~31) + 32;
  const size_t small = 1024 /*alignment slack*/ + (size_t)P.n_pad * 4 * (P.has_ln ? 2 : 1) +
                       (P.has_ln ? (size_t)kStatSlots * kTileM * 8 : 0) + (5 * kMaxStages + 4) * 8 + 16;
  const size_t cap = 227 * 1024;
  int stages, b_stages = 0;
  size_t ring_bytes;
  if (halo) {
    // two halo tiles (the next channel block lands while the nine taps of this one run), the rest of the ring memory in
    // weight tiles (one per tap: at least 4 so that the producer runs half a channel block ahead of the tensor core)
    const size_t b_stage = (size_t)(1 + precise) * P.b_plane_bytes;
    stages = 2;
    size_t left = cap - small - kEpiWarps * kSlabBytes - (size_t)stages * P.stage_bytes;
    b_stages = (int)(left / b_stage);
    if (b_stages > kMaxStages) b_stages = kMaxStages;
    if (b_stages < 3) return STF_E_SHAPE;
    if (b_stages >= 6 && (size_t)(stages + 1) * P.stage_bytes + 4 * b_stage + small + kEpiWarps * kSlabBytes <= cap)
      stages = 3, b_stages = (int)((cap - small - kEpiWarps * kSlabBytes - 3 * (size_t)P.stage_bytes) / b_stage);
    if (b_stages > kMaxStages) b_stages = kMaxStages;
    ring_bytes = (size_t)stages * P.stage_bytes + (size_t)b_stages * b_stage;
  } else {
    stages = (int)((cap - small - kEpiWarps * 1 * kSlabBytes) / P.stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) return STF_E_SHAPE;
    ring_bytes = (size_t)stages * P.stage_bytes;
  }
  P.stages = stages, P.b_stages = b_stages;
  int slabs = (int)((cap - small - ring_bytes) / (kEpiWarps * kSlabBytes));
  if (slabs > kMaxSlabs) slabs = kMaxSlabs;
  P.slabs = slabs;
  static const int dbg = getenv("STF_B200_CONV_DEBUG") ? atoi(getenv("STF_B200_CONV_DEBUG")) : 0;
  P.debug = dbg;
  const size_t fixed = small + (size_t)kEpiWarps * slabs * kSlabBytes;
  const size_t smem = fixed + ring_bytes;
  const int sms = a->max_ctas > 0 && a->max_ctas < kNumSMs ? a->max_ctas : kNumSMs;
  const int grid = P.total_tiles < sms ? P.total_tiles : sms;
  auto kern = halo ? (precise ? conv_tf32_kernel<1, 0, 1> : conv_tf32_kernel<0, 0, 1>)
                   : precise ? (P.has_ln ? conv_tf32_kernel<1, 1, 0> : conv_tf32_kernel<1, 0, 0>)
                             : (P.has_ln ? conv_tf32_kernel<0, 1, 0> : conv_tf32_kernel<0, 0, 0>);
  static std::once_flag attr_once[6];
  std::call_once(attr_once[halo ? 4 + precise : precise * 2 + P.has_ln], [&] {
    (void)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap);
  });
  kern<<<grid, kThreads, smem, (cudaStream_t)stream>>>(P);
  return check_launch();
}

extern "C" int stf_conv2d_out_hw(int H, int W, int ksize, int stride, int *Ho, int *Wo) {
  if (H <= 0 || W <= 0 || (ksize != 1 && ksize != 3 && ksize != 5) || (stride != 1 && stride != 2)) return STF_E_SHAPE;
  const Geometry g = geometry(H, W, ksize, stride);
  if (Ho) *Ho = g.Ho;
  if (Wo) *Wo = g.Wo;
  return STF_OK;
}

}  // namespace stf
