// Device-side rANS *decoder* for the slice loop of decompress():  RansDecoder.decode_stream
// (compressai/cpp_exts/rans/rans_interface.cpp:285-350, rans64.h:107-142) for up to 32 independent streams per warp.
//
// Why on the device although rANS is a sequential integer state machine: decompress() needs the symbols of slice i-1 before
// it can compute the parameters of slice i (stf.py:757-779), so the host decoder costs two PCIe hops and a host
// synchronisation per slice (36 per batch of 64 images), and with eight ranks sharing one host (4 threads per rank) the
// slice loop becomes host-bound.  One image is one stream (the reference's format), but the images of a sub-batch are
// independent: lane b of a warp decodes image b, all lanes in lockstep -- the SIMT shape of the host coder's W-way
// interleave.  A sub-batch then decodes with NO host round trip (one CUDA graph), and the single decoding warp runs
// concurrently with the other sub-batches' convolutions on the rest of the GPU.
// Same integers as the host decoder (csrc/rans_host.cpp): symbols are bit-identical (tests/test_gpu_entropy.py).
//
// Table image (built on the host by stf_rans_device_table_pack, kept in shared memory by the kernel):
//   header  {rows, cdf_entries, lut_off, cdf_off} int32 x 4
//   rowinfo [rows] {int32 offset, int32 escape, uint32 cbase, uint32 pad}
//   lut     [rows][1024] uint16  symbol holding cumulative value (bucket << 6): at most ~3 symbols per bucket for the widest row
//                                (3133 symbols), so the search behind the LUT is one three-entry probe (the host coder's 256
//                                buckets + linear scan cost ~12 dependent shared-memory loads there, x the slowest lane)
//   cdf     [cdf_entries] uint16 cumulative counts, 65536 stored as 0 (only ever the last entry of a row)
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include <vector>

#include "common.cuh"

struct stf_rans_table;   // csrc/rans_host.cpp
extern "C" int stf_rans_table_export(const stf_rans_table *t, int *rows, const int32_t **sizes, const int32_t **offsets,
                                     const uint32_t **cdf, const uint32_t **cbase);

namespace stf {
namespace {

constexpr uint64_t kLow = 1ull << 31;
constexpr int kDevLutBits = 10;

struct DecArgs {
  const uint8_t *table;     // device copy of the table image
  int table_bytes;
  const uint32_t *streams;  // all streams of the sub-batch back to back (32-bit words)
  const int64_t *offsets;   // [count] first word of stream b
  const int32_t *lengths;   // [count] words of stream b
  uint64_t *state_x;        // [count] decoder state between slices
  uint32_t *state_pos;      // [count] next unread word
  int32_t *status;          // [count] 0 ok, STF_E_STREAM / STF_E_ARG (sticky)
  const int32_t *indexes;   // slice indexes in coding order, image b at indexes + b * idx_stride
  int64_t idx_stride;
  int32_t *symbols;         // decoded symbols, image b at symbols + b * sym_stride
  int64_t sym_stride;
  int count, first, lanes;
  int64_t n;
};

constexpr int kTile = 32;     // symbols per stream between two staging points
constexpr int kPitch = 33;    // row pitch of the index / symbol tiles: lane l walks row l, bank (l + j) % 32
constexpr int kRing = 64;     // stream words kept in shared memory per lane (word-interleaved: bank = lane)

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// One decoding warp per CTA, `lanes` streams per warp (lane l = stream blockIdx.x * lanes + l).  Everything the state
// machine touches per symbol is in shared memory:
//   * indexes: tile t + 1 arrives by cp.async (coalesced: the whole warp fetches 32 consecutive indexes of one stream per
//     instruction) while tile t is decoded; symbols leave the same way round, 32 per stream per store;
//   * stream words: a 64-word ring per lane, topped up by cp.async once per tile, one tile ahead of its use (a tile consumes
//     at most 16 words unless it holds escapes; past the ring the lane reads global memory directly);
//   * the index -> row-info fetch runs two / one symbols ahead of the state chain
//       state -> LUT -> 4 CDF entries -> select -> multiply -> refill (branch-free)  ~ 120 cycles.
// Round-2 first version (per-lane global index loads and symbol stores inside the chain) measured 514-754 cycles per step.
__global__ void __launch_bounds__(128) rans_decode_kernel(const DecArgs a) {
  extern __shared__ __align__(16) uint8_t tbl[];
  for (int i = threadIdx.x; i < a.table_bytes / 16; i += blockDim.x)
    reinterpret_cast<int4 *>(tbl)[i] = __ldg(reinterpret_cast<const int4 *>(a.table) + i);
  __syncthreads();
  if (threadIdx.x >= 32) return;   // the helper warps only load the table
  const int lane = threadIdx.x;
  const int b0 = blockIdx.x * a.lanes;
  const int L = min(a.lanes, a.count - b0);
  const bool active = lane < L;
  const int b = b0 + (active ? lane : 0);
  const int32_t *hdr = reinterpret_cast<const int32_t *>(tbl);
  const uint32_t rows = (uint32_t)hdr[0];
  const int4 *rowinfo = reinterpret_cast<const int4 *>(tbl + 16);
  const uint16_t *lut = reinterpret_cast<const uint16_t *>(tbl + hdr[2]);
  const uint16_t *cdf = reinterpret_cast<const uint16_t *>(tbl + hdr[3]);
  int32_t *idx_s = reinterpret_cast<int32_t *>(tbl + a.table_bytes);          // [2][lanes][kPitch] (+ 2 pad words)
  int32_t *sym_s = idx_s + 2 * a.lanes * kPitch + 4;                          // [lanes][kPitch]
  uint32_t *ring = reinterpret_cast<uint32_t *>(sym_s + a.lanes * kPitch);    // [kRing][32]

  const uint32_t *w = a.streams + a.offsets[b];
  const uint32_t nwords = (uint32_t)a.lengths[b];
  uint64_t x = 0;
  uint32_t pos = 2;
  int st = a.first ? 0 : a.status[b];
  if (a.first) {
    if (nwords < 2) st = STF_E_STREAM;
    else x = (uint64_t)w[0] | ((uint64_t)w[1] << 32);                          // Rans64DecInit
  } else {
    x = a.state_x[b], pos = a.state_pos[b];
  }
  if (!active) st = 1;                       // idle lanes only help with the staging
  uint32_t ring_end = pos;                   // words [pos, ring_usable) are readable, [ring_usable, ring_end) in flight
  uint32_t ring_usable = pos;

  const int64_t ntiles = (a.n + kTile - 1) / kTile;
  auto stage_tile = [&](int64_t t) {         // indexes of tile t (all streams) + this lane's next stream words
    if (t < ntiles) {
      int32_t *dst = idx_s + (t & 1) * a.lanes * kPitch;
      const int64_t i = t * kTile + lane;
      if (i < a.n)
        for (int l = 0; l < L; ++l) cp_async4(dst + l * kPitch + lane, a.indexes + (int64_t)(b0 + l) * a.idx_stride + i);
    }
    if (st == 0) {
      const uint32_t want = min(nwords, pos + (uint32_t)kRing);
      for (uint32_t k = max(ring_end, pos); k < want; ++k) cp_async4(ring + (k & (kRing - 1)) * 32 + lane, w + k);
      ring_end = max(ring_end, want);
    }
    cp_async_commit();
  };

  stage_tile(0);
  for (int64_t t = 0; t < ntiles; ++t) {
    cp_async_wait_all();
    ring_usable = ring_end;
    __syncwarp();
    stage_tile(t + 1);
    const int32_t *it = idx_s + (t & 1) * a.lanes * kPitch + lane * kPitch;
    int32_t *ot = sym_s + lane * kPitch;
    const int steps = (int)min((int64_t)kTile, a.n - t * kTile);
    if (st == 0) {
      uint32_t row = (uint32_t)it[0];
      if (row >= rows) st = STF_E_ARG, row = 0;
      int4 ri = rowinfo[row];
      uint32_t row_n = (uint32_t)it[1];      // (pad words keep it[steps], it[steps + 1] inside the buffer)
      for (int j = 0; j < steps && st == 0; ++j) {
        const uint32_t row_nn = (uint32_t)it[j + 2];
        const int4 ri_n = rowinfo[min(row_n, rows - 1)];
        const uint32_t wv = pos < ring_usable ? ring[(pos & (kRing - 1)) * 32 + lane] : (pos < nwords ? w[pos] : 0u);
        const uint16_t *c = cdf + (uint32_t)ri.z;
        const uint32_t cum = (uint32_t)x & 0xFFFFu;
        uint32_t s = lut[(row << kDevLutBits) + (cum >> (16 - kDevLutBits))];
        uint32_t start, hi;
        for (;;) {                           // c[s] <= cum by construction; find s with c[s + 1] > cum
          const uint32_t c0 = c[s];
          uint32_t c1 = c[s + 1], c2 = c[s + 2], c3 = c[s + 3];   // (reads past the row end stay inside the image: padded)
          c1 = c1 ? c1 : 65536u, c2 = c2 ? c2 : 65536u, c3 = c3 ? c3 : 65536u;
          const bool p1 = c1 <= cum, p2 = p1 && c2 <= cum, p3 = p2 && c3 <= cum;
          if (p3) {
            s += 3;
            continue;
          }
          start = p2 ? c2 : (p1 ? c1 : c0);
          hi = p2 ? c3 : (p1 ? c2 : c1);
          s += (uint32_t)p1 + (uint32_t)p2;
          break;
        }
        x = (uint64_t)(hi - start) * (x >> 16) + (cum - start);    // Rans64DecAdvance
        const bool need = x < kLow;                                // Rans64DecRenorm, branch-free
        if (need && pos >= nwords) st = STF_E_STREAM;
        x = need ? ((x << 32) | wv) : x;
        pos += need ? 1u : 0u;
        int32_t v = (int32_t)s;
        if (v == ri.y) {                     // escape: nibble-coded bypass value (rans_interface.cpp:320-343); rare
          auto nibble = [&]() -> int {
            const int nib = (int)(x & 15u);
            x >>= 4;
            if (x < kLow) {
              if (pos >= nwords) {
                st = STF_E_STREAM;
              } else {
                const uint32_t wn = pos < ring_usable ? ring[(pos & (kRing - 1)) * 32 + lane] : w[pos];
                x = (x << 32) | wn;
                ++pos;
              }
            }
            return nib;
          };
          int nib = nibble(), nn = nib;
          while (nib == 15 && st == 0) {
            nib = nibble();
            nn += nib;
          }
          int32_t raw = 0;
          for (int k = 0; k < nn && st == 0; ++k) {
            nib = nibble();
            if (k < 8) raw |= nib << (k * 4);
          }
          v = raw >> 1;
          v = (raw & 1) ? -v - 1 : v + ri.y;
        }
        ot[j] = v + ri.x;
        if (row_n >= rows && j + 1 < steps) st = STF_E_ARG;
        row = min(row_n, rows - 1), ri = ri_n, row_n = row_nn;
      }
    }
    __syncwarp();
    const int64_t i = t * kTile + lane;
    if (i < a.n)
      for (int l = 0; l < L; ++l) a.symbols[(int64_t)(b0 + l) * a.sym_stride + i] = sym_s[l * kPitch + lane];
    __syncwarp();
  }
  cp_async_wait_all();
  if (active) a.state_x[b] = x, a.state_pos[b] = pos, a.status[b] = st;
}

}  // namespace
}  // namespace stf

using namespace stf;

extern "C" int64_t stf_rans_device_table_bytes(const stf_rans_table *t) {
  int rows;
  const int32_t *sizes, *offsets;
  const uint32_t *cdf, *cbase;
  if (!t || stf_rans_table_export(t, &rows, &sizes, &offsets, &cdf, &cbase) != STF_OK) return STF_E_ARG;
  int64_t entries = 0;
  for (int r = 0; r < rows; ++r) entries += sizes[r];
  int64_t bytes = 16 + (int64_t)rows * 16 + (int64_t)rows * (1 << kDevLutBits) * 2 + (entries + 8) * 2;
  return (bytes + 15) & ~(int64_t)15;
}

extern "C" int stf_rans_device_table_pack(const stf_rans_table *t, void *host_out) {
  int rows;
  const int32_t *sizes, *offsets;
  const uint32_t *cdf, *cbase;
  if (!t || !host_out || stf_rans_table_export(t, &rows, &sizes, &offsets, &cdf, &cbase) != STF_OK) return STF_E_ARG;
  const int64_t bytes = stf_rans_device_table_bytes(t);
  uint8_t *o = static_cast<uint8_t *>(host_out);
  memset(o, 0, (size_t)bytes);
  int64_t entries = 0;
  for (int r = 0; r < rows; ++r) entries += sizes[r];
  int32_t *hdr = reinterpret_cast<int32_t *>(o);
  hdr[0] = rows, hdr[1] = (int32_t)entries, hdr[2] = 16 + rows * 16, hdr[3] = hdr[2] + rows * (1 << kDevLutBits) * 2;
  int32_t *ri = reinterpret_cast<int32_t *>(o + 16);
  uint16_t *lut = reinterpret_cast<uint16_t *>(o + hdr[2]);
  uint16_t *c16 = reinterpret_cast<uint16_t *>(o + hdr[3]);
  uint32_t run = 0;
  for (int r = 0; r < rows; ++r) {
    const uint32_t *c = cdf + cbase[r];
    ri[4 * r] = offsets[r], ri[4 * r + 1] = sizes[r] - 2, ri[4 * r + 2] = (int32_t)run, ri[4 * r + 3] = 0;
    for (int j = 0; j < sizes[r]; ++j) c16[run + j] = (uint16_t)(c[j] & 0xFFFFu);   // 65536 -> 0
    uint32_t s = 0;
    for (uint32_t bkt = 0; bkt < (1u << kDevLutBits); ++bkt) {
      const uint32_t cum = bkt << (16 - kDevLutBits);
      while (c[s + 1] <= cum) ++s;
      lut[(r << kDevLutBits) + bkt] = (uint16_t)s;
    }
    run += (uint32_t)sizes[r];
  }
  return STF_OK;
}

extern "C" int stf_rans_decode_device(const void *table_dev, int64_t table_bytes, const uint32_t *streams,
                                      const int64_t *stream_offsets, const int32_t *stream_words, uint64_t *state_x,
                                      uint32_t *state_pos, int32_t *status, int first, const int32_t *indexes,
                                      int64_t idx_batch_stride, int32_t *symbols_out, int64_t sym_batch_stride, int count,
                                      int64_t n, void *stream) {
  if (!table_dev || !streams || !stream_offsets || !stream_words || !state_x || !state_pos || !status || count < 0 || n < 0)
    return STF_E_ARG;
  if (n > 0 && (!indexes || !symbols_out)) return STF_E_ARG;
  if (table_bytes <= 0 || table_bytes % 16 || table_bytes > 200 * 1024 || !aligned16(table_dev)) return STF_E_TABLE;
  if (count == 0) return STF_OK;
  DecArgs a{};
  a.table = static_cast<const uint8_t *>(table_dev), a.table_bytes = (int)table_bytes;
  a.streams = streams, a.offsets = stream_offsets, a.lengths = stream_words;
  a.state_x = state_x, a.state_pos = state_pos, a.status = status;
  a.indexes = indexes, a.idx_stride = idx_batch_stride, a.symbols = symbols_out, a.sym_stride = sym_batch_stride;
  a.count = count, a.first = first, a.n = n;
  // Streams per decoding warp: fewer lanes = less divergence per step and more SMs decoding side by side (one CTA per
  // `lanes` streams; a CTA holds the ~185 KB table image, so it owns its SM).  STF_B200_RANS_DEVICE_LANES, default 8.
  static const int lanes_env = [] {
    const char *e = getenv("STF_B200_RANS_DEVICE_LANES");
    const int v = e ? atoi(e) : 8;
    return v >= 1 && v <= 32 ? v : 8;
  }();
  a.lanes = lanes_env;
  const size_t smem = (size_t)table_bytes + ((size_t)3 * a.lanes * kPitch + 4) * 4 + (size_t)kRing * 32 * 4;
  if (smem > 227 * 1024) return STF_E_TABLE;
  static std::atomic<int> attr_set{0};
  if (!attr_set.load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(rans_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set.store(1, std::memory_order_release);
  }
  rans_decode_kernel<<<(count + a.lanes - 1) / a.lanes, 128, smem, (cudaStream_t)stream>>>(a);
  return check_launch();
}
