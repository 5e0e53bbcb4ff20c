// Host-side rANS codec of libstf_b200 (CPU by design: rANS is a sequential integer state machine
// and stays on the host in the reference too -- SURVEY.md section 2.1 / 8(f) rank 1).
//
// Bit-exact replacement of compressai.ans:
//   BufferedRansEncoder / RansEncoder   compressai/cpp_exts/rans/rans_interface.cpp:99-204
//   RansDecoder                         compressai/cpp_exts/rans/rans_interface.cpp:206-350
//   rANS64 state machine                third_party/ryg_rans/rans64.h:59-142
//   pmf_to_quantized_cdf                compressai/cpp_exts/ops/ops.cpp:24-81
//
// What differs from the reference (the stream does not):
//   - int32 buffers in, bytes out: no Python lists, no per-call conversion of the 64x3133 table;
//   - the encoder walks the symbols backwards and emits directly (no staged symbol vector);
//   - x / freq is an exact multiply-shift with a per-symbol 64-bit reciprocal;
//   - the decoder finds the symbol through a bucket LUT per CDF row (256 or 4096 buckets) instead of a linear scan;
//   - renormalisation is branch-free on high-entropy runs (chosen per run, see kLutBitsWide);
//   - independent streams (one per image) are coded on separate host threads.
#include <stdlib.h>
#include <string.h>

#include <cmath>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "../../include/stf_b200.h"

namespace {

constexpr uint32_t kProbBits = 16;
constexpr uint32_t kProbScale = 1u << kProbBits;
constexpr uint32_t kNibbleBits = 4;
constexpr int32_t kNibbleMax = 15;
constexpr uint64_t kLow = 1ull << 31;  // RANS64_L
// Decoder LUT resolution per CDF row: 256 buckets for rows of up to 256 symbols, 4096 for wider ones (16 cumulative counts
// per bucket: the widest row of the scale table, 3133 symbols, then has < 1 symbol per bucket; with 256 buckets the scan behind
// the LUT took ~6 dependent steps there and the decoder ran at 23 ns per symbol instead of 6).  Wide rows are also the
// high-entropy ones (~log2(4.1 sigma) bits per symbol), where "does the state renormalise" is a coin flip: they take the
// branch-free step, narrow rows the branchy one (well predicted there, and speculation past a predicted branch is cheaper
// than a select in the dependency chain: 2.8 vs 4.2 ns per symbol at 0.3 bits per symbol).
constexpr int kLutBitsNarrow = 8, kLutBitsWide = 12;
constexpr int kWideRow = 40;   // symbols: sigma ~ 3, ~3.6 bits per symbol (below it the branchy decoder wins, at it they tie)
// images one thread codes in lockstep (env STF_B200_RANS_LOCKSTEP, 1..8)
static const int kMaxLockstep = [] {
  const char *e = getenv("STF_B200_RANS_LOCKSTEP");
  int v = e ? atoi(e) : 4;
  return v < 1 ? 1 : (v > 8 ? 8 : v);
}();

struct EncEntry {    // 16 bytes: the widest row (3133 symbols) is 50 KB, about one L1
  uint64_t rcp;       // freq >= 2: ceil(2^(63+shift) / freq), so that x / freq == mulhi(x, rcp) >> shift exactly;
                      // freq == 1: 2^64 - 1 with shift 0, so that the "quotient" is x - 1 (made up for by `bias`)
  uint32_t bias;      // start (freq >= 2);  start + 65535 (freq == 1):  x + bias + (x - 1) * 65535 == x * 65536 + start
  uint16_t cmpl;      // 65536 - freq (freq >= 1); the renormalisation bound is x >= 2^47 * freq <=> (x >> 47) >= 65536 - cmpl
  uint8_t shift;      // ceil(log2 freq) - 1
  uint8_t pad;
};
static_assert(sizeof(EncEntry) == 16, "EncEntry layout");

}  // namespace

struct stf_rans_table {
  int rows = 0;
  std::vector<int32_t> sizes, offsets;
  std::vector<uint32_t> base;   // first entry of each row in enc / cdf
  std::vector<EncEntry> enc;    // one per (row, symbol)
  std::vector<uint32_t> cdf;    // flattened valid CDF entries, row r: cdf[base[r] + r .. + sizes[r])
  std::vector<uint32_t> cbase;  // first cdf entry of each row
  // symbol holding cumulative value (bucket << (16 - bits)), per row: 256 buckets for the branchy step (5 KB per ten rows:
  // stays in L1 when a run uses a handful of narrow rows), 4096 for the branch-free one (see kLutBitsWide)
  std::vector<uint16_t> lut_narrow, lut_wide;
  struct Row {                  // everything a coding step needs to know about a CDF row, one 16-byte load
    int32_t offset, escape;     // symbol offset; escape symbol = size - 2
    uint32_t base;              // first enc entry of the row; its first cdf entry is base + row
    uint32_t wide;              // 1: more than kWideRow symbols
  };
  std::vector<Row> row;
};

struct stf_rans_decoder {
  uint64_t x = 0;
  const uint32_t *w = nullptr, *end = nullptr;
  std::vector<uint32_t> words;
};

namespace {

inline void make_entry(EncEntry *e, uint32_t start, uint32_t freq) {
  e->cmpl = (uint16_t)(kProbScale - freq);
  e->pad = 0;
  if (freq < 2) {
    e->rcp = ~0ull;
    e->bias = start + kProbScale - 1;
    e->shift = 0;
    return;
  }
  uint32_t sh = 0;
  while (freq > (1u << sh)) ++sh;  // sh = ceil(log2 freq) >= 1
  unsigned __int128 num = ((unsigned __int128)1 << (63 + sh)) + (freq - 1);
  e->rcp = (uint64_t)(num / freq);
  e->bias = start;
  e->shift = (uint8_t)(sh - 1);
}

// x <- C(s, x) for a modelled symbol; identical state sequence to Rans64EncPut (rans64.h:77-93).
// Branch-free: whether a state renormalises (about every third symbol at 10 bits per symbol) is not predictable, and a
// mispredicted branch costs more than the whole step.  The word is stored unconditionally below the write pointer, which
// only moves when the store was real (the buffer bound, 8 bytes per symbol, always leaves that word free).
template <bool kBranchFree>
inline void put_symbol(uint64_t &x, uint32_t *&w, const EncEntry &e) {
  if (!kBranchFree) {
    if ((x >> 47) >= kProbScale - (uint32_t)e.cmpl) {
      *--w = (uint32_t)x;
      x >>= 32;
    }
    const uint64_t q = (uint64_t)(((unsigned __int128)x * e.rcp) >> 64) >> e.shift;
    x = x + e.bias + q * e.cmpl;
    return;
  }
  const uint64_t need = (uint64_t)((x >> 47) >= kProbScale - (uint32_t)e.cmpl);  // x >= ((L >> 16) << 32) * freq
  w[-1] = (uint32_t)x;
  w -= need;
  x >>= (need << 5);
  const uint64_t q = (uint64_t)(((unsigned __int128)x * e.rcp) >> 64) >> e.shift;
  x = x + e.bias + q * e.cmpl;
}

// Raw 4-bit value (Rans64EncPutBits, rans_interface.cpp:59-77): freq = 2^12, x_max = 2^59.
inline void put_nibble(uint64_t &x, uint32_t *&w, uint32_t val) {
  if (x >= (1ull << 59)) {
    *--w = (uint32_t)x;
    x >>= 32;
  }
  x = (x << kNibbleBits) | val;
}

// One symbol of one stream (returns false on an out-of-range index).
template <bool kBranchFree>
inline bool encode_step(const stf_rans_table *t, int32_t sym, int32_t row, uint64_t &x, uint32_t *&w) {
  if ((uint32_t)row >= (uint32_t)t->rows) return false;
  const stf_rans_table::Row ri = t->row[row];
  const int32_t escape = ri.escape;
  int32_t v = sym - ri.offset;
  if ((uint32_t)v < (uint32_t)escape) {
    put_symbol<kBranchFree>(x, w, t->enc[ri.base + v]);
    return true;
  }
  // escape: staged order is [escape symbol][count nibbles][value nibbles]; emit it reversed
  uint32_t raw = v < 0 ? (uint32_t)(-2 * v - 1) : (uint32_t)(2 * (v - escape));
  int32_t nn = 0;
  while ((raw >> (nn * kNibbleBits)) != 0) ++nn;
  for (int32_t j = nn - 1; j >= 0; --j) put_nibble(x, w, (raw >> (j * kNibbleBits)) & kNibbleMax);
  int32_t full = nn / kNibbleMax, rest = nn % kNibbleMax;  // count = 15,15,...,rest
  put_nibble(x, w, (uint32_t)rest);
  for (int32_t j = 0; j < full; ++j) put_nibble(x, w, kNibbleMax);
  put_symbol<false>(x, w, t->enc[ri.base + escape]);
  return true;
}

inline int64_t encode_finish(uint64_t x, uint32_t *w, uint32_t *buf_end) {
  *--w = (uint32_t)(x >> 32);  // Rans64EncFlush: low word first in memory
  *--w = (uint32_t)x;
  return (int64_t)(buf_end - w) * 4;
}

// Does a run of n symbols mostly use wide (high-entropy) rows?  64 evenly spaced samples.
template <class IdxT>
inline bool wide_run(const stf_rans_table *t, const IdxT *indexes, int64_t n) {
  if (n <= 0) return false;
  const int64_t stride = n / 64 > 0 ? n / 64 : 1;
  int wide = 0, seen = 0;
  for (int64_t i = 0; i < n && seen < 64; i += stride, ++seen) {
    const int32_t row = (int32_t)indexes[i];
    if ((uint32_t)row < (uint32_t)t->rows) wide += (int)t->row[row].wide;
  }
  return 4 * wide >= seen;
}

template <bool kBranchFree, class SymT, class IdxT>
__attribute__((noinline)) int64_t encode_into_impl(const stf_rans_table *t, const SymT *symbols, const IdxT *indexes, int64_t n,
                         uint32_t *buf_end) {
  uint32_t *w = buf_end;
  uint64_t x = kLow;
  for (int64_t i = n - 1; i >= 0; --i)
    if (!encode_step<kBranchFree>(t, (int32_t)symbols[i], (int32_t)indexes[i], x, w)) return STF_E_ARG;
  return encode_finish(x, w, buf_end);
}

template <class SymT, class IdxT>
int64_t encode_into(const stf_rans_table *t, const SymT *symbols, const IdxT *indexes, int64_t n, uint32_t *buf_end) {
  return wide_run(t, indexes, n) ? encode_into_impl<true>(t, symbols, indexes, n, buf_end)
                                 : encode_into_impl<false>(t, symbols, indexes, n, buf_end);
}

// W independent streams in lockstep: a stream is one serial dependency chain (state -> multiply-high -> state, ~15 cycles
// per symbol), so a thread that owns W images interleaves them and the core overlaps the W chains (W = 2 or 4; with eight
// ranks sharing a host each rank has a handful of threads for 21+ images).  Same bytes as W encode_into calls.
template <int W, bool kBranchFree, class SymT, class IdxT>
__attribute__((noinline)) void encode_intoW_impl(const stf_rans_table *t, const SymT *const *sym, const IdxT *const *idx, const int64_t *n,
                       uint32_t *const *buf_end, int64_t *nb) {
  uint32_t *w[W];
  uint64_t x[W];
  int64_t i[W];
  int64_t common = n[0];
  for (int k = 0; k < W; ++k) {
    w[k] = buf_end[k], x[k] = kLow, i[k] = n[k] - 1;
    if (n[k] < common) common = n[k];
  }
  bool ok = true;
  for (int64_t step = 0; ok && step < common; ++step) {
#pragma GCC unroll 8
    for (int k = 0; k < W; ++k) {
      ok = encode_step<kBranchFree>(t, (int32_t)sym[k][i[k]], (int32_t)idx[k][i[k]], x[k], w[k]) && ok;
      --i[k];
    }
  }
  for (int k = 0; k < W; ++k)
    for (; ok && i[k] >= 0; --i[k]) ok = encode_step<kBranchFree>(t, (int32_t)sym[k][i[k]], (int32_t)idx[k][i[k]], x[k], w[k]);
  for (int k = 0; k < W; ++k) nb[k] = ok ? encode_finish(x[k], w[k], buf_end[k]) : (int64_t)STF_E_ARG;
}

template <int W, class SymT, class IdxT>
void encode_intoW(const stf_rans_table *t, const SymT *const *sym, const IdxT *const *idx, const int64_t *n,
                  uint32_t *const *buf_end, int64_t *nb) {
  if (wide_run(t, idx[0], n[0])) encode_intoW_impl<W, true, SymT, IdxT>(t, sym, idx, n, buf_end, nb);
  else encode_intoW_impl<W, false, SymT, IdxT>(t, sym, idx, n, buf_end, nb);
}

// Decoder position kept in registers while a run is in flight (the decoder object is only read at the start of a run and
// written back at its end).
struct Cursor {
  uint64_t x;
  const uint32_t *w, *end;
};

inline bool refill(Cursor &c) {   // branchy form, for the rare paths (escape nibbles)
  if (c.x < kLow) {
    if (c.w >= c.end) return false;
    c.x = (c.x << 32) | *c.w++;
  }
  return true;
}

inline bool get_nibble(Cursor &c, int32_t *val) {
  *val = (int32_t)(c.x & ((1u << kNibbleBits) - 1));
  c.x >>= kNibbleBits;
  return refill(c);
}

// One symbol of one stream.  The two data-dependent decisions of a step -- does the bucket's first symbol already hold
// `cum`, does the state need a new word -- are taken without branches (at ~10 bits per symbol both are coin flips); only
// the rare cases (a third symbol in a 16-count bucket, stream exhausted, escape) branch.
template <bool kBranchFree>
inline int decode_step(Cursor &c, const stf_rans_table *t, int32_t row, int32_t *out) {
  if ((uint32_t)row >= (uint32_t)t->rows) return STF_E_ARG;
  const stf_rans_table::Row ri = t->row[row];
  const uint32_t *cdf = t->cdf.data() + ri.base + (uint32_t)row;
  const int32_t escape = ri.escape;
  const uint64_t x = c.x;
  const uint32_t cum = (uint32_t)x & (kProbScale - 1);
  uint32_t s;
  if (kBranchFree) {
    s = t->lut_wide[((size_t)row << kLutBitsWide) + (cum >> (kProbBits - kLutBitsWide))];
    s += (uint32_t)(cdf[s + 1] <= cum);
    while (cdf[s + 1] <= cum) ++s;
    const uint32_t start = cdf[s], freq = cdf[s + 1] - start;
    uint64_t nx = freq * (x >> kProbBits) + cum - start;  // Rans64DecAdvance, rans64.h:126-142
    const uint64_t need = (uint64_t)(nx < kLow);            // Rans64DecRenorm
    const bool have = c.w < c.end;
    if (need && !have) return STF_E_STREAM;
    const uint64_t word = have ? *c.w : 0u;
    const uint64_t mask = 0 - need;
    nx = ((nx << (need << 5)) | (word & mask));
    c.w += need;
    c.x = nx;
  } else {
    s = t->lut_narrow[((size_t)row << kLutBitsNarrow) + (cum >> (kProbBits - kLutBitsNarrow))];
    while (cdf[s + 1] <= cum) ++s;
    const uint32_t start = cdf[s], freq = cdf[s + 1] - start;
    c.x = freq * (x >> kProbBits) + cum - start;
    if (!refill(c)) return STF_E_STREAM;
  }
  int32_t v = (int32_t)s;
  if (v == escape) {  // rans_interface.cpp:320-343
    int32_t nib, nn;
    if (!get_nibble(c, &nib)) return STF_E_STREAM;
    nn = nib;
    while (nib == kNibbleMax) {
      if (!get_nibble(c, &nib)) return STF_E_STREAM;
      nn += nib;
    }
    int32_t raw = 0;
    for (int32_t j = 0; j < nn; ++j) {
      if (!get_nibble(c, &nib)) return STF_E_STREAM;
      if (j < 8) raw |= nib << (j * kNibbleBits);
    }
    v = raw >> 1;
    v = (raw & 1) ? -v - 1 : v + escape;
  }
  *out = v + ri.offset;
  return STF_OK;
}

template <bool kBranchFree, class IdxT>
__attribute__((noinline)) int decode_run_impl(stf_rans_decoder *d, const stf_rans_table *t, const IdxT *indexes, int64_t n, int32_t *out) {
  Cursor c{d->x, d->w, d->end};
  for (int64_t i = 0; i < n; ++i) {
    const int rc = decode_step<kBranchFree>(c, t, (int32_t)indexes[i], out + i);
    if (rc) return rc;
  }
  d->x = c.x, d->w = c.w;
  return STF_OK;
}

template <class IdxT>
int decode_run(stf_rans_decoder *d, const stf_rans_table *t, const IdxT *indexes, int64_t n, int32_t *out) {
  return wide_run(t, indexes, n) ? decode_run_impl<true, IdxT>(d, t, indexes, n, out)
                                 : decode_run_impl<false, IdxT>(d, t, indexes, n, out);
}

// W decoders in lockstep (see encode_intoW).
template <int W, bool kBranchFree, class IdxT>
__attribute__((noinline)) void decode_runW_impl(stf_rans_decoder *const *d, const stf_rans_table *t, const IdxT *const *idx, const int64_t *n,
                      int32_t *const *out, int *rc_out) {
  Cursor c[W];
  int rc[W];   // local: the caller's status array is shared between threads (one cache line for several groups)
  int64_t m = n[0];
  for (int k = 0; k < W; ++k) {
    c[k] = Cursor{d[k]->x, d[k]->w, d[k]->end}, rc[k] = STF_OK;
    if (n[k] < m) m = n[k];
  }
  int64_t i = 0;
  int any = 0;
  for (; i < m && !any; ++i) {
#pragma GCC unroll 8
    for (int k = 0; k < W; ++k) {
      rc[k] = decode_step<kBranchFree>(c[k], t, (int32_t)idx[k][i], out[k] + i);
      any |= rc[k];
    }
  }
  if (!any)
    for (int k = 0; k < W; ++k) {
      for (int64_t j = i; j < n[k] && !rc[k]; ++j) rc[k] = decode_step<kBranchFree>(c[k], t, (int32_t)idx[k][j], out[k] + j);
      if (!rc[k]) d[k]->x = c[k].x, d[k]->w = c[k].w;
    }
  for (int k = 0; k < W; ++k) rc_out[k] = rc[k];
}

template <int W, class IdxT>
void decode_runW(stf_rans_decoder *const *d, const stf_rans_table *t, const IdxT *const *idx, const int64_t *n,
                 int32_t *const *out, int *rc_out) {
  if (wide_run(t, idx[0], n[0])) decode_runW_impl<W, true, IdxT>(d, t, idx, n, out, rc_out);
  else decode_runW_impl<W, false, IdxT>(d, t, idx, n, out, rc_out);
}

// Persistent worker pool: decode_batch is called once per slice (12-13 times per image batch), so
// spawning std::threads per call (~25 us each) would cost more than the decoding itself.
// The pool is the only process-wide state of the library; it is created on first use, its workers
// sleep on a condition variable between calls, and concurrent callers are serialised.
class Pool {
 public:
  static Pool &get() {
    static Pool p;
    return p;
  }
  template <class F>
  void run(int count, int threads, F f) {
    if (count <= 0) return;
    if (threads > count) threads = count;
    if (threads <= 1) {
      for (int i = 0; i < count; ++i) f(i);
      return;
    }
    std::lock_guard<std::mutex> caller(run_mutex_);
    ensure_workers(threads - 1);
    std::function<void(int)> fn = f;
    {
      std::lock_guard<std::mutex> lk(m_);
      fn_ = &fn;
      next_ = 0;
      count_ = count;
      pending_ = count;
      helpers_ = threads - 1;
      ++generation_;
    }
    cv_work_.notify_all();
    drain();  // the calling thread works too
    std::unique_lock<std::mutex> lk(m_);
    cv_done_.wait(lk, [&] { return pending_ == 0 && busy_ == 0; });
    fn_ = nullptr;
  }

 private:
  Pool() = default;
  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_work_.notify_all();
    for (auto &t : workers_) t.join();
  }
  void ensure_workers(int n) {
    while ((int)workers_.size() < n) {
      const int id = (int)workers_.size();
      workers_.emplace_back([this, id] { worker(id); });
    }
  }
  void drain() {
    for (;;) {
      int i;
      {
        std::lock_guard<std::mutex> lk(m_);
        if (next_ >= count_) return;
        i = next_++;
      }
      (*fn_)(i);
      std::lock_guard<std::mutex> lk(m_);
      if (--pending_ == 0) cv_done_.notify_all();
    }
  }
  void worker(int id) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_work_.wait(lk, [&] { return stop_ || (generation_ != seen && id < helpers_ && next_ < count_); });
        if (stop_) return;
        seen = generation_;
        ++busy_;
      }
      drain();
      std::lock_guard<std::mutex> lk(m_);
      if (--busy_ == 0 && pending_ == 0) cv_done_.notify_all();
    }
  }
  std::mutex run_mutex_, m_;
  std::condition_variable cv_work_, cv_done_;
  std::vector<std::thread> workers_;
  const std::function<void(int)> *fn_ = nullptr;
  int next_ = 0, count_ = 0, pending_ = 0, helpers_ = 0, busy_ = 0;
  uint64_t generation_ = 0;
  bool stop_ = false;
};

template <class F>
void parallel_for(int count, int threads, F f) {
  Pool::get().run(count, threads, f);
}

}  // namespace

extern "C" stf_rans_table *stf_rans_table_create(const int32_t *cdf, int rows, int row_stride,
                                                 const int32_t *sizes, const int32_t *offsets) {
  if (!cdf || !sizes || !offsets || rows <= 0 || row_stride < 2) return nullptr;
  stf_rans_table *t = new (std::nothrow) stf_rans_table;
  if (!t) return nullptr;
  t->rows = rows;
  t->sizes.assign(sizes, sizes + rows);
  t->offsets.assign(offsets, offsets + rows);
  t->base.resize(rows);
  t->cbase.resize(rows);
  for (int r = 0; r < rows; ++r) {
    const int32_t *c = cdf + (size_t)r * row_stride;
    const int sz = sizes[r];
    bool ok = sz >= 2 && sz <= row_stride && sz <= 65537 && c[0] == 0 && c[sz - 1] == (int32_t)kProbScale;
    for (int j = 0; ok && j + 1 < sz; ++j) ok = c[j + 1] > c[j];
    if (!ok) {
      delete t;
      return nullptr;
    }
    t->base[r] = (uint32_t)t->enc.size();
    t->cbase[r] = (uint32_t)t->cdf.size();
    t->row.push_back({offsets[r], sz - 2, t->base[r], sz - 1 > kWideRow ? 1u : 0u});
    for (int j = 0; j < sz; ++j) t->cdf.push_back((uint32_t)c[j]);
    for (int j = 0; j + 1 < sz; ++j) {
      EncEntry e;
      make_entry(&e, (uint32_t)c[j], (uint32_t)(c[j + 1] - c[j]));
      t->enc.push_back(e);
    }
    for (int pass = 0; pass < 2; ++pass) {
      const int bits = pass ? kLutBitsWide : kLutBitsNarrow;
      std::vector<uint16_t> &lut = pass ? t->lut_wide : t->lut_narrow;
      lut.resize(((size_t)r + 1) << bits);
      uint32_t s = 0;
      for (uint32_t b = 0; b < (1u << bits); ++b) {
        const uint32_t cum = b << (kProbBits - bits);
        while ((uint32_t)c[s + 1] <= cum) ++s;
        lut[((size_t)r << bits) + b] = (uint16_t)s;
      }
    }
  }
  return t;
}

extern "C" void stf_rans_table_destroy(stf_rans_table *t) { delete t; }

// (library-internal) read access for the device-table packer in csrc/rans_device.cu
extern "C" int stf_rans_table_export(const stf_rans_table *t, int *rows, const int32_t **sizes, const int32_t **offsets,
                                     const uint32_t **cdf, const uint32_t **cbase) {
  if (!t) return STF_E_ARG;
  *rows = t->rows, *sizes = t->sizes.data(), *offsets = t->offsets.data(), *cdf = t->cdf.data(), *cbase = t->cbase.data();
  return STF_OK;
}

extern "C" int64_t stf_rans_encode_bound(int64_t n) { return n < 0 ? STF_E_ARG : 8 * n + 64; }

namespace {

template <class SymT, class IdxT>
int64_t encode_one(const stf_rans_table *t, const SymT *symbols, const IdxT *indexes, int64_t n, uint8_t *out,
                   int64_t out_cap) {
  if (!t || !out || n < 0 || (n > 0 && (!symbols || !indexes))) return STF_E_ARG;
  const int64_t bound = stf_rans_encode_bound(n);
  if (out_cap >= bound && ((uintptr_t)out & 3u) == 0) {  // code straight into the caller's buffer
    uint32_t *end = reinterpret_cast<uint32_t *>(out) + out_cap / 4;
    int64_t nb = encode_into(t, symbols, indexes, n, end);
    if (nb < 0) return nb;
    memmove(out, reinterpret_cast<uint8_t *>(end) - nb, (size_t)nb);
    return nb;
  }
  std::vector<uint32_t> tmp((size_t)bound / 4);
  int64_t nb = encode_into(t, symbols, indexes, n, tmp.data() + tmp.size());
  if (nb < 0) return nb;
  if (nb > out_cap) return STF_E_OVERFLOW;
  memcpy(out, reinterpret_cast<uint8_t *>(tmp.data() + tmp.size()) - nb, (size_t)nb);
  return nb;
}

template <class SymT, class IdxT>
int encode_batch(const stf_rans_table *t, int count, const SymT *const *symbols, const IdxT *const *indexes,
                 const int64_t *n, uint8_t *const *out, const int64_t *out_cap, int64_t *out_lens, int threads) {
  if (!t || count < 0 || !symbols || !indexes || !n || !out || !out_cap || !out_lens) return STF_E_ARG;
  bool groupable = count > threads && threads >= 1;
  for (int i = 0; groupable && i < count; ++i)
    groupable = n[i] >= 0 && symbols[i] && indexes[i] && out[i] && out_cap[i] >= stf_rans_encode_bound(n[i]) &&
                ((uintptr_t)out[i] & 3u) == 0;
  if (!groupable) {
    parallel_for(count, threads,
                 [&](int i) { out_lens[i] = encode_one(t, symbols[i], indexes[i], n[i], out[i], out_cap[i]); });
  } else {
    // more images than threads: the images are split evenly over the threads (contiguous ranges whose sizes differ by at
    // most one -- fixed groups of four would leave half the threads idle in the last round of 21 images on 4 threads) and
    // every thread codes its range in lockstep groups of 4, 3, 2 or 1, straight into the callers' buffers
    parallel_for(threads, threads, [&](int tix) {
      const int lo = (int)((int64_t)count * tix / threads), hi = (int)((int64_t)count * (tix + 1) / threads);
      for (int a = lo; a < hi;) {
        const int m = hi - a < kMaxLockstep ? hi - a : kMaxLockstep;
        const SymT *sy[8];
        const IdxT *ix[8];
        int64_t nn[8], nb[8];
        uint32_t *end[8];
        for (int k = 0; k < m; ++k)
          sy[k] = symbols[a + k], ix[k] = indexes[a + k], nn[k] = n[a + k],
          end[k] = reinterpret_cast<uint32_t *>(out[a + k]) + out_cap[a + k] / 4;
        if (m == 8) encode_intoW<8>(t, sy, ix, nn, end, nb);
        else if (m == 7) encode_intoW<7>(t, sy, ix, nn, end, nb);
        else if (m == 6) encode_intoW<6>(t, sy, ix, nn, end, nb);
        else if (m == 5) encode_intoW<5>(t, sy, ix, nn, end, nb);
        else if (m == 4) encode_intoW<4>(t, sy, ix, nn, end, nb);
        else if (m == 3) encode_intoW<3>(t, sy, ix, nn, end, nb);
        else if (m == 2) encode_intoW<2>(t, sy, ix, nn, end, nb);
        else nb[0] = encode_into(t, sy[0], ix[0], nn[0], end[0]);
        for (int k = 0; k < m; ++k) {
          if (nb[k] > 0) memmove(out[a + k], reinterpret_cast<uint8_t *>(end[k]) - nb[k], (size_t)nb[k]);
          out_lens[a + k] = nb[k];
        }
        a += m;
      }
    });
  }
  for (int i = 0; i < count; ++i)
    if (out_lens[i] < 0) return (int)out_lens[i];
  return STF_OK;
}

template <class IdxT>
int decode_batch(stf_rans_decoder *const *d, const stf_rans_table *t, int count, const IdxT *const *indexes,
                 const int64_t *n, int32_t *const *symbols_out, int threads) {
  if (!d || !t || count < 0 || !indexes || !n || !symbols_out) return STF_E_ARG;
  std::vector<int> rc((size_t)count, 0);
  bool groupable = count > threads && threads >= 1;
  for (int i = 0; i < count; ++i) {
    const bool ok = d[i] && n[i] >= 0 && (n[i] == 0 || (indexes[i] && symbols_out[i]));
    if (!ok) return STF_E_ARG;
  }
  if (!groupable) {
    parallel_for(count, threads, [&](int i) { rc[i] = decode_run(d[i], t, indexes[i], n[i], symbols_out[i]); });
  } else {
    parallel_for(threads, threads, [&](int tix) {   // even contiguous ranges, lockstep groups of 4 / 3 / 2 / 1 (see encode)
      const int lo = (int)((int64_t)count * tix / threads), hi = (int)((int64_t)count * (tix + 1) / threads);
      for (int a = lo; a < hi;) {
        const int m = hi - a < kMaxLockstep ? hi - a : kMaxLockstep;
        int r[8];
        if (m == 8) decode_runW<8>(d + a, t, indexes + a, n + a, symbols_out + a, r);
        else if (m == 7) decode_runW<7>(d + a, t, indexes + a, n + a, symbols_out + a, r);
        else if (m == 6) decode_runW<6>(d + a, t, indexes + a, n + a, symbols_out + a, r);
        else if (m == 5) decode_runW<5>(d + a, t, indexes + a, n + a, symbols_out + a, r);
        else if (m == 4) decode_runW<4>(d + a, t, indexes + a, n + a, symbols_out + a, r);
        else if (m == 3) decode_runW<3>(d + a, t, indexes + a, n + a, symbols_out + a, r);
        else if (m == 2) decode_runW<2>(d + a, t, indexes + a, n + a, symbols_out + a, r);
        else r[0] = decode_run(d[a], t, indexes[a], n[a], symbols_out[a]);
        for (int k = 0; k < m; ++k) rc[a + k] = r[k];
        a += m;
      }
    });
  }
  for (int i = 0; i < count; ++i)
    if (rc[i]) return rc[i];
  return STF_OK;
}

}  // namespace

extern "C" int64_t stf_rans_encode(const stf_rans_table *t, const int32_t *symbols, const int32_t *indexes,
                                   int64_t n, uint8_t *out, int64_t out_cap) {
  return encode_one(t, symbols, indexes, n, out, out_cap);
}

extern "C" int stf_rans_encode_batch(const stf_rans_table *t, int count, const int32_t *const *symbols,
                                     const int32_t *const *indexes, const int64_t *n, uint8_t *const *out,
                                     const int64_t *out_cap, int64_t *out_lens, int threads) {
  return encode_batch(t, count, symbols, indexes, n, out, out_cap, out_lens, threads);
}

// The same coder on the narrow transfer format of stf_slice_step_nhwc (int16 symbols, uint8 indexes: 3 instead of 8 bytes
// per symbol over PCIe and through the host caches).  Same bytes out.
extern "C" int stf_rans_encode_batch_narrow(const stf_rans_table *t, int count, const int16_t *const *symbols,
                                            const uint8_t *const *indexes, const int64_t *n, uint8_t *const *out,
                                            const int64_t *out_cap, int64_t *out_lens, int threads) {
  return encode_batch(t, count, symbols, indexes, n, out, out_cap, out_lens, threads);
}

extern "C" stf_rans_decoder *stf_rans_decoder_create(const uint8_t *stream, int64_t nbytes) {
  if (!stream || nbytes < 8 || (nbytes & 3)) return nullptr;
  stf_rans_decoder *d = new (std::nothrow) stf_rans_decoder;
  if (!d) return nullptr;
  d->words.resize((size_t)nbytes / 4);
  memcpy(d->words.data(), stream, (size_t)nbytes);
  d->w = d->words.data();
  d->end = d->w + d->words.size();
  d->x = (uint64_t)d->w[0] | ((uint64_t)d->w[1] << 32);  // Rans64DecInit, rans64.h:107-115
  d->w += 2;
  return d;
}

// Same, without copying the stream: the caller keeps `stream` alive and unchanged for the decoder's lifetime
// (a batch of 32 Kodak-size y-strings is 19 MB of memcpy otherwise).  Unaligned buffers fall back to the copy.
extern "C" stf_rans_decoder *stf_rans_decoder_create_view(const uint8_t *stream, int64_t nbytes) {
  if (!stream || nbytes < 8 || (nbytes & 3)) return nullptr;
  if ((uintptr_t)stream & 3u) return stf_rans_decoder_create(stream, nbytes);
  stf_rans_decoder *d = new (std::nothrow) stf_rans_decoder;
  if (!d) return nullptr;
  d->w = reinterpret_cast<const uint32_t *>(stream);
  d->end = d->w + nbytes / 4;
  d->x = (uint64_t)d->w[0] | ((uint64_t)d->w[1] << 32);
  d->w += 2;
  return d;
}

extern "C" void stf_rans_decoder_destroy(stf_rans_decoder *d) { delete d; }

extern "C" int stf_rans_decode(stf_rans_decoder *d, const stf_rans_table *t, const int32_t *indexes, int64_t n,
                               int32_t *symbols_out) {
  if (!d || !t || n < 0 || (n > 0 && (!indexes || !symbols_out))) return STF_E_ARG;
  return decode_run(d, t, indexes, n, symbols_out);
}

extern "C" int stf_rans_decode_batch(stf_rans_decoder *const *d, const stf_rans_table *t, int count,
                                     const int32_t *const *indexes, const int64_t *n,
                                     int32_t *const *symbols_out, int threads) {
  return decode_batch(d, t, count, indexes, n, symbols_out, threads);
}

// uint8 indexes (a quarter of the device -> host bytes of every slice of decompress()); symbols stay int32: what a
// stream decodes to is not bounded.
extern "C" int stf_rans_decode_batch_u8(stf_rans_decoder *const *d, const stf_rans_table *t, int count,
                                        const uint8_t *const *indexes, const int64_t *n,
                                        int32_t *const *symbols_out, int threads) {
  return decode_batch(d, t, count, indexes, n, symbols_out, threads);
}

extern "C" int stf_pmf_to_quantized_cdf(const float *pmf, int n, int precision, uint32_t *cdf) {
  if (!pmf || !cdf || n < 1 || precision < 1 || precision > 16) return STF_E_ARG;
  const int m = n + 1;
  const uint32_t scale = 1u << precision;
  cdf[0] = 0;
  uint32_t total = 0;
  for (int i = 0; i < n; ++i) {
    cdf[i + 1] = (uint32_t)std::round(pmf[i] * (float)scale);
    total += cdf[i + 1];
  }
  if (total == 0) return STF_E_TABLE;
  uint32_t run = 0;
  for (int i = 0; i < m; ++i) {
    run += (uint32_t)(((uint64_t)scale * cdf[i]) / total);
    cdf[i] = run;
  }
  cdf[m - 1] = scale;
  for (int i = 0; i + 1 < m; ++i) {
    if (cdf[i] != cdf[i + 1]) continue;
    uint32_t best = ~0u;  // cheapest donor with more than one count (ops.cpp:52-60)
    int donor = -1;
    for (int j = 0; j + 1 < m; ++j) {
      uint32_t f = cdf[j + 1] - cdf[j];
      if (f > 1 && f < best) best = f, donor = j;
    }
    if (donor < 0) return STF_E_TABLE;
    if (donor < i)
      for (int j = donor + 1; j <= i; ++j) cdf[j]--;
    else
      for (int j = i + 1; j <= donor; ++j) cdf[j]++;
  }
  return STF_OK;
}
