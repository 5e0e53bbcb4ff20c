/*
 * oracle/rans_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU checker; never on the product path).
 *
 * Plain-C restatement of the reference's entropy-coder back end, written from its published
 * algorithm so the tests can check stf_b200's host codec and tables bit-for-bit:
 *
 *   - rANS64 primitives (L = 2^31, 32-bit renormalisation words, 16-bit probabilities)
 *       reference: third_party/ryg_rans/rans64.h:59-142
 *   - symbol staging with per-symbol CDF index, escape symbol + 4-bit "bypass" nibbles
 *       reference: compressai/cpp_exts/rans/rans_interface.cpp:99-164 (encode_with_indexes)
 *                  compressai/cpp_exts/rans/rans_interface.cpp:166-191 (flush, reverse order)
 *                  compressai/cpp_exts/rans/rans_interface.cpp:206-275, 285-350 (decoders)
 *   - pmf -> 16-bit quantised cdf with "steal from the cheapest symbol" repair
 *       reference: compressai/cpp_exts/ops/ops.cpp:24-81
 *
 * Parity pin: tests/test_oracle_pins.py checks this file against (a) the reference's own compiled
 * extension in oracle/_ref (built from /root/reference by oracle/Makefile) on random streams and
 * (b) the known-answer vectors recorded from the live reference (tests/golden/kat.json).
 *
 * Deliberately simple: linear CDF search exactly like the reference's std::find_if, one staged
 * record per symbol, no tables, no threads.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PROB_BITS 16u
#define NIBBLE_BITS 4u
#define NIBBLE_MAX 15
#define RANS_LOW (1ull << 31)

typedef struct {
  uint16_t start;
  uint16_t range;
  uint8_t raw; /* 1: write `start` as NIBBLE_BITS raw bits */
} staged_t;

typedef struct {
  staged_t *v;
  size_t n, cap;
} staged_vec;

static int push(staged_vec *s, uint32_t start, uint32_t range, int raw) {
  if (s->n == s->cap) {
    size_t cap = s->cap ? s->cap * 2 : 1024;
    staged_t *nv = (staged_t *)realloc(s->v, cap * sizeof(staged_t));
    if (!nv) return -1;
    s->v = nv;
    s->cap = cap;
  }
  s->v[s->n].start = (uint16_t)start;
  s->v[s->n].range = (uint16_t)range;
  s->v[s->n].raw = (uint8_t)raw;
  s->n++;
  return 0;
}

/* rans_interface.cpp:99-164 */
static int stage_symbols(staged_vec *out, const int32_t *symbols, const int32_t *indexes, long n,
                         const int32_t *cdfs, int cdf_stride, const int32_t *cdf_sizes,
                         const int32_t *offsets) {
  for (long i = 0; i < n; ++i) {
    const int32_t t = indexes[i];
    const int32_t *cdf = cdfs + (size_t)t * cdf_stride;
    const int32_t escape = cdf_sizes[t] - 2;
    int32_t v = symbols[i] - offsets[t];
    uint32_t raw = 0;
    if (v < 0) {
      raw = (uint32_t)(-2 * v - 1);
      v = escape;
    } else if (v >= escape) {
      raw = (uint32_t)(2 * (v - escape));
      v = escape;
    }
    if (push(out, (uint32_t)cdf[v], (uint32_t)(cdf[v + 1] - cdf[v]), 0)) return -1;
    if (v == escape) {
      int32_t nn = 0;
      while ((raw >> (nn * NIBBLE_BITS)) != 0) ++nn;
      int32_t left = nn;
      while (left >= NIBBLE_MAX) {
        if (push(out, NIBBLE_MAX, NIBBLE_MAX + 1, 1)) return -1;
        left -= NIBBLE_MAX;
      }
      if (push(out, (uint32_t)left, (uint32_t)left + 1, 1)) return -1;
      for (int32_t j = 0; j < nn; ++j) {
        uint32_t nib = (raw >> (j * NIBBLE_BITS)) & NIBBLE_MAX;
        if (push(out, nib, nib + 1, 1)) return -1;
      }
    }
  }
  return 0;
}

/* rans64.h:77-93 (Rans64EncPut) and rans_interface.cpp:59-77 (Rans64EncPutBits) */
static inline void enc_put(uint64_t *x, uint32_t **w, uint32_t start, uint32_t freq, uint32_t bits) {
  uint64_t s = *x;
  uint64_t lim = ((RANS_LOW >> bits) << 32) * freq;
  if (s >= lim) {
    *w -= 1;
    **w = (uint32_t)s;
    s >>= 32;
  }
  *x = ((s / freq) << bits) + (s % freq) + start;
}
static inline void enc_put_raw(uint64_t *x, uint32_t **w, uint32_t val, uint32_t nbits) {
  uint64_t s = *x;
  uint32_t freq = 1u << (16 - nbits);
  uint64_t lim = ((RANS_LOW >> 16) << 32) * freq;
  if (s >= lim) {
    *w -= 1;
    **w = (uint32_t)s;
    s >>= 32;
  }
  *x = (s << nbits) | val;
}

/*
 * Encode n symbols; returns the byte length of the stream written to out[0..), or -1 when
 * out_cap is too small / out of memory.  Stream = what RansEncoder.encode_with_indexes returns
 * (rans_interface.cpp:193-204): native-endian 32-bit words, final state first.
 */
long oracle_rans_encode(const int32_t *symbols, const int32_t *indexes, long n, const int32_t *cdfs,
                        int cdf_stride, const int32_t *cdf_sizes, const int32_t *offsets,
                        uint8_t *out, long out_cap) {
  staged_vec st = {0, 0, 0};
  if (stage_symbols(&st, symbols, indexes, n, cdfs, cdf_stride, cdf_sizes, offsets)) {
    free(st.v);
    return -1;
  }
  size_t words = st.n + 2;
  uint32_t *buf = (uint32_t *)malloc(words * sizeof(uint32_t));
  if (!buf) {
    free(st.v);
    return -1;
  }
  uint32_t *w = buf + words;
  uint64_t x = RANS_LOW; /* rans64.h:68 */
  for (size_t k = st.n; k-- > 0;) {
    const staged_t *e = &st.v[k];
    if (e->raw)
      enc_put_raw(&x, &w, e->start, NIBBLE_BITS);
    else
      enc_put(&x, &w, e->start, e->range, PROB_BITS);
  }
  w -= 2; /* rans64.h:96-103 */
  w[0] = (uint32_t)x;
  w[1] = (uint32_t)(x >> 32);
  long nbytes = (long)((buf + words) - w) * 4;
  long rc = -1;
  if (nbytes <= out_cap) {
    memcpy(out, w, (size_t)nbytes);
    rc = nbytes;
  }
  free(buf);
  free(st.v);
  return rc;
}

/* Decoder state kept by the caller so that decode_stream (rans_interface.cpp:277-350) can be
 * restated: several calls consume one stream. */
typedef struct {
  uint64_t x;
  const uint32_t *w;
} oracle_dec_t;

void oracle_rans_dec_init(oracle_dec_t *d, const uint8_t *stream) {
  const uint32_t *w = (const uint32_t *)stream; /* rans64.h:107-115 */
  d->x = (uint64_t)w[0] | ((uint64_t)w[1] << 32);
  d->w = w + 2;
}

static inline uint32_t dec_raw(oracle_dec_t *d, uint32_t nbits) { /* rans_interface.cpp:79-96 */
  uint64_t s = d->x;
  uint32_t val = (uint32_t)(s & ((1u << nbits) - 1));
  s >>= nbits;
  if (s < RANS_LOW) {
    s = (s << 32) | *d->w;
    d->w += 1;
  }
  d->x = s;
  return val;
}

void oracle_rans_dec_run(oracle_dec_t *d, const int32_t *indexes, long n, const int32_t *cdfs,
                         int cdf_stride, const int32_t *cdf_sizes, const int32_t *offsets,
                         int32_t *out) {
  for (long i = 0; i < n; ++i) {
    const int32_t t = indexes[i];
    const int32_t *cdf = cdfs + (size_t)t * cdf_stride;
    const int32_t escape = cdf_sizes[t] - 2;
    const uint32_t cum = (uint32_t)(d->x & ((1u << PROB_BITS) - 1)); /* rans64.h:118-121 */
    int32_t s = 0; /* linear search like std::find_if, rans_interface.cpp:228-232 */
    while ((uint32_t)cdf[s + 1] <= cum) ++s;
    {
      /* rans64.h:126-142 */
      uint64_t xs = d->x;
      uint32_t start = (uint32_t)cdf[s], freq = (uint32_t)(cdf[s + 1] - cdf[s]);
      xs = freq * (xs >> PROB_BITS) + (xs & ((1ull << PROB_BITS) - 1)) - start;
      if (xs < RANS_LOW) {
        xs = (xs << 32) | *d->w;
        d->w += 1;
      }
      d->x = xs;
    }
    int32_t v = s;
    if (v == escape) {
      int32_t nib = (int32_t)dec_raw(d, NIBBLE_BITS);
      int32_t nn = nib;
      while (nib == NIBBLE_MAX) {
        nib = (int32_t)dec_raw(d, NIBBLE_BITS);
        nn += nib;
      }
      int32_t raw = 0;
      for (int32_t j = 0; j < nn; ++j) {
        nib = (int32_t)dec_raw(d, NIBBLE_BITS);
        raw |= nib << (j * NIBBLE_BITS);
      }
      v = raw >> 1;
      if (raw & 1)
        v = -v - 1;
      else
        v += escape;
    }
    out[i] = v + offsets[t];
  }
}

/* ops.cpp:24-81.  cdf has n+1 entries.  Returns 0, or -1 if no symbol can donate frequency. */
int oracle_pmf_to_quantized_cdf(const float *pmf, int n, int precision, uint32_t *cdf) {
  const int m = n + 1;
  cdf[0] = 0;
  for (int i = 0; i < n; ++i) cdf[i + 1] = (uint32_t)roundf(pmf[i] * (float)(1 << precision));
  uint32_t total = 0;
  for (int i = 0; i < m; ++i) total += cdf[i];
  for (int i = 0; i < m; ++i)
    cdf[i] = (uint32_t)((((uint64_t)(1u << precision)) * cdf[i]) / total);
  for (int i = 1; i < m; ++i) cdf[i] += cdf[i - 1];
  cdf[m - 1] = 1u << precision;
  for (int i = 0; i < m - 1; ++i) {
    if (cdf[i] != cdf[i + 1]) continue;
    uint32_t best = ~0u;
    int donor = -1;
    for (int j = 0; j < m - 1; ++j) {
      uint32_t f = cdf[j + 1] - cdf[j];
      if (f > 1 && f < best) {
        best = f;
        donor = j;
      }
    }
    if (donor < 0) return -1;
    if (donor < i) {
      for (int j = donor + 1; j <= i; ++j) cdf[j]--;
    } else {
      for (int j = i + 1; j <= donor; ++j) cdf[j]++;
    }
  }
  return 0;
}
