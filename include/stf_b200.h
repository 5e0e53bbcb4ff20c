/*
 * stf_b200.h -- C ABI of libstf_b200.so: the B200 (sm_100a) implementation of the STF / WACNN
 * data-parallel hot path.  This header is the drop-in boundary: plain
 * pointers and sizes, no torch / C++ types, no exceptions across it.
 *
 * Conventions (all entry points)
 *   - return 0 on success; < 0 = argument / shape error (STF_E_*); > 0 = cudaError_t.
 *   - device pointers are raw CUDA device addresses, fp32 / int32, contiguous, 16-byte aligned.
 *   - the caller owns every buffer (outputs and workspaces included); the library allocates no
 *     device memory and keeps no mutable global state, so calls are thread-safe.
 *   - all GPU work is enqueued on `stream` (a cudaStream_t passed as void*); no host sync.
 *   - kernels are deterministic run-to-run and batch-invariant (no atomics, fixed reduction
 *     order): decode must rebuild the encoder's indexes bit-for-bit (reference stf.py:767).
 *
 * Each function cites the reference interface it replaces (paths relative to the reference
 * repository memory4963/STF).  The reference has no FFI for this path -- it is PyTorch eager
 * plus two pybind11 modules -- so the binding a maintainer adds is the ctypes stub shown in
 * INTEGRATION.md (stf_b200/_C.py is that stub).
 */
#ifndef STF_B200_H
#define STF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STF_OK 0
#define STF_E_ARG (-1)       /* null pointer / negative size */
#define STF_E_SHAPE (-2)     /* unsupported or inconsistent shape */
#define STF_E_ALIGN (-3)     /* pointer not 16-byte aligned */
#define STF_E_TABLE (-4)     /* invalid CDF / scale table */
#define STF_E_OVERFLOW (-5)  /* output buffer too small */
#define STF_E_STREAM (-6)    /* corrupt / truncated bitstream */

/* Library / build identification: "stf_b200 <ver> sm_100a". */
const char *stf_version(void);
/* Number of CUDA kernels this library has launched in the calling process (for bench.py's
 * gpu_launches claim).  Monotonic, relaxed atomic. */
int64_t stf_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Entropy-model element-wise kernels (HBM-bound).
 * ------------------------------------------------------------------------------------------ */

/* GaussianConditional.build_indexes (compressai/entropy_models/entropy_models.py:661-666):
 *   sigma = max(scales, scale_bound);  idx = (levels-1) - #{ i < levels-1 : sigma <= table[i] }
 * NaN -> levels-1.  `table_host` is a HOST array of `levels` (<= 64) fp32 values.  8 B/element. */
int stf_build_indexes(const float *scales, int32_t *indexes, int64_t n, const float *table_host,
                      int levels, float scale_bound, void *stream);

/* One compress step of the slice loop (stf.py:717-719 / cnn.py:246-248) fused:
 *   indexes = build_indexes(scales); symbols = int32(round_half_even(y - means));
 *   y_hat = float(symbols) + means
 * y may be a channel slice of a larger NCHW tensor: element (b, c, p) is read at
 * y[b * y_batch_stride + c * plane + p] with c < channels, p < plane; scales / means / y_hat are
 * dense (batch, channels, plane).  symbols / indexes are written at
 * out[b * out_batch_stride + c * plane + p] so that all slices of one image land contiguously in
 * the reference's coding order (slice-major, then channel, row, column; stf.py:721-722).
 * Any of symbols / indexes / y_hat may be NULL.  24 B/element with all three outputs. */
int stf_gaussian_compress_step(const float *y, int64_t y_batch_stride, const float *scales,
                               const float *means, int32_t *symbols, int32_t *indexes,
                               int64_t out_batch_stride, float *y_hat, int batch, int channels,
                               int64_t plane, const float *table_host, int levels,
                               float scale_bound, void *stream);

/* The same slice steps on NHWC operands -- the layout of the convolution kernel below, so the slice loop runs without a
 * single layout copy: mu / scale / y / y_hat are pixel-major (element (b, p, c) at [(b * plane + p) * ld + c]; a channel
 * slice of a wider NHWC tensor is fine, e.g. y_hat goes straight into its 32-channel slot of the support buffer), while
 * symbols / indexes / likelihoods are written in the reference's coding order (b, c, p) (stf.py:721-722).  Which step runs
 * follows from the non-NULL pointers:
 *   likelihood != NULL            forward  (entropy_models.py:645-659 + stf.py:623-626): y, scales, means -> likelihood, y_hat
 *   else y != NULL                encode   (stf.py:717-722): -> symbols_out, indexes_out (if scales), y_hat (if non-NULL)
 *   else symbols_in != NULL       decode   (stf.py:770-772): y_hat = float(symbols_in) + means; + indexes_out if scales
 *   else                          indexes  (stf.py:767): scales -> indexes_out
 * Arithmetic is the NCHW kernels': identical bits for identical inputs.  channels <= 32.
 * narrow != 0: symbols_out is an int16_t buffer and indexes_out a uint8_t buffer (strides still in elements) -- 3 instead
 * of 8 bytes per symbol cross PCIe to the host coder (stf_rans_encode_batch_narrow / stf_rans_decode_batch_u8); *overflow
 * (device int32, zeroed by the caller) is set to 1 when a symbol does not fit int16 or an index does not fit uint8, in
 * which case the caller repeats the step with narrow == 0. */
typedef struct {
  const float *y; int y_ld;
  const float *scales; int scales_ld;
  const float *means; int means_ld;
  const int32_t *symbols_in; int64_t symbols_in_batch_stride;
  int32_t *symbols_out; int32_t *indexes_out; int64_t out_batch_stride;
  float *y_hat; int y_hat_ld;
  float *likelihood; int64_t likelihood_batch_stride;
  int batch, channels; int64_t plane;
  const float *table_host; int levels;   /* HOST scale table (needed when indexes_out != NULL) */
  float scale_bound, lik_bound;
  int ste_round;                         /* forward: y_hat = ((round(t) - t) + t) + mu (ops/ops.py:34) */
  int narrow; int32_t *overflow;         /* int16 symbols_out / uint8 indexes_out + device overflow flag (see above) */
} stf_slice_args;
int stf_slice_step_nhwc(const stf_slice_args *args, void *stream);

/* EntropyModel.quantize(x, "symbols", means) (entropy_models.py:126-150); means may be NULL. */
int stf_quantize_symbols(const float *x, const float *means, int32_t *symbols, int64_t n,
                         void *stream);

/* EntropyModel.quantize(x, "dequantize", means): out = round_half_even(x - means) + means, kept in
 * fp32 (no int32 round trip, like torch.round; entropy_models.py:137-146); means may be NULL. */
int stf_quantize_dequantize(const float *x, const float *means, float *out, int64_t n, void *stream);

/* EntropyModel.dequantize (entropy_models.py:158-165) for one decoded slice:
 *   y_hat[b,c,p] = float(symbols[b * sym_batch_stride + c * plane + p]) + means[b,c,p] */
int stf_dequantize(const int32_t *symbols, int64_t sym_batch_stride, const float *means,
                   float *y_hat, int batch, int channels, int64_t plane, void *stream);

/* GaussianConditional.forward in eval mode fused with the caller's ste_round
 * (entropy_models.py:645-659, 626-643; stf.py:623-626):
 *   y_hat = round_half_even(y - means) + means
 *   v = |y_hat - means|; sigma = max(scales, scale_bound)
 *   lik = max( 0.5*erfc(-(0.5 - v)/sigma/sqrt2) - 0.5*erfc(-(-0.5 - v)/sigma/sqrt2), lik_bound )
 * y is addressed like in stf_gaussian_compress_step; y_hat may be NULL.  16-20 B/element.
 * ste_round != 0: y_hat is the forward value of `ste_round(y - means) + means`, i.e.
 * ((round(t) - t) + t) + means evaluated left to right (ops/ops.py:34), which can differ from
 * round(t) + means in the last bit; the likelihood always uses round(t) + means. */
int stf_gaussian_likelihood(const float *y, int64_t y_batch_stride, const float *scales,
                            const float *means, float *y_hat, float *likelihood, int batch,
                            int channels, int64_t plane, float scale_bound, float lik_bound,
                            int ste_round, void *stream);

/* EntropyBottleneck.forward in eval mode (entropy_models.py:446-489, 400-433) without the two
 * permutes: z is (batch, channels, plane) NCHW.  `params` is a device array of
 * channels * STF_EB_PARAMS floats packed per channel by stf_b200/entropy_models.py:
 *   [softplus(M0) 3][b0 3][tanh(f0) 3] [softplus(M1) 9][b1 3][tanh(f1) 3] ... [softplus(M4) 3][b4 1]
 *   (58 values) then [median][0 pad] -> STF_EB_PARAMS = 60 floats per channel
 *   z_hat = round_half_even(z - median) + median;  lik = max(|sig(s*u) - sig(s*l)|, lik_bound)
 * Optionally also emits int32 symbols = round(z - median) (EntropyBottleneck.compress path,
 * entropy_models.py:508-515).  z_hat / likelihood / symbols may each be NULL.  ste_round as in
 * stf_gaussian_likelihood (stf.py:602-604). */
#define STF_EB_PARAMS 60
#define STF_EB_MEDIAN_SLOT 58
int stf_entropy_bottleneck(const float *z, const float *params, float *z_hat, float *likelihood,
                           int32_t *symbols, int batch, int channels, int64_t plane,
                           float lik_bound, int ste_round, void *stream);

/* ------------------------------------------------------------------------------------------
 * Window-attention path (tensor-core bound).
 *
 * One persistent tcgen05 / TMEM GEMM kernel  Y = epilogue( LN?(gather(X)) . W^T )  with TF32 operands
 * and fp32 accumulation covers qkv, proj, fc1, fc2, PatchMerging and PatchSplit; the producer warps do
 * the window-partition / cyclic-shift / 2x2-merge gathers as index math on cp.async source addresses,
 * the LayerNorm is folded through the GEMM (row statistics gathered on the fly), and the epilogue does
 * bias, q-scaling, exact-erf GELU, residual add and the window-reverse / un-shift / pixel-shuffle scatters.  A second kernel does the per-window softmax(QK^T+B+mask)V.
 * ------------------------------------------------------------------------------------------ */

/* Pack a torch Linear (weight W[N][K] row-major fp32, optional bias[N]) -- and, when the layer is
 * preceded by a LayerNorm over its K inputs, that LayerNorm's gamma[K] / beta[K] -- into the image
 * the GEMM kernel streams with 1-D bulk TMA copies.  `packed` (device) receives
 * stf_packed_linear_floats(N, K) = N*K + 3*N floats:
 *   [N / n_tile][K / 4][n_tile][4]   gamma o W rounded to TF32 (round-to-nearest), n_tile = stf_linear_n_tile(N)
 *   s[N] = sum_k tf32(gamma_k W_nk)    t[N] = sum_k beta_k W_nk + bias_n    b[N] = bias_n
 * so that LN(x).W^T + bias = rstd * (x.(gamma o W)^T - mean * s) + t is evaluated in the GEMM epilogue.
 * All four source pointers are device pointers; bias and (ln_gamma, ln_beta) may be NULL. */
/* Arithmetic of the tensor-core GEMMs.  The reference's matmuls are true fp32 (torch's allow_tf32 is off for
 * matmul), so STF_PREC_FP32 is the parity mode: every operand is split into two TF32 halves x = hi + lo
 * (lo = tf32(x - hi), |x - hi - lo| <= 2^-22 |x|) and the tensor core accumulates hi.hi + lo.hi + hi.lo in
 * fp32 ("3xTF32"); the dropped lo.lo term is below fp32 round-off.  STF_PREC_TF32 is the single-pass mode. */
enum {
  STF_PREC_TF32 = 0, /* operands rounded to TF32 (10-bit mantissa), one MMA per k-step: ~1e-3 relative per GEMM */
  STF_PREC_FP32 = 1  /* 3xTF32 split: fp32-grade results (~1e-6 relative), three MMAs per k-step            */
};
int stf_linear_n_tile(int N);                       /* STF_PREC_TF32 */
int stf_linear_n_tile_prec(int N, int precision);   /* n_tile of the packed image for either precision */
/* Packed size in floats: (1 + precision) * N*K + 3*N  (STF_PREC_FP32 stores a hi and a lo image per k-block). */
int64_t stf_packed_linear_floats(int N, int K, int precision);
int stf_pack_linear(const float *weight, const float *bias, const float *ln_gamma, const float *ln_beta,
                    float *packed, int N, int K, int precision, void *stream);

/* Row gather applied to X before the GEMM (what each of the 128 rows of an M-tile reads). */
enum {
  STF_ROWS_DENSE = 0,   /* row r = X[r, :K]                                                      */
  STF_ROWS_WINDOW = 1,  /* row g (window order) = token after cyclic shift + window partition;    */
                        /* pad tokens (h >= H or w >= W) are zero AFTER the LayerNorm             */
                        /* (stf.py:155-175: norm1, F.pad, torch.roll, window_partition)           */
  STF_ROWS_MERGE = 2    /* row = concat of the 2x2 neighbourhood (x0,x1,x2,x3), zero pad BEFORE   */
                        /* the LayerNorm (stf.py:218-232)                                         */
};
/* Epilogue applied to the fp32 accumulator before the store. */
enum {
  STF_EPI_STORE = 0,        /* Y[r, n] = acc (+bias)                                              */
  STF_EPI_QKV = 1,          /* (+bias), columns < q_cols scaled by q_scale (stf.py:97-100)        */
  STF_EPI_GELU = 2,         /* exact-erf GELU(acc + bias) (stf.py:35-36), stored rounded to TF32  */
  STF_EPI_RESIDUAL = 3,     /* Y[r] = residual[r] + acc + bias  (stf.py:197)                      */
  STF_EPI_WINDOW_RESIDUAL = 4, /* row g -> token t via window_reverse + un-shift, pad rows dropped:*/
                            /* Y[t] = residual[t] + acc + bias (stf.py:181-196)                   */
  STF_EPI_PIXEL_SHUFFLE = 5 /* PatchSplit: feature f of token (h,w) -> token (2h+(f%4)/2,         */
                            /* 2w+f%2), channel f/4 (stf.py:256-259)                              */
};

typedef struct {
  /* problem */
  int M;            /* rows of the GEMM (tokens; for WINDOW: B * nWh * nWw * ws * ws incl. pad) */
  int N;            /* output features */
  int K;            /* input features (for MERGE: 4 * C) */
  const float *x;   /* input activations (token-major, row stride = ldx floats) */
  int ldx;
  const float *w_packed; /* from stf_pack_linear (weights, bias and LayerNorm affine folded in) */
  float *y;         /* output */
  int ldy;
  /* prologue */
  int rows;               /* STF_ROWS_* */
  int has_ln;             /* 1: LayerNorm over the K gathered inputs (packed with ln_gamma / ln_beta) */
  int x_is_tf32;          /* 1: caller guarantees X holds TF32-exact values (low 13 mantissa bits zero), e.g. the */
                          /*    output of STF_EPI_GELU or stf_window_attention: the in-kernel rounding pass is skipped */
  float ln_eps;
  /* epilogue */
  int epilogue;           /* STF_EPI_* */
  const float *residual;  /* for the two residual epilogues (row stride ldy) */
  int q_cols;             /* STF_EPI_QKV */
  float q_scale;
  /* geometry for WINDOW / MERGE / PIXEL_SHUFFLE: feature map of `batch` images H x W tokens */
  int batch, H, W;
  int window;             /* window size ws (WINDOW) */
  int shift;              /* cyclic shift (0 or ws/2) */
  int precision;          /* STF_PREC_*; must match the precision w_packed was packed with */
  int max_ctas;           /* persistent grid size cap (0 = 148), see stf_conv_args */
} stf_linear_args;

/* Fused linear layer on tcgen05 tensor cores.  Replaces, depending on the arguments:
 * norm1+pad+roll+window_partition+qkv (stf.py:155-175,97), proj+window_reverse+roll+residual
 * (stf.py:119,181-196), norm2+fc1+GELU (stf.py:197,35-36), fc2+residual (stf.py:38,197),
 * PatchMerging (stf.py:209-235) and PatchSplit (stf.py:251-260). */
int stf_linear(const stf_linear_args *args, void *stream);

/* Per-window multi-head attention core (stf.py:100-118 / layers/win_attention.py:94-112):
 *   S = q k^T + table[rel_idx(n,m)][head] (+ mask(n,m));  P = softmax_m(S);  O = P v
 * qkv: (num_windows_total * N, 3*C) rows in window order, features [3][heads][d], q already
 * scaled; out: (num_windows_total * N, C) head-major concat.  N = ws*ws in {16, 64},
 * d = C / heads in {16, 24, 32, 40}.  The shifted-window mask ({0,-100}, stf.py:316-334) is
 * computed analytically from the window position when shift > 0: windows are numbered
 * image-major then row-major over the (Hp/ws, Wp/ws) grid.  Independently, `mask` may point to an
 * explicit additive (mask_windows, N, N) fp32 tensor applied as mask[window % mask_windows]
 * (the WindowAttention.forward(x, mask) signature, stf.py:108-110); NULL = none.
 * tf32_out != 0: the output is stored rounded to TF32 (for a STF_PREC_TF32 proj GEMM called with
 * x_is_tf32); 0: full fp32 output. */
int stf_window_attention(const float *qkv, float *out, const float *bias_table, const float *mask,
                         int mask_windows, int64_t num_windows, int C, int heads, int ws, int shift,
                         int Hp, int Wp, int tf32_out, void *stream);

/* The same attention core for 4x4 windows / head_dim 16 (every STF stage) on TOKEN-order operands, contractions on tensor
 * cores: qkv (batch, H, W, 3C) as the dense qkv GEMM (stf_conv2d, ksize 1) leaves it, out (batch, H, W, C) as the proj GEMM
 * reads it.  window_partition / torch.roll / F.pad / window_reverse / un-roll / crop (stf.py:155-196) are the kernel's
 * address arithmetic on per-token bulk (TMA) copies; pad tokens (zero after norm1, stf.py:155-162) take the row pad_qkv[3C]
 * (= qkv bias with the q third pre-scaled; may be NULL when H and W are multiples of 4) and their outputs are dropped.
 * QK^T and PV run as mma.sync.m16n8k8 TF32 tiles, one warp per (window, head) -- 3xTF32 (hi / lo split) when precision is
 * STF_PREC_FP32 --, bias + analytic mask are added on the accumulator fragments, softmax reduces rows with warp shuffles. */
int stf_window_attention_tokens(const float *qkv, float *out, const float *bias_table, const float *pad_qkv, int batch,
                                int H, int W, int C, int heads, int ws, int shift, int precision, void *stream);

/* Fused Swin MLP half-block (stf.py:196-197 with Mlp = stf.py:25-40):
 *   y = x + fc2( GELU( fc1( LayerNorm(x) ) ) )      x, y: (M tokens, C) fp32 rows of x_ld / y_ld floats; hidden = fc1 width
 * One kernel per 128-token tile: the hidden activations stay in shared memory / TMEM (HBM sees x in and y out only).
 * w1_packed / w2_packed: stf_pack_conv images (ksize 1) of fc1 -- with the LayerNorm folded, has_ln -- and fc2, packed for the
 * precision of this call.  C: multiple of 16, <= 192; hidden: multiple of 32.  y may alias x. */
typedef struct {
  const float *x; int x_ld;
  float *y; int y_ld;
  int64_t M; int C, hidden;
  const float *w1_packed, *w2_packed;
  float ln_eps;
  int precision;                        /* STF_PREC_FP32 (3xTF32) or STF_PREC_TF32 */
  int max_ctas;                         /* > 0: cap the persistent grid */
} stf_mlp_args;
int stf_swin_mlp(const stf_mlp_args *args, void *stream);

/* ------------------------------------------------------------------------------------------
 * Convolution stacks either side of the entropy kernels: the five-layer 3x3 cc_mean / cc_scale /
 * lrp stacks of the slice loop (stf.py:510-548 called at :613-633, :706-729, :757-779; cnn.py:89-127), the hyperprior
 * h_a / h_mean_s / h_scale_s incl. subpel_conv3x3 = conv + PixelShuffle(2) (stf.py:472-509, layers/layers.py:47-51) and the
 * 5x5 end_conv (stf.py:466).  One implicit-GEMM tcgen05 kernel, NHWC, operands and results moved by tensor-map TMA:
 *   Y[b,oy,ox,n] = act( sum_{ky,kx,c} X[b, oy*s+ky-p, ox*s+kx-p, c] * W[n,c,ky,kx] + bias[n] ),   p = ksize / 2
 * X is the channel concatenation of n_src (<= 3) NHWC tensors (torch.cat([latent, y_hat_0, ...], 1) is never materialised);
 * the zero padding, the tile halo and ragged edges are TMA out-of-bounds fill; stride 2 is the tensor map's element stride;
 * pixel_shuffle = 2 stores conv channel 4c + 2i + j of pixel (y, x) at channel c of pixel (2y+i, 2x+j) (nn.PixelShuffle).
 * Every output element is one fixed-order K loop (tap-major, source, channel) whatever the batch size or tiling: results are
 * batch-invariant bit for bit, which is what lets a stream encoded in a batch be decoded alone (stf.py:767).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int batch, H, W;            /* input feature map: batch images of H x W pixels, NHWC */
  int n_src;                  /* 1..3 channel-concatenated sources */
  const float *src[3];        /* device pointers, 16-byte aligned */
  int src_channels[3];        /* channels of each source (multiples of 4) */
  int src_ld[3];              /* pixel stride of each source in floats (>= channels; a channel slice of a wider tensor is fine) */
  int N;                      /* conv output channels (multiple of 16) */
  int ksize;                  /* 1, 3 or 5; padding = ksize / 2 */
  int stride;                 /* 1 or 2 */
  const float *w_packed;      /* from stf_pack_conv (same n_src / channels / N / ksize / pixel_shuffle / precision) */
  float *y;                   /* output NHWC: (batch, Ho, Wo, N), or (batch, 2Ho, 2Wo, N/4) with pixel_shuffle */
  int ldy;                    /* output pixel stride in floats */
  int act;                    /* 0: none, 1: exact-erf GELU (nn.GELU), 2: residual + 0.5 * tanh(.) -- the LRP tail */
                              /* `y_hat_slice + 0.5 * torch.tanh(lrp)` (stf.py:631-633), 3: residual + (.) -- the */
                              /* shortcut add behind fc2 / proj (stf.py:196-197); 2 and 3 without pixel shuffle */
  const float *residual;      /* act 2 / 3: NHWC (batch, Ho, Wo, N) tensor, may alias y (updated in place) */
  int res_ld;                 /* its pixel stride in floats */
  int pixel_shuffle;          /* 0 or 2 */
  int has_ln;                 /* ksize 1, one source: a LayerNorm over the source's channels is folded through the GEMM */
  float ln_eps;               /* (row statistics gathered in-kernel while the rows stream through shared memory) */
  int precision;              /* STF_PREC_TF32: one MMA per k-step on the raw fp32 activations (what cuDNN's default TF32 */
                              /* convolutions do); STF_PREC_FP32: 3xTF32 split of both operands, fp32-grade */
  int max_ctas;               /* persistent grid size cap (0 = one CTA per SM, 148).  A caller that runs long single-warp */
                              /* kernels beside this one (the device rANS decoder of other sub-batches) leaves them an SM each: */
                              /* a persistent CTA needs a whole SM's shared memory and would queue behind them */
} stf_conv_args;
/* Floats of the packed weight image: planes * N * Kp + 2 N, Kp = ksize^2 * sum_s ceil32(src_channels[s]). */
int64_t stf_packed_conv_floats(const stf_conv_args *args);
/* Pack an nn.Conv2d weight (N, sum C_s, ksize, ksize) contiguous fp32 (+ bias[N] or NULL), device pointers, into the K-major
 * [N][tap][source][channel padded to 32] image (TF32 hi plane, plus the lo plane for STF_PREC_FP32) followed by two N-vectors
 * t (= bias) and s (= 0), output channels permuted to sub-pixel-major order when pixel_shuffle = 2.
 * The same entry packs an nn.Linear (ksize 1; a token-major (M, K) matrix is the NHWC image (1, 1, M, K)) together with the
 * LayerNorm in front of it (args->has_ln, ln_gamma / ln_beta of K floats): the image holds gamma o W,
 * s = sum_k (gamma o W)[n][k], t = beta . W^T + bias, and stf_conv2d evaluates
 * LN(x) . W^T + bias = rstd * (x . (gamma o W)^T - mean * s) + t in its epilogue.  Output columns < scale_cols are multiplied
 * by row_scale (the q third of a qkv Linear times d^-1/2, stf.py:99); pass 0 / 1.0f otherwise. */
int stf_pack_conv(const stf_conv_args *args, const float *weight, const float *bias, const float *ln_gamma,
                  const float *ln_beta, int scale_cols, float row_scale, float *packed, void *stream);
int stf_conv2d(const stf_conv_args *args, void *stream);
/* Output size of the convolution: Ho = (H + 2p - k) / s + 1 (before any pixel shuffle). */
int stf_conv2d_out_hw(int H, int W, int ksize, int stride, int *Ho, int *Wo);

/* ------------------------------------------------------------------------------------------
 * Element-wise glue around the cuDNN convolution stacks (the callers either side of the hot path).
 * ------------------------------------------------------------------------------------------ */

/* x <- act(x + bias[c]) in place on an NHWC (channels_last) fp32 tensor of n elements, c = element index % channels;
 * act 0 = none, 1 = exact-erf GELU.  Replaces the strided broadcast bias add + GELU launches after every convolution of
 * the hyperprior / cc_mean / cc_scale / lrp stacks (stf.py:472-548).  channels % 4 == 0. */
int stf_bias_act(float *x, const float *bias, int channels, int64_t n, int act, void *stream);

/* PatchEmbed (stf.py:350-381): Conv2d(in_chans -> embed_dim, kernel = stride = patch) on the NCHW image x (batch, in_chans,
 * H, W), zero-padded to a multiple of the patch, + LayerNorm(embed_dim) when ln_gamma / ln_beta are non-NULL, written
 * token-major: tokens (batch * ceil(H/patch) * ceil(W/patch), embed_dim).  weight: (embed_dim, in_chans, patch, patch)
 * contiguous.  embed_dim in {48, 96}, in_chans * patch^2 <= 48.  Fixed summation order per token (batch-invariant). */
int stf_patch_embed(const float *x, const float *weight, const float *bias, const float *ln_gamma, const float *ln_beta,
                    float *tokens, int batch, int in_chans, int H, int W, int patch, int embed_dim, float ln_eps,
                    void *stream);

/* y = LayerNorm(x) over the C <= 768 channels of M token-major rows (PatchEmbed.norm, stf.py:375-379). */
int stf_layernorm_fwd(const float *x, const float *gamma, const float *beta, float *y, int64_t M, int C, float eps,
                      void *stream);

/* ------------------------------------------------------------------------------------------
 * Training step (BASELINE config 5): backward kernels.  The backward GEMMs dX = dY . W are stf_linear calls
 * on the packed transposed weight; weight gradients dW = dY^T . X are plain library GEMMs on the caller's side.
 * Row / window reductions are two-stage and atomic-free: every CTA writes its partial sums to its own slot of
 * `partials`, the caller adds the slots in a fixed order (deterministic gradients).
 * ------------------------------------------------------------------------------------------ */

/* Backward of stf_window_attention for 4x4 and 8x8 windows (autograd of stf.py:100-118 / layers/win_attention.py:94-112):
 *   dqkv (num_windows*16, 3C): [dq * q_scale | dk | dv]   (q in `qkv` is the pre-scaled q the forward consumed;
 *   dq is returned already multiplied by q_scale = d(q_scaled)/d(q), so dqkv is the gradient of the qkv Linear output)
 *   dbias_partials: (stf_attention_bwd_slots(...), (2*ws-1)^2, heads) partial sums of d relative_position_bias_table.
 * stf_attention_bwd_ctas returns the number of partial slots (CTAs) for a problem size. */
int stf_attention_bwd_ctas(int64_t num_windows, int C, int heads, int *windows_per_cta_out);   /* ws = 4 */
int stf_attention_bwd_slots(int64_t num_windows, int C, int heads, int ws);   /* partial slots for ws = 4 or 8 */
int stf_window_attention_bwd(const float *qkv, const float *dout, const float *bias_table, float *dqkv,
                             float *dbias_partials, int64_t num_windows, int C, int heads, int ws, int shift,
                             int Hp, int Wp, float q_scale, void *stream);

/* LayerNorm backward over rows of C <= 768 features (autograd of nn.LayerNorm, stf.py:155,197):
 *   g = dL/d(LN output) (M, C);  dx = LN'(g) + res (res = gradient of the residual path or NULL);
 *   xn (optional) = LN(x) recomputed (the wgrad operand of the Linear behind the norm);
 *   partials: (stf_layernorm_bwd_ctas(M), 2, C) partial sums of (dgamma, dbeta). */
int stf_layernorm_bwd_ctas(int64_t M);
int stf_layernorm_bwd(const float *x, const float *g, const float *gamma, const float *beta, const float *res,
                      float *dx, float *xn, float *partials, int64_t M, int C, float eps, void *stream);

/* Column sums of a row-major (M, C) fp32 matrix (the bias gradients db = sum_rows dY): partials is
 * (stf_colsum_ctas(M), C); C % 4 == 0, C <= 1024. */
int stf_colsum_ctas(int64_t M);
int stf_colsum(const float *a, float *partials, int64_t M, int C, void *stream);

/* Exact-erf GELU backward: dpre = dh * (Phi(pre) + pre * phi(pre)); n % 4 == 0. */
int stf_gelu_bwd(const float *pre, const float *dh, float *dpre, int64_t n, void *stream);

/* GaussianConditional.forward in training mode (entropy_models.py:131-135, 645-659): the likelihood of
 * y + noise (noise = U(-1/2, 1/2) supplied by the caller, NULL = 0) under N(means, max(scales, bound)),
 * floored at lik_bound; and its backward with the LowerBound gradient rule (ops/bound_ops.py:21-27: the
 * gradient passes where x >= bound or grad < 0).  dmean may be NULL. */
int stf_gaussian_likelihood_train(const float *y, const float *scales, const float *means, const float *noise,
                                  float *likelihood, int64_t n, float scale_bound, float lik_bound, void *stream);
int stf_gaussian_likelihood_train_bwd(const float *y, const float *scales, const float *means, const float *noise,
                                      const float *dlik, float *dy, float *dscale, float *dmean, int64_t n,
                                      float scale_bound, float lik_bound, void *stream);

/* ------------------------------------------------------------------------------------------
 * Host-side rANS codec (CPU; replaces compressai.ans, cpp_exts/rans/rans_interface.cpp:99-350,
 * bit-exact streams, no Python lists, LUT symbol search, thread-parallel over streams).
 * ------------------------------------------------------------------------------------------ */

typedef struct stf_rans_table stf_rans_table;

/* Prepare a CDF table set: cdf is (rows, row_stride) int32, sizes[r] = valid entries of row r
 * (cdf_length), offsets[r] = symbol offset.  Precision is 16 bits.  Returns NULL on a malformed
 * table (row not starting at 0, not ending at 65536, or not strictly increasing). */
stf_rans_table *stf_rans_table_create(const int32_t *cdf, int rows, int row_stride,
                                      const int32_t *sizes, const int32_t *offsets);
void stf_rans_table_destroy(stf_rans_table *t);

/* RansEncoder.encode_with_indexes (rans_interface.cpp:193-204): returns the stream length in
 * bytes (written to out[0..)), or STF_E_OVERFLOW / STF_E_ARG.  stf_rans_encode_bound(n) bytes
 * of `out` are always sufficient. */
int64_t stf_rans_encode_bound(int64_t n);
int64_t stf_rans_encode(const stf_rans_table *t, const int32_t *symbols, const int32_t *indexes,
                        int64_t n, uint8_t *out, int64_t out_cap);

/* Encode `count` independent streams on up to `threads` host threads (one image each).
 * out_lens[i] receives the byte length or a negative error. */
int stf_rans_encode_batch(const stf_rans_table *t, int count, const int32_t *const *symbols,
                          const int32_t *const *indexes, const int64_t *n, uint8_t *const *out,
                          const int64_t *out_cap, int64_t *out_lens, int threads);
/* The same on the narrow transfer format of stf_slice_step_nhwc (int16 symbols, uint8 indexes); same bytes out. */
int stf_rans_encode_batch_narrow(const stf_rans_table *t, int count, const int16_t *const *symbols,
                                 const uint8_t *const *indexes, const int64_t *n, uint8_t *const *out,
                                 const int64_t *out_cap, int64_t *out_lens, int threads);

/* RansDecoder (rans_interface.cpp:277-350): set_stream + repeated decode_stream calls. */
typedef struct stf_rans_decoder stf_rans_decoder;
stf_rans_decoder *stf_rans_decoder_create(const uint8_t *stream, int64_t nbytes);
/* Zero-copy variant: `stream` must stay alive and unchanged until stf_rans_decoder_destroy. */
stf_rans_decoder *stf_rans_decoder_create_view(const uint8_t *stream, int64_t nbytes);
void stf_rans_decoder_destroy(stf_rans_decoder *d);
int stf_rans_decode(stf_rans_decoder *d, const stf_rans_table *t, const int32_t *indexes,
                    int64_t n, int32_t *symbols_out);
/* Decode the next n[i] symbols of `count` independent decoders in parallel. */
int stf_rans_decode_batch(stf_rans_decoder *const *d, const stf_rans_table *t, int count,
                          const int32_t *const *indexes, const int64_t *n,
                          int32_t *const *symbols_out, int threads);
/* The same with uint8 indexes (symbols stay int32: what a stream decodes to is not bounded). */
int stf_rans_decode_batch_u8(stf_rans_decoder *const *d, const stf_rans_table *t, int count,
                             const uint8_t *const *indexes, const int64_t *n,
                             int32_t *const *symbols_out, int threads);

/* Device-side decoder for the slice loop of decompress() (stf.py:757-779): RansDecoder.set_stream + decode_stream
 * (rans_interface.cpp:277-350, rans64.h:107-142) for `count` independent streams, lane b of a warp decoding stream b in
 * lockstep -- same integers as stf_rans_decode, bit-identical symbols, no host round trip per slice.
 *   table image: stf_rans_device_table_pack() writes stf_rans_device_table_bytes() bytes into a HOST buffer (row info, 256-bucket
 *                LUTs, 16-bit CDFs); the caller copies it to the device once; the kernel keeps it in shared memory (<= 200 KB).
 *   streams:     all streams back to back as 32-bit words (device), stream b = streams[stream_offsets[b] .. + stream_words[b]).
 *   state:       state_x / state_pos / status [count] (device) carry the decoders between calls; first != 0 initialises them
 *                from the streams (Rans64DecInit).  status[b]: 0, STF_E_STREAM (truncated stream) or STF_E_ARG (bad index).
 *   per call:    n symbols per stream; indexes / symbols_out of stream b at + b * batch_stride (int32, device). */
int64_t stf_rans_device_table_bytes(const stf_rans_table *t);
int stf_rans_device_table_pack(const stf_rans_table *t, void *host_out);
int stf_rans_decode_device(const void *table_dev, int64_t table_bytes, const uint32_t *streams, const int64_t *stream_offsets,
                           const int32_t *stream_words, uint64_t *state_x, uint32_t *state_pos, int32_t *status, int first,
                           const int32_t *indexes, int64_t idx_batch_stride, int32_t *symbols_out, int64_t sym_batch_stride,
                           int count, int64_t n, void *stream);

/* compressai._CXX.pmf_to_quantized_cdf (cpp_exts/ops/ops.cpp:24-81): cdf_out has n+1 entries. */
int stf_pmf_to_quantized_cdf(const float *pmf, int n, int precision, uint32_t *cdf_out);

#ifdef __cplusplus
}
#endif
#endif /* STF_B200_H */
