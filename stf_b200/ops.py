"""Tensor-level wrappers over the C ABI (include/stf_b200.h).  Each function allocates its outputs
with torch, passes raw device pointers + the current CUDA stream through ctypes and raises on a
non-zero status.  No CPU / eager fallback: CPU tensors are rejected.
"""
import ctypes
import math
import os

import numpy as np
import torch

from . import _C, profiler

_f32p = ctypes.POINTER(ctypes.c_float)
_SYNC_LAUNCH = os.environ.get("STF_B200_SYNC_LAUNCH", "0") == "1"


def _launch(kernel, nbytes, fn, *args, flops=0):
    """Call one C-ABI entry point; raise on a non-zero status.  `kernel` / `nbytes` name the CUDA kernel
    and its algorithmic bytes for profiler.capture()."""
    cap = profiler.ACTIVE
    if _SYNC_LAUNCH:     # bring-up: synchronise after every launch so that a device fault names its kernel and arguments
        rc = fn(*args)
        try:
            torch.cuda.synchronize()
        except Exception:
            import sys
            desc = []
            for a in args:
                obj = getattr(a, "_obj", None)     # ctypes.byref(struct)
                if obj is not None and hasattr(obj, "_fields_"):
                    desc.append({f[0]: (list(getattr(obj, f[0])) if hasattr(getattr(obj, f[0]), "__len__") else getattr(obj, f[0]))
                                 for f in obj._fields_})
                else:
                    desc.append(a)
            print(f"[stf_b200] device fault after {kernel} ({fn.__name__}): {desc}", file=sys.stderr, flush=True)
            raise
    elif cap is None:
        rc = fn(*args)
    else:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        cap.add(kernel, nbytes, e0, e1, flops)
    _C.check(rc, fn.__name__)


def _dev(t, name, dtype=torch.float32):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"stf_b200: `{name}` must be a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"stf_b200: `{name}` must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"stf_b200: `{name}` must be contiguous")
    return t


def _host_table(table):
    """fp32 host copy of a scale table + ctypes pointer (kept alive by the returned array)."""
    if isinstance(table, torch.Tensor):
        table = table.detach().cpu().numpy()
    arr = np.ascontiguousarray(table, dtype=np.float32).reshape(-1)
    return arr, arr.ctypes.data_as(_f32p)


# ------------------------------------------------------------------------------------ entropy

def build_indexes(scales, table, scale_bound=0.11):
    """GaussianConditional.build_indexes (reference entropy_models.py:661-666) -> int32."""
    s = _dev(scales.contiguous() if scales.is_cuda else scales, "scales")
    out = torch.empty(s.shape, dtype=torch.int32, device=s.device)
    if s.numel() == 0:
        return out
    arr, p = _host_table(table)
    _launch("compress_step_kernel(build_indexes)", 8 * s.numel(), _C.lib().stf_build_indexes, s.data_ptr(),
            out.data_ptr(), s.numel(), p, arr.size, float(scale_bound), _C.stream())
    return out


def quantize_symbols(x, means=None):
    """EntropyModel.quantize(x, "symbols", means) (reference entropy_models.py:126-150)."""
    x = _dev(x, "inputs")
    m = None
    if means is not None:
        m = _dev(means.expand_as(x).contiguous(), "means")
    out = torch.empty(x.shape, dtype=torch.int32, device=x.device)
    if x.numel() == 0:
        return out
    _launch("quantize_symbols_kernel", (8 if m is None else 12) * x.numel(), _C.lib().stf_quantize_symbols,
            x.data_ptr(), _C.ptr(m), out.data_ptr(), x.numel(), _C.stream())
    return out


def quantize_dequantize(x, means=None):
    """EntropyModel.quantize(x, "dequantize", means): round_half_even(x - means) + means, in fp32."""
    x = _dev(x, "inputs")
    m = None
    if means is not None:
        m = _dev(means.expand_as(x).contiguous(), "means")
    out = torch.empty_like(x)
    if x.numel() == 0:
        return out
    _launch("quantize_dequantize_kernel", (8 if m is None else 12) * x.numel(), _C.lib().stf_quantize_dequantize,
            x.data_ptr(), _C.ptr(m), out.data_ptr(), x.numel(), _C.stream())
    return out


def _slice_geom(y, channel_offset, channels):
    """(B, Ctot, h, w) contiguous tensor -> base pointer of the channel slice, batch stride, plane."""
    B, Ctot = y.shape[0], y.shape[1]
    plane = int(np.prod(y.shape[2:])) if y.dim() > 2 else 1
    if channel_offset < 0 or channel_offset + channels > Ctot:
        raise ValueError("channel slice out of range")
    return y.data_ptr() + 4 * channel_offset * plane, Ctot * plane, plane, B


def gaussian_compress_step(y, channel_offset, scales, means, table, symbols_out, indexes_out, out_offset,
                           scale_bound=0.11, want_y_hat=True):
    """Fused build_indexes + quantize("symbols") + dequantize of one slice (stf.py:717-719).

    y: full (B, M, h, w) latent; the slice is channels [channel_offset, channel_offset + C_s).
    scales / means: dense (B, C_s, h, w).  symbols_out / indexes_out: (B, total) int32 buffers; the
    slice is written at element offset `out_offset` of every row (reference coding order)."""
    y = _dev(y, "y")
    scales, means = _dev(scales, "scales"), _dev(means, "means")
    Cs = scales.shape[1]
    yp, ystride, plane, B = _slice_geom(y, channel_offset, Cs)
    symbols_out, indexes_out = _dev(symbols_out, "symbols", torch.int32), _dev(indexes_out, "indexes", torch.int32)
    total = symbols_out.shape[1]
    if out_offset + Cs * plane > total:
        raise ValueError("symbol buffer too small")
    y_hat = torch.empty_like(scales) if want_y_hat else None
    arr, p = _host_table(table)
    _launch("compress_step_kernel", (24 if want_y_hat else 20) * B * Cs * plane, _C.lib().stf_gaussian_compress_step,
            yp, ystride, scales.data_ptr(), means.data_ptr(), symbols_out.data_ptr() + 4 * out_offset,
            indexes_out.data_ptr() + 4 * out_offset, total, _C.ptr(y_hat), B, Cs, plane, p, arr.size,
            float(scale_bound), _C.stream())
    return y_hat


def dequantize(symbols, sym_offset, means):
    """EntropyModel.dequantize for one decoded slice: symbols (B, total) int32 device buffer, the slice
    starts at element `sym_offset` of each row; means (B, C_s, h, w) -> y_hat like means."""
    symbols = _dev(symbols, "symbols", torch.int32)
    means = _dev(means, "means")
    B, Cs = means.shape[0], means.shape[1]
    plane = means[0, 0].numel()
    out = torch.empty_like(means)
    _launch("dequantize_kernel", 12 * B * Cs * plane, _C.lib().stf_dequantize, symbols.data_ptr() + 4 * sym_offset,
            symbols.shape[1], means.data_ptr(), out.data_ptr(), B, Cs, plane, _C.stream())
    return out


def gaussian_likelihood(y, channel_offset, scales, means, scale_bound=0.11, lik_bound=1e-9, want_y_hat=True,
                        ste_round=False):
    """GaussianConditional.forward (eval) + ste_round of the slice (entropy_models.py:645-659, stf.py:623-626).
    Returns (y_hat or None, likelihood)."""
    y = _dev(y, "y")
    scales = _dev(scales, "scales")
    means = _dev(means, "means") if means is not None else None
    Cs = scales.shape[1]
    yp, ystride, plane, B = _slice_geom(y, channel_offset, Cs)
    lik = torch.empty_like(scales)
    y_hat = torch.empty_like(scales) if want_y_hat else None
    _launch("gaussian_likelihood_kernel", (20 if want_y_hat else 16) * B * Cs * plane,
            _C.lib().stf_gaussian_likelihood, yp, ystride, scales.data_ptr(), _C.ptr(means), _C.ptr(y_hat),
            lik.data_ptr(), B, Cs, plane, float(scale_bound), float(lik_bound), int(bool(ste_round)), _C.stream())
    return y_hat, lik


def _nhwc(t, name, C=None):
    """(B, h, w, C) fp32 CUDA view with unit channel stride and a uniform pixel stride -> (ptr, ld)."""
    if t is None:
        return None, 0
    if not t.is_cuda or t.dtype != torch.float32:
        raise RuntimeError(f"stf_b200: `{name}` must be a CUDA fp32 tensor (there is no CPU path)")
    if t.dim() != 4:
        raise ValueError(f"stf_b200: `{name}` must be a 4-D NHWC tensor, got {tuple(t.shape)}")
    B, h, w, c = t.shape
    # pixel stride from the first dimension that has one (strides of size-1 dimensions are arbitrary)
    ld = t.stride(2) if w > 1 else (t.stride(1) if h > 1 else (t.stride(0) if B > 1 else c))
    ok = (c == 1 or t.stride(3) == 1) and (w == 1 or t.stride(2) == ld) and (h == 1 or t.stride(1) == w * ld) and \
        (B == 1 or t.stride(0) == h * w * ld) and ld >= c and (C is None or c == C)
    if not ok:
        raise ValueError(f"stf_b200: `{name}` must be an NHWC view with a uniform pixel stride, got {tuple(t.shape)} / {t.stride()}")
    return t.data_ptr(), ld


def slice_step_nhwc(*, y=None, scales=None, means=None, symbols_in=None, sym_in_offset=0, symbols_out=None,
                    indexes_out=None, out_offset=0, y_hat=None, likelihood=None, lik_offset=0, table=None,
                    scale_bound=0.11, lik_bound=1e-9, ste_round=False, overflow=None):
    """One slice step on NHWC operands (stf_slice_step_nhwc in include/stf_b200.h).

    y / scales / means / y_hat: (B, h, w, C<=32) NHWC views (channel slices of wider tensors are fine; y_hat is written).
    symbols_in / symbols_out / indexes_out: (B, total) int32 buffers in coding order, the slice at element offset
    sym_in_offset / out_offset of every row.  likelihood: (B, M, h, w) NCHW tensor, the slice at channel lik_offset.
    The step (forward / encode / decode / indexes) follows from which arguments are given.
    Narrow outputs: symbols_out int16 and / or indexes_out uint8 (both narrow when both are given) with `overflow`, a
    zero-initialised int32 CUDA tensor of one element that the kernel sets when a value does not fit."""
    ref = next(t for t in (y, scales, means, y_hat) if t is not None)
    B, h, w, C = ref.shape
    plane = h * w
    a = _C.SliceArgs()
    a.y, a.y_ld = _nhwc(y, "y", C)
    a.scales, a.scales_ld = _nhwc(scales, "scales", C)
    a.means, a.means_ld = _nhwc(means, "means", C)
    a.y_hat, a.y_hat_ld = _nhwc(y_hat, "y_hat", C)
    if symbols_in is not None:
        symbols_in = _dev(symbols_in, "symbols_in", torch.int32)
        a.symbols_in, a.symbols_in_batch_stride = symbols_in.data_ptr() + 4 * sym_in_offset, symbols_in.shape[1]
    total = None
    outs = [(n_, b_, w_, nd_) for n_, b_, w_, nd_ in (("symbols_out", symbols_out, torch.int32, torch.int16),
                                                       ("indexes_out", indexes_out, torch.int32, torch.uint8)) if b_ is not None]
    narrow = bool(outs) and all(b_.dtype == nd_ for _, b_, _, nd_ in outs)
    for name, buf, wide_dt, narrow_dt in outs:
        buf = _dev(buf, name, narrow_dt if narrow else wide_dt)
        if total is not None and buf.shape[1] != total:
            raise ValueError("symbols_out and indexes_out must have the same row length")
        total = buf.shape[1]
        if out_offset + C * plane > total:
            raise ValueError(f"{name} buffer too small")
        setattr(a, name, buf.data_ptr() + buf.element_size() * out_offset)
    a.out_batch_stride = total or 0
    if narrow:
        if overflow is None:
            raise ValueError("narrow symbols_out / indexes_out need an `overflow` flag tensor")
        a.narrow, a.overflow = 1, _dev(overflow, "overflow", torch.int32).data_ptr()
    if likelihood is not None:
        likelihood = _dev(likelihood, "likelihood")
        a.likelihood = likelihood.data_ptr() + 4 * lik_offset * plane
        a.likelihood_batch_stride = likelihood.shape[1] * plane
    a.batch, a.channels, a.plane = B, C, plane
    keep = None
    if indexes_out is not None:
        keep, a.table_host = _host_table(table)
        a.levels = keep.size
    a.scale_bound, a.lik_bound, a.ste_round = float(scale_bound), float(lik_bound), int(bool(ste_round))
    streams = sum(t is not None for t in (y, scales, means, symbols_in, symbols_out, indexes_out, y_hat, likelihood))
    _launch("slice_step_nhwc_kernel", 4 * streams * B * C * plane, _C.lib().stf_slice_step_nhwc, ctypes.byref(a), _C.stream())


def entropy_bottleneck(z, params, lik_bound=1e-9, want_z_hat=True, want_lik=True, want_symbols=False,
                       ste_round=False):
    """EntropyBottleneck.forward (eval) without the permutes; params: (C, 60) packed (see header)."""
    z = _dev(z, "z")
    params = _dev(params, "params")
    B, C = z.shape[0], z.shape[1]
    plane = z[0, 0].numel()
    z_hat = torch.empty_like(z) if want_z_hat else None
    lik = torch.empty_like(z) if want_lik else None
    sym = torch.empty(z.shape, dtype=torch.int32, device=z.device) if want_symbols else None
    nout = int(want_z_hat) + int(want_lik) + int(want_symbols)
    _launch("entropy_bottleneck_kernel", 4 * (1 + nout) * z.numel(), _C.lib().stf_entropy_bottleneck, z.data_ptr(),
            params.data_ptr(), _C.ptr(z_hat), _C.ptr(lik), _C.ptr(sym), B, C, plane, float(lik_bound),
            int(bool(ste_round)), _C.stream())
    return z_hat, lik, sym


# ------------------------------------------------------------------------------------ tensor-core linear

# Arithmetic of the tensor-core GEMMs (include/stf_b200.h STF_PREC_*).  "fp32" (default) is the parity mode: the
# reference's matmuls are true fp32, so operands are split hi + lo and multiplied as 3xTF32 (fp32-grade results).
# "tf32" is the single-pass fast mode (~1e-3 relative per GEMM).  Select with STF_B200_PRECISION or set_precision().
_PRECISIONS = {"fp32": _C.PREC_FP32, "tf32": _C.PREC_TF32}
_precision = _PRECISIONS[os.environ.get("STF_B200_PRECISION", "fp32").lower()]


def set_precision(name):
    """'fp32' (3xTF32, parity with the reference's fp32 matmuls) or 'tf32' (single pass).  Returns the old name."""
    global _precision
    old = precision()
    _precision = _PRECISIONS[name.lower()]
    return old


def precision():
    return "fp32" if _precision == _C.PREC_FP32 else "tf32"


def precision_code():
    return _precision


class PackedLinear:
    """A torch Linear (weight (N, K), optional bias) -- and the LayerNorm in front of it, if any -- packed for
    the tcgen05 kernel (stf_pack_linear: TF32 tile image(s) + the three LayerNorm-folding vectors)."""

    def __init__(self, weight, bias=None, ln=None, prec=None):
        self.precision = _precision if prec is None else int(prec)
        w = _dev(weight.detach().contiguous(), "weight")
        self.N, self.K = w.shape
        b = None if bias is None else _dev(bias.detach().contiguous(), "bias")
        g = be = None
        self.ln_eps = 0.0
        if ln is not None:
            g, be = _dev(ln[0].detach().contiguous(), "ln.weight"), _dev(ln[1].detach().contiguous(), "ln.bias")
            self.ln_eps = float(ln[2])
            if g.numel() != self.K or be.numel() != self.K:
                raise ValueError("LayerNorm width does not match the Linear's input features")
        self.has_ln = ln is not None
        n = int(_C.lib().stf_packed_linear_floats(self.N, self.K, self.precision))
        self.packed = torch.empty(n, dtype=torch.float32, device=w.device)
        _launch("pack_weight_kernel", 8 * w.numel(), _C.lib().stf_pack_linear, w.data_ptr(), _C.ptr(b), _C.ptr(g),
                _C.ptr(be), self.packed.data_ptr(), self.N, self.K, self.precision, _C.stream())


def linear(x, lin, *, M=None, rows=_C.ROWS_DENSE, epilogue=_C.EPI_STORE, residual=None, q_cols=0, q_scale=1.0,
           geom=None, out=None, out_rows=None, out_cols=None, x_is_tf32=False):
    """Y = epilogue(LN?(gather(X)) . W^T)   (stf_linear in include/stf_b200.h).

    x: (rows_in, ldx) fp32; lin: PackedLinear (carries bias and the optional LayerNorm);
    geom: (batch, H, W, window, shift) for the WINDOW / MERGE / PIXEL_SHUFFLE index math."""
    x = _dev(x, "x")
    x2 = x.reshape(-1, x.shape[-1])
    if x2.shape[1] * (4 if rows == _C.ROWS_MERGE else 1) != lin.K:
        raise ValueError(f"stf_linear: input features {x2.shape[1]} do not match the weight's K={lin.K}")
    a = _C.LinearArgs()
    a.M = x2.shape[0] if M is None else int(M)
    a.N, a.K = lin.N, lin.K
    a.x, a.ldx = x2.data_ptr(), x2.shape[1]
    a.w_packed = lin.packed.data_ptr()
    if out is None:
        out = torch.empty((a.M if out_rows is None else out_rows, lin.N if out_cols is None else out_cols),
                          dtype=torch.float32, device=x.device)
    out = _dev(out, "out")
    a.y, a.ldy = out.data_ptr(), out.shape[-1]
    a.rows = rows
    a.has_ln, a.ln_eps = int(lin.has_ln), lin.ln_eps
    a.x_is_tf32 = int(bool(x_is_tf32)) if lin.precision == _C.PREC_TF32 else 0
    a.precision = lin.precision
    a.max_ctas = MAX_CTAS
    a.epilogue = epilogue
    if residual is not None:
        residual = _dev(residual, "residual")
        a.residual = residual.data_ptr()
    a.q_cols, a.q_scale = int(q_cols), float(q_scale)
    if geom is not None:
        a.batch, a.H, a.W, a.window, a.shift = (int(v) for v in geom)
    # algorithmic bytes: activations in + weights + outputs (+ residual read)
    nbytes = 4 * (a.M * a.K + a.N * a.K + a.M * a.N * (2 if residual is not None else 1))
    _launch("linear_tf32_kernel", nbytes, _C.lib().stf_linear, ctypes.byref(a), _C.stream(), flops=2 * a.M * a.N * a.K)
    return out


def window_attention_core(qkv, bias_table, num_windows, C, heads, ws, shift, Hp=0, Wp=0, mask=None, tf32_out=None):
    """softmax(q k^T + B + mask) v per (window, head); qkv (num_windows*ws*ws, 3C), q pre-scaled.
    tf32_out (default: the TF32 precision mode is active): store the output rounded to TF32 for the proj GEMM."""
    if tf32_out is None:
        tf32_out = _precision == _C.PREC_TF32
    qkv = _dev(qkv, "qkv")
    bias_table = _dev(bias_table.detach(), "relative_position_bias_table")
    out = torch.empty((qkv.shape[0], C), dtype=torch.float32, device=qkv.device)
    mw = 0
    if mask is not None:
        mask = _dev(mask.contiguous(), "mask")
        mw = mask.shape[0]
    _launch("window_attention_kernel", 4 * qkv.shape[0] * 4 * C, _C.lib().stf_window_attention, qkv.data_ptr(),
            out.data_ptr(), bias_table.data_ptr(), _C.ptr(mask), mw, int(num_windows), C, heads, ws, shift, Hp, Wp,
            int(bool(tf32_out)), _C.stream())
    return out


# ------------------------------------------------------------------------------------ convolution stacks

def conv_precision_code():
    """Arithmetic of the conv kernel.  Default = the library's GEMM precision for the strict "fp32" parity runs only when
    STF_B200_CONV_PRECISION=fp32 (or set_conv_precision("fp32")): otherwise single-pass TF32 on the raw fp32 activations,
    which is what the reference's convolutions do on a GPU (torch's cudnn.allow_tf32 default)."""
    return _conv_precision


def set_conv_precision(name):
    global _conv_precision
    old = "fp32" if _conv_precision == _C.PREC_FP32 else "tf32"
    _conv_precision = _PRECISIONS[name.lower()]
    return old


_conv_precision = _PRECISIONS[os.environ.get("STF_B200_CONV_PRECISION", "tf32").lower()]
# Linear layers of the Swin blocks on the TMA / tcgen05 GEMM engine (stf_conv2d with ksize 1) instead of the cp.async-fed
# stf_linear kernel; "0" keeps the older kernel (A/B measurements).
GEMM_ENGINE = os.environ.get("STF_B200_GEMM_ENGINE", "1") != "0"
# Grid cap of the persistent kernels (0 = 148, one CTA per SM).  models.decompress() lowers it while device rANS decoders of
# other sub-batches may be running: each of those is ONE warp on one SM for milliseconds, and a persistent CTA (a whole SM's
# shared memory) assigned to that SM would wait for it -- head-of-line blocking of the entire kernel.
FUSED_MLP = os.environ.get("STF_B200_FUSED_MLP", "1") != "0"
FUSED_MLP_MAX_C = int(os.environ.get("STF_B200_FUSED_MLP_MAX_C", "96"))
MAX_CTAS = 0
NUM_SMS = 148


class PackedConv:
    """An nn.Conv2d weight (N, sum C_s, k, k) + bias packed for stf_conv2d: K-major [N][tap][source][channel padded to 32]
    TF32 image(s) + bias, output channels in sub-pixel-major order when a PixelShuffle(2) follows."""

    def __init__(self, weight, bias, src_channels=None, stride=1, pixel_shuffle=0, prec=None, ln=None, row_scale=None):
        """weight: conv (N, C, k, k) or Linear (N, K); ln = (gamma, beta, eps) folds the LayerNorm in front of a Linear
        through the GEMM; row_scale = (cols, factor) multiplies the first `cols` output features (q * d^-1/2)."""
        self.precision = _conv_precision if prec is None else int(prec)
        w = _dev(weight.detach().contiguous(), "weight")      # (N, C, k, k) in NCHW-contiguous order for the packer
        if w.dim() == 2:
            w = w.reshape(w.shape[0], w.shape[1], 1, 1)
        self.N, ctot, self.ksize, k2 = w.shape
        if src_channels is None:
            src_channels = (ctot,)
        if k2 != self.ksize or sum(src_channels) != ctot:
            raise ValueError("stf_conv2d: weight shape does not match the sources' channels")
        self.src_channels = tuple(int(c) for c in src_channels)
        self.stride, self.pixel_shuffle = int(stride), int(pixel_shuffle)
        self.has_ln, self.ln_eps = ln is not None, 0.0 if ln is None else float(ln[2])
        b = None if bias is None else _dev(bias.detach().contiguous(), "bias")
        g = be = None
        if ln is not None:
            g, be = _dev(ln[0].detach().contiguous(), "ln.weight"), _dev(ln[1].detach().contiguous(), "ln.bias")
            if g.numel() != ctot or be.numel() != ctot or self.ksize != 1 or len(self.src_channels) != 1:
                raise ValueError("stf_conv2d: a folded LayerNorm needs a Linear (ksize 1, one source) of matching width")
        cols, factor = (0, 1.0) if row_scale is None else (int(row_scale[0]), float(row_scale[1]))
        a = self.args()
        n = int(_C.lib().stf_packed_conv_floats(ctypes.byref(a)))
        if n < 0:
            _C.check(n, "stf_packed_conv_floats")
        self.packed = torch.empty(n, dtype=torch.float32, device=w.device)
        _launch("pack_conv_kernel", 8 * w.numel(), _C.lib().stf_pack_conv, ctypes.byref(a), w.data_ptr(), _C.ptr(b),
                _C.ptr(g), _C.ptr(be), cols, factor, self.packed.data_ptr(), _C.stream())

    def args(self):
        a = _C.ConvArgs()
        a.n_src = len(self.src_channels)
        for i, c in enumerate(self.src_channels):
            a.src_channels[i] = c
        a.N, a.ksize, a.stride = self.N, self.ksize, self.stride
        a.pixel_shuffle, a.precision = self.pixel_shuffle, self.precision
        a.has_ln, a.ln_eps = int(self.has_ln), self.ln_eps
        return a


def conv_out_hw(H, W, ksize, stride):
    ho, wo = ctypes.c_int(), ctypes.c_int()
    _C.check(_C.lib().stf_conv2d_out_hw(int(H), int(W), int(ksize), int(stride), ctypes.byref(ho), ctypes.byref(wo)),
             "stf_conv2d_out_hw")
    return ho.value, wo.value


def conv2d(srcs, pc, act=False, out=None, residual=None):
    """act(conv2d(cat(srcs, channel), W) + bias) on NHWC tensors (stf_conv2d in include/stf_b200.h).

    srcs: list of (B, H, W, C_s) fp32 CUDA tensors whose last dimension is dense (a channel slice of a wider NHWC tensor
    is fine: only the pixel stride must be uniform); pc: PackedConv; out: optional (B, Ho, Wo, N) / (B, 2Ho, 2Wo, N/4)
    destination view (same stride rules).  act: False / True ("gelu") / "lrp" (out = residual + 0.5 * tanh(conv + bias), the
    residual defaulting to `out` itself, updated in place: stf.py:631-633).  Returns the NHWC output."""
    a = pc.args()
    B, H, W = srcs[0].shape[:3]
    if len(srcs) != a.n_src:
        raise ValueError("stf_conv2d: number of sources does not match the packed weight")
    for i, t in enumerate(srcs):
        if not t.is_cuda or t.dtype != torch.float32:
            raise RuntimeError("stf_conv2d: sources must be CUDA fp32 tensors (there is no CPU path)")
        if t.dim() != 4 or tuple(t.shape[:3]) != (B, H, W) or t.shape[3] != pc.src_channels[i]:
            raise ValueError(f"stf_conv2d: source {i} has shape {tuple(t.shape)}")
        a.src[i], a.src_ld[i] = _nhwc(t, f"source {i}")
    a.batch, a.H, a.W = B, H, W
    Ho, Wo = conv_out_hw(H, W, pc.ksize, pc.stride)
    r = 2 if pc.pixel_shuffle else 1
    shape = (B, Ho * r, Wo * r, pc.N // (r * r))
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=srcs[0].device)
    elif tuple(out.shape) != shape:
        raise ValueError(f"stf_conv2d: `out` must be an NHWC view of shape {shape}")
    a.w_packed = pc.packed.data_ptr()
    a.max_ctas = MAX_CTAS
    a.y, a.ldy = _nhwc(out, "out")
    a.act = {False: 0, True: 1, None: 0, "gelu": 1, "lrp": 2, "residual": 3}[act]
    if a.act >= 2:      # out <- residual + 0.5 * tanh(conv + bias) / residual + conv + bias; default residual = out's content
        res = out if residual is None else residual
        a.residual, a.res_ld = _nhwc(res, "residual", shape[3])
    ctot = sum(pc.src_channels)
    # algorithmic bytes: activations in + weights + outputs (+ the residual read)
    nbytes = 4 * (B * H * W * ctot + pc.N * ctot * pc.ksize ** 2 + B * Ho * Wo * pc.N * (2 if a.act >= 2 else 1))
    _launch("conv_tf32_kernel:linear" if pc.ksize == 1 else "conv_tf32_kernel:conv", nbytes, _C.lib().stf_conv2d,
            ctypes.byref(a), _C.stream(), flops=2 * B * Ho * Wo * pc.N * ctot * pc.ksize ** 2)
    return out


def gemm(x, pc, act=False, residual=None, out=None):
    """Linear layer on the TMA / tcgen05 GEMM engine: act(LN?(x) . W^T + bias) for token-major x (M, K) -> (M, N).
    The matrix is handed to stf_conv2d as the NHWC image (1, 1, M, K) (ksize 1): 128-row tiles, LayerNorm folded through the
    GEMM, act in {False, True / "gelu", "residual"} with residual (M, N)."""
    if x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("stf_b200.gemm: x must be a token-major (M, K) matrix")
    M = x.shape[0]
    x4 = x.as_strided((1, 1, M, x.shape[1]), (M * x.stride(0), M * x.stride(0), x.stride(0), 1))
    o4 = r4 = None
    if out is not None:
        o4 = out.as_strided((1, 1, M, out.shape[1]), (M * out.stride(0), M * out.stride(0), out.stride(0), 1))
    if residual is not None:
        r4 = residual.as_strided((1, 1, M, residual.shape[1]),
                                 (M * residual.stride(0), M * residual.stride(0), residual.stride(0), 1))
    y = conv2d([x4], pc, act=act, out=o4, residual=r4)
    return y.reshape(M, -1) if out is None else out


def swin_mlp(x, pc1, pc2, out=None):
    """x + fc2(GELU(fc1(LayerNorm(x)))) for token-major x (M, C) in ONE kernel (stf_swin_mlp): the hidden activations never
    reach HBM.  pc1: PackedConv of fc1 with the LayerNorm folded (ln=...), pc2: PackedConv of fc2, both in the current
    precision.  out: optional (M, C) destination (may be x)."""
    if x.dim() != 2 or x.stride(1) != 1 or not x.is_cuda or x.dtype != torch.float32:
        raise ValueError("stf_b200.swin_mlp: x must be a token-major CUDA fp32 (M, C) matrix")
    M, C = x.shape
    if not pc1.has_ln or pc1.ksize != 1 or pc2.ksize != 1 or pc1.src_channels != (C,) or pc2.N != C or \
            pc2.src_channels != (pc1.N,) or pc1.precision != pc2.precision:
        raise ValueError("stf_b200.swin_mlp: fc1 must carry the folded LayerNorm and fc2 must map hidden -> C")
    if out is None:
        out = torch.empty((M, C), dtype=torch.float32, device=x.device)
    a = _C.MlpArgs()
    a.x, a.x_ld, a.y, a.y_ld = x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0)
    a.M, a.C, a.hidden = M, C, pc1.N
    a.w1_packed, a.w2_packed = pc1.packed.data_ptr(), pc2.packed.data_ptr()
    a.ln_eps, a.precision, a.max_ctas = pc1.ln_eps, pc1.precision, MAX_CTAS
    _launch("swin_mlp_kernel", 8 * M * C, _C.lib().stf_swin_mlp, ctypes.byref(a), _C.stream(), flops=4 * M * C * pc1.N)
    return out


def mlp_fusable(C, hidden):
    """Shapes stf_swin_mlp takes and where it pays (tools/bench_mlp.py): C = 48 in both GEMM modes (2.90 vs 3.09 ms per
    6.3 M tokens in the 3xTF32 mode, 2.20 vs 2.83 ms single-pass), C = 96 in the single-pass mode only (1.24 vs 1.47 ms; in
    the 3xTF32 mode the two-launch form wins there, 1.87 vs 2.19 ms)."""
    return FUSED_MLP and hidden == 4 * C and (C == 48 or (C == 96 and _precision == 0)) and C <= FUSED_MLP_MAX_C


def window_attention_tokens(qkv, bias_table, pad_qkv, B, H, W, C, heads, ws, shift):
    """Attention core + partition / shift / pad / reverse on token-order qkv (B*H*W, 3C) -> (B*H*W, C)
    (stf_window_attention_tokens in include/stf_b200.h); 4x4 windows, head_dim 16."""
    qkv = _dev(qkv, "qkv")
    bias_table = _dev(bias_table.detach(), "relative_position_bias_table")
    out = torch.empty((qkv.shape[0], C), dtype=torch.float32, device=qkv.device)
    if pad_qkv is not None:
        pad_qkv = _dev(pad_qkv, "pad_qkv")
    _launch("window_attention_tok_kernel", 4 * qkv.shape[0] * 4 * C, _C.lib().stf_window_attention_tokens, qkv.data_ptr(),
            out.data_ptr(), bias_table.data_ptr(), _C.ptr(pad_qkv), int(B), int(H), int(W), int(C), int(heads), int(ws),
            int(shift), _precision, _C.stream(), flops=4 * qkv.shape[0] * ws * ws * C)
    return out


def bias_act_(x, bias, gelu):
    """x <- act(x + bias[c]) in place; x: (B, C, H, W) in channels_last memory format (dense NHWC)."""
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 4 or not x.is_contiguous(memory_format=torch.channels_last):
        raise ValueError("stf_b200.bias_act_: expected a CUDA fp32 (B, C, H, W) tensor in channels_last format")
    bias = _dev(bias.detach(), "bias")
    _launch("bias_act_kernel", 8 * x.numel(), _C.lib().stf_bias_act, x.data_ptr(), bias.data_ptr(), x.shape[1], x.numel(),
            1 if gelu else 0, _C.stream())
    return x


def patch_embed(x, weight, bias, ln_weight, ln_bias, patch, eps):
    """PatchEmbed (stf.py:350-381) on the NCHW image -> (tokens (B * Wh * Ww, E), Wh, Ww)."""
    x = _dev(x.contiguous(), "x")
    B, Cin, H, W = x.shape
    w = _dev(weight.detach().contiguous(), "weight")
    E = w.shape[0]
    Wh, Ww = -(-H // patch), -(-W // patch)
    out = torch.empty((B * Wh * Ww, E), dtype=torch.float32, device=x.device)
    g = None if ln_weight is None else _dev(ln_weight.detach().contiguous(), "norm.weight")
    be = None if ln_bias is None else _dev(ln_bias.detach().contiguous(), "norm.bias")
    b = None if bias is None else _dev(bias.detach().contiguous(), "bias")
    _launch("patch_embed_kernel", 4 * (x.numel() + out.numel()), _C.lib().stf_patch_embed, x.data_ptr(), w.data_ptr(),
            _C.ptr(b), _C.ptr(g), _C.ptr(be), out.data_ptr(), B, Cin, H, W, int(patch), E, float(eps), _C.stream())
    return out, Wh, Ww


def layernorm(x, weight, bias, eps):
    """Token-major LayerNorm: x (M, C) contiguous -> (M, C)."""
    x = _dev(x, "x")
    out = torch.empty_like(x)
    _launch("layernorm_fwd_kernel", 8 * x.numel(), _C.lib().stf_layernorm_fwd, x.data_ptr(),
            _dev(weight.detach(), "weight").data_ptr(), _dev(bias.detach(), "bias").data_ptr(), out.data_ptr(), x.shape[0],
            x.shape[1], float(eps), _C.stream())
    return out


_replayed = 0


def count_replayed_launches(n):
    """Kernels of this library re-issued by a CUDA-graph replay (the C-side counter only sees captures)."""
    global _replayed
    _replayed += int(n)


def launch_count():
    """stf_b200 kernels launched by this process: direct C-ABI launches + launches replayed inside CUDA graphs."""
    return int(_C.lib().stf_launch_count()) + _replayed


def ceil_to(v, m):
    return int(math.ceil(v / m)) * m
