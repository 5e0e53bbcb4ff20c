"""Autograd functions of the training step (BASELINE config 5; SURVEY.md section 8 rows a1/a4/a5 backward,
a11 "noise", a13, a15, a16).

Forward = the same fused kernels as inference (stf_linear / stf_window_attention / entropy kernels).
Backward:
  * dX = dY . W              stf_linear on the packed TRANSPOSED weight; the window gather / scatter of the forward
                             pass swap roles (a forward row gather is a backward row scatter and vice versa)
  * attention core           stf_window_attention_bwd   (csrc/train_kernels.cu)
  * LayerNorm (+ residual)   stf_layernorm_bwd          (also re-emits LN(x), the wgrad operand)
  * GELU                     stf_gelu_bwd on the pre-activation, which is RECOMPUTED by one more fc1 GEMM
                             (the forward keeps only GELU(h): 4C floats per token instead of 8C)
  * dW = dY^T . X            plain library GEMM (torch -> cuBLAS; TF32 tensor cores by default, see _wgrad_precision)
  * db = column sums         stf_colsum (two-stage, deterministic)
  * Gaussian likelihood      stf_gaussian_likelihood_train{,_bwd} with the LowerBound gradient rule
Reductions are two-stage and atomic-free (deterministic gradients).  Feature maps that are not a multiple of the window go
through the same zero-pad index math as the forward kernels (pad tokens are zero after norm1 and take part as keys).
"""
import torch

from . import _C, ops


_ROW_INDEX_CACHE = {}


def window_row_index(B, H, W, ws, shift, device):
    """Cached per geometry (the same few shapes recur every step)."""
    key = (B, H, W, ws, shift, str(device))
    if key not in _ROW_INDEX_CACHE:
        if len(_ROW_INDEX_CACHE) >= 64:
            _ROW_INDEX_CACHE.clear()
        _ROW_INDEX_CACHE[key] = _window_row_index(B, H, W, ws, shift, device)
    return _ROW_INDEX_CACHE[key]


def _window_row_index(B, H, W, ws, shift, device):
    """Token index of every row of the window-ordered layout (F.pad + roll(-shift) + window_partition, stf.py:158-171), -1
    for the zero-pad tokens: used to bring token-ordered tensors into window order for the weight-gradient GEMMs."""
    Hp, Wp = ops.ceil_to(H, ws), ops.ceil_to(W, ws)
    idx = torch.full((B, Hp, Wp), -1, device=device, dtype=torch.int64)
    idx[:, :H, :W] = torch.arange(B * H * W, device=device, dtype=torch.int64).view(B, H, W)
    if shift:
        idx = torch.roll(idx, shifts=(-shift, -shift), dims=(1, 2))
    idx = idx.view(B, Hp // ws, ws, Wp // ws, ws).permute(0, 1, 3, 2, 4)
    return idx.reshape(-1)


def gather_rows(t, idx, has_pad):
    """t[idx] with all-zero rows where idx < 0 (pad tokens are zero after norm1 and their outputs are cropped).  `has_pad`
    comes from the geometry on the host: inspecting idx on the device would synchronise the launch queue in every backward."""
    if not has_pad:
        return t.index_select(0, idx)
    return t.index_select(0, idx.clamp_min(0)) * (idx >= 0).unsqueeze(1).to(t.dtype)


def layernorm_bwd(x, g, gamma, beta, eps, res=None, want_xn=True):
    """-> (dx [+ res], LN(x) or None, dgamma, dbeta)."""
    M, C = x.shape
    dx = torch.empty_like(x)
    xn = torch.empty_like(x) if want_xn else None
    L = _C.lib()
    ctas = int(L.stf_layernorm_bwd_ctas(M))
    part = torch.empty((ctas, 2, C), dtype=torch.float32, device=x.device)
    ops._launch("layernorm_bwd_kernel", 4 * M * C * (4 + int(want_xn) + int(res is not None)), L.stf_layernorm_bwd,
                x.data_ptr(), g.data_ptr(), gamma.data_ptr(), _C.ptr(beta), _C.ptr(res), dx.data_ptr(), _C.ptr(xn),
                part.data_ptr(), M, C, float(eps), _C.stream())
    s = part.sum(0)
    return dx, xn, s[0], s[1]


def colsum(a):
    """Column sums of a contiguous (M, C) matrix (bias gradients), deterministic two-stage reduction."""
    M, C = a.shape
    L = _C.lib()
    if C % 4 or C > 1024:
        return a.sum(0)
    ctas = int(L.stf_colsum_ctas(M))
    part = torch.empty((ctas, C), dtype=torch.float32, device=a.device)
    ops._launch("colsum_kernel", 4 * M * C, L.stf_colsum, a.data_ptr(), part.data_ptr(), M, C, _C.stream())
    return part.sum(0)


def wgrad(dy, x):
    """dW = dY^T . X, a plain library GEMM (cuBLAS through torch).  Its arithmetic follows the library's precision mode
    (ops.precision()): "fp32" -> 3xTF32-grade via a manual hi/lo split of both operands on TF32 tensor cores (three GEMMs,
    fp32 accumulate; the dropped lo.lo term is below fp32 round-off), "tf32" -> one TF32 GEMM.  Each product runs under
    a scoped, lock-protected TF32 override that always restores the caller's setting (STF_B200_WGRAD_FP32=1: plain fp32
    SIMT SGEMM instead, ~8x slower)."""
    import os
    if os.environ.get("STF_B200_WGRAD_FP32", "0") == "1":
        with _tf32_matmul(False):
            return dy.t().mm(x)
    with _tf32_matmul(True):
        if ops.precision() != "fp32":
            return dy.t().mm(x)
        # split in the operands' own layout (dy^T is a view: cuBLAS takes the transposed operand without a copy)
        a_hi, b_hi = _tf32_trunc(dy), _tf32_trunc(x)
        a_lo, b_lo = dy - a_hi, x - b_hi         # exact in fp32; the GEMM reads its upper 11 bits: x = hi + lo to 2^-22
        out = a_hi.t().mm(b_hi)
        out.addmm_(a_lo.t(), b_hi)
        out.addmm_(a_hi.t(), b_lo)
        return out


def _tf32_trunc(t):
    """Keep the upper 19 bits (sign, exponent, 10 mantissa bits): the value a TF32 tensor core reads from an fp32 word."""
    return (t.contiguous().view(torch.int32) & -8192).view(torch.float32)


class _tf32_matmul:
    """Scoped TF32 switch for torch.mm with a re-entrancy lock: autograd runs backward nodes of one device on one
    worker thread, but user code may call matmuls from other threads; the lock keeps the flag flip + GEMM atomic with
    respect to other wgrad calls and the old value is always restored."""
    import threading
    _lock = threading.RLock()

    def __init__(self, on):
        self.on = on

    def __enter__(self):
        self._lock.acquire()
        self.old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = self.on

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.old
        self._lock.release()
        return False


def gelu_bwd(pre, dh):
    out = torch.empty_like(pre)
    ops._launch("gelu_bwd_kernel", 12 * pre.numel(), _C.lib().stf_gelu_bwd, pre.data_ptr(), dh.data_ptr(),
                out.data_ptr(), pre.numel(), _C.stream())
    return out


def attention_bwd(qkv, d_o, table, num_windows, C, heads, ws, shift, Hp, Wp, q_scale):
    L = _C.lib()
    ctas = int(L.stf_attention_bwd_slots(num_windows, C, heads, ws))
    if ctas < 0:
        _C.check(ctas, "stf_attention_bwd_slots")
    dqkv = torch.empty_like(qkv)
    part = torch.empty((ctas, (2 * ws - 1) ** 2, heads), dtype=torch.float32, device=qkv.device)
    ops._launch("window_attention16_bwd_kernel", 4 * qkv.shape[0] * 7 * C, L.stf_window_attention_bwd, qkv.data_ptr(),
                d_o.data_ptr(), table.data_ptr(), dqkv.data_ptr(), part.data_ptr(), int(num_windows), C, heads, ws,
                shift, Hp, Wp, float(q_scale), _C.stream())
    return dqkv, part.sum(0)


class AttentionBranch(torch.autograd.Function):
    """x -> x + proj(attn(qkv(LN1(x))))   (stf.py:152-196 without DropPath), token-major (B*L, C)."""

    @staticmethod
    def forward(ctx, x, g1, b1, wqkv, bqkv, table, wproj, bproj, blk, geom):
        B, H, W, ws, shift = geom
        Hp, Wp = ops.ceil_to(H, ws), ops.ceil_to(W, ws)      # zero-pad to whole windows (stf.py:158-163)
        C = x.shape[1]
        attn = blk.attn
        norm1 = getattr(blk, "norm1", None)           # WinBasedAttention (WACNN) has no LayerNorm in front of qkv
        qkv = ops.linear(x, attn.packed_qkv(norm1), M=B * Hp * Wp, rows=_C.ROWS_WINDOW, epilogue=_C.EPI_QKV,
                         q_cols=C, q_scale=attn.scale, geom=geom)
        o = ops.window_attention_core(qkv, table, B * (Hp // ws) * (Wp // ws), C, attn.num_heads, ws, shift, Hp, Wp)
        x1 = ops.linear(o, attn.packed_proj(), epilogue=_C.EPI_WINDOW_RESIDUAL, residual=x, geom=geom,
                        out_rows=B * H * W, x_is_tf32=True)
        ctx.save_for_backward(x, g1, b1, wqkv, table, wproj, qkv, o)
        ctx.blk, ctx.geom = blk, geom
        ctx.has_bqkv, ctx.has_bproj = bqkv is not None, bproj is not None
        return x1

    @staticmethod
    def backward(ctx, dx1):
        x, g1, b1, wqkv, table, wproj, qkv, o = ctx.saved_tensors
        blk, geom = ctx.blk, ctx.geom
        B, H, W, ws, shift = geom
        Hp, Wp = ops.ceil_to(H, ws), ops.ceil_to(W, ws)
        attn = blk.attn
        C = x.shape[1]
        dx1 = dx1.contiguous()
        idx = window_row_index(B, H, W, ws, shift, x.device)
        # proj: y_w = o . Wp^T + b, scattered to tokens (pad rows cropped: their gradient is zero, which is what the
        # WINDOW row gather feeds for them)
        d_o = ops.linear(dx1, attn._pp.get_t(wproj), M=B * Hp * Wp, rows=_C.ROWS_WINDOW, geom=geom)
        has_pad = (Hp != H) or (Wp != W)
        dy_w = gather_rows(dx1, idx, has_pad)
        dwproj = wgrad(dy_w, o)
        dbproj = colsum(dx1)
        # attention core
        dqkv, dtable = attention_bwd(qkv, d_o, table, B * (Hp // ws) * (Wp // ws), C, attn.num_heads, ws, shift, Hp, Wp,
                                     attn.scale)
        if g1 is None:
            # no LayerNorm: the qkv input gradient lands in token order on top of the shortcut gradient directly
            dx = ops.linear(dqkv, attn._pq.get_t(wqkv), epilogue=_C.EPI_WINDOW_RESIDUAL, residual=dx1, geom=geom,
                            out_rows=B * H * W)
            dwqkv = wgrad(dqkv, gather_rows(x, idx, has_pad))
            return (dx, None, None, dwqkv, colsum(dqkv) if ctx.has_bqkv else None, dtable, dwproj,
                    dbproj if ctx.has_bproj else None, None, None)
        # qkv Linear (window-ordered rows -> token order) then LayerNorm 1, plus the shortcut gradient
        zeros = torch.zeros_like(x)
        g = ops.linear(dqkv, attn._pq.get_t(wqkv), epilogue=_C.EPI_WINDOW_RESIDUAL, residual=zeros, geom=geom,
                       out_rows=B * H * W)
        dx, xn, dg1, db1 = layernorm_bwd(x, g, g1, b1, blk.norm1.eps, res=dx1)
        dwqkv = wgrad(dqkv, gather_rows(xn, idx, has_pad))
        dbqkv = colsum(dqkv) if ctx.has_bqkv else None      # qkv_bias=False: the input was None (stf.py:66)
        return dx, dg1, db1, dwqkv, dbqkv, dtable, dwproj, dbproj if ctx.has_bproj else None, None, None


class MlpBranch(torch.autograd.Function):
    """x -> x + fc2(GELU(fc1(LN2(x))))   (stf.py:197, 34-40 without DropPath)."""

    @staticmethod
    def forward(ctx, x, g2, b2, w1, bb1, w2, bb2, blk):
        mlp = blk.mlp
        h = ops.linear(x, mlp._p1.get(w1, bb1, blk.norm2), epilogue=_C.EPI_GELU)
        y = ops.linear(h, mlp._p2.get(w2, bb2), epilogue=_C.EPI_RESIDUAL, residual=x, x_is_tf32=True)
        ctx.save_for_backward(x, g2, b2, w1, bb1, w2, h)
        ctx.blk = blk
        return y

    @staticmethod
    def backward(ctx, dy):
        x, g2, b2, w1, bb1, w2, h = ctx.saved_tensors
        blk = ctx.blk
        mlp = blk.mlp
        dy = dy.contiguous()
        dh = ops.linear(dy, mlp._p2.get_t(w2))
        dw2 = wgrad(dy, h)
        db2 = colsum(dy)
        pre = ops.linear(x, mlp._p1.get(w1, bb1, blk.norm2))          # recompute the fc1 pre-activation
        dpre = gelu_bwd(pre, dh)
        g = ops.linear(dpre, mlp._p1.get_t(w1))
        dx, xn, dg2, dbeta2 = layernorm_bwd(x, g, g2, b2, blk.norm2.eps, res=dy)
        dw1 = wgrad(dpre, xn)
        db1 = colsum(dpre)
        return dx, dg2, dbeta2, dw1, db1, dw2, db2, None


class GaussianLikelihoodTrain(torch.autograd.Function):
    """GaussianConditional.forward in training mode (entropy_models.py:645-659) -> likelihood of y + noise."""

    @staticmethod
    def forward(ctx, y, scales, means, noise, scale_bound, lik_bound):
        y, scales = y.contiguous(), scales.contiguous()
        means = None if means is None else means.contiguous()
        noise = None if noise is None else noise.contiguous()
        lik = torch.empty_like(y)
        ops._launch("gaussian_train_fwd_kernel", 20 * y.numel(), _C.lib().stf_gaussian_likelihood_train, y.data_ptr(),
                    scales.data_ptr(), _C.ptr(means), _C.ptr(noise), lik.data_ptr(), y.numel(), float(scale_bound),
                    float(lik_bound), _C.stream())
        ctx.save_for_backward(y, scales, means, noise)
        ctx.bounds = (float(scale_bound), float(lik_bound))
        return lik

    @staticmethod
    def backward(ctx, dlik):
        y, scales, means, noise = ctx.saved_tensors
        dlik = dlik.contiguous()
        dy, ds = torch.empty_like(y), torch.empty_like(y)
        dm = torch.empty_like(y) if means is not None else None
        ops._launch("gaussian_train_bwd_kernel", 32 * y.numel(), _C.lib().stf_gaussian_likelihood_train_bwd,
                    y.data_ptr(), scales.data_ptr(), _C.ptr(means), _C.ptr(noise), dlik.data_ptr(), dy.data_ptr(),
                    ds.data_ptr(), _C.ptr(dm), y.numel(), ctx.bounds[0], ctx.bounds[1], _C.stream())
        return dy, ds, dm, None, None, None


class LowerBoundFunction(torch.autograd.Function):
    """max(x, bound) with the reference's gradient rule (ops/bound_ops.py:21-27): pass where x >= bound or grad < 0.
    Tiny element-wise glue around the entropy bottleneck (18 k elements); the Gaussian path has it fused in its kernel."""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, g):
        x, bound = ctx.saved_tensors
        return ((x >= bound) | (g < 0)).type(g.dtype) * g, None


def ste_round(x):
    """round(x) with identity gradient (ops/ops.py:20-34)."""
    return torch.round(x) - x.detach() + x


def drop_path_scale(x, drop_prob, training):
    """Per-sample stochastic-depth factor (timm 0.4.12 drop_path: floor(keep + U) / keep), shape (B, 1, 1)."""
    if drop_prob == 0.0 or not training:
        return None
    keep = 1.0 - drop_prob
    r = keep + torch.rand((x.shape[0], 1, 1), dtype=x.dtype, device=x.device)
    return r.floor_() / keep
