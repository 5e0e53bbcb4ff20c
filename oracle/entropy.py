"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's entropy-model arithmetic.

Every function states the reference lines it follows (paths relative to /root/reference).
Floating-point steps use torch CPU fp32 ops in the reference's order; integer steps use numpy /
oracle/rans_oracle.c.  State is passed as plain tensors / dicts (no nn.Module), so the file is a
restatement of the algorithm, not of the reference's class layout.
"""
import ctypes
import math
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SCALE_MIN, SCALE_MAX, SCALE_LEVELS = 0.11, 256, 64          # stf.py:16-18
LIKELIHOOD_FLOOR = 1e-9                                      # entropy_models.py:82
GAUSS_TAIL_MULT = 6.1094102048693975   # -scipy.stats.norm.ppf(1e-9 / 2), entropy_models.py:600


def lib():
    """oracle/_build/liboracle.so (built by `make -C oracle` / __graft_entry__.build())."""
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", _HERE, "_build/liboracle.so"])
        L = ctypes.CDLL(path)
        i32p, u8p, u32p, f32p = (ctypes.POINTER(t) for t in (ctypes.c_int32, ctypes.c_uint8, ctypes.c_uint32, ctypes.c_float))
        L.oracle_rans_encode.restype = ctypes.c_long
        L.oracle_rans_encode.argtypes = [i32p, i32p, ctypes.c_long, i32p, ctypes.c_int, i32p, i32p, u8p, ctypes.c_long]
        L.oracle_rans_dec_init.restype = None
        L.oracle_rans_dec_init.argtypes = [ctypes.c_void_p, u8p]
        L.oracle_rans_dec_run.restype = None
        L.oracle_rans_dec_run.argtypes = [ctypes.c_void_p, i32p, ctypes.c_long, i32p, ctypes.c_int, i32p, i32p, i32p]
        L.oracle_pmf_to_quantized_cdf.restype = ctypes.c_int
        L.oracle_pmf_to_quantized_cdf.argtypes = [f32p, ctypes.c_int, ctypes.c_int, u32p]
        _LIB = L
    return _LIB


# ----------------------------------------------------------------------------- tables

def scale_table(lo=SCALE_MIN, hi=SCALE_MAX, levels=SCALE_LEVELS):
    """stf.py:21-22 / cnn.py:19-20: exp(linspace(ln lo, ln hi, levels)), fp32."""
    return torch.exp(torch.linspace(math.log(lo), math.log(hi), levels))


def pmf_to_quantized_cdf(pmf, precision=16):
    """entropy_models.py:60-63 -> ops.cpp:24-81 (restated in rans_oracle.c). pmf: 1-D fp32."""
    p = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32))
    out = np.zeros(p.size + 1, dtype=np.uint32)
    rc = lib().oracle_pmf_to_quantized_cdf(
        p.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), p.size, precision,
        out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)))
    if rc != 0:
        raise RuntimeError("pmf_to_quantized_cdf: no donor symbol")
    return out.astype(np.int32)


def _rows_to_cdf(pmf, tail_mass, lengths, max_length, precision=16):
    """entropy_models.py:172-180 (_pmf_to_cdf)."""
    cdf = np.zeros((len(lengths), max_length + 2), dtype=np.int32)
    for i in range(len(lengths)):
        row = torch.cat((pmf[i, : int(lengths[i])], tail_mass[i]), dim=0).numpy()
        q = pmf_to_quantized_cdf(row, precision)
        cdf[i, : q.size] = q
    return cdf


def std_normal_cdf(v):
    """entropy_models.py:578-582: 0.5 * erfc(-(2 ** -0.5) * v)."""
    return 0.5 * torch.erfc(float(-(2 ** -0.5)) * v)


def gaussian_tables(table=None, tail_mass=1e-9):
    """GaussianConditional.update, entropy_models.py:599-624.
    Returns (quantized_cdf int32 (L, maxlen+2), cdf_length int32 (L,), offset int32 (L,))."""
    table = scale_table() if table is None else torch.as_tensor(table, dtype=torch.float32)
    center = torch.ceil(table * GAUSS_TAIL_MULT).int()
    length = 2 * center + 1
    max_length = int(length.max())
    k = torch.abs(torch.arange(max_length).int() - center[:, None]).float()
    s = table.unsqueeze(1).float()
    upper = std_normal_cdf((0.5 - k) / s)
    lower = std_normal_cdf((-0.5 - k) / s)
    pmf = upper - lower
    tail = 2 * lower[:, :1]
    cdf = _rows_to_cdf(pmf, tail, length, max_length)
    return cdf, (length + 2).numpy().astype(np.int32), (-center).numpy().astype(np.int32)


# ----------------------------------------------------------------------------- quantize

def quantize(x, mode, means=None):
    """EntropyModel.quantize (eval modes only), entropy_models.py:126-150.
    torch.round == round-half-to-even."""
    if mode not in ("dequantize", "symbols"):
        raise ValueError(f'Invalid quantization mode: "{mode}"')
    v = x.clone()
    if means is not None:
        v = v - means
    v = torch.round(v)
    if mode == "dequantize":
        return v + means if means is not None else v
    return v.int()


def dequantize(sym, means=None):
    """EntropyModel.dequantize, entropy_models.py:158-165."""
    if means is None:
        return sym.float()
    return sym.type_as(means) + means


def ste_round_value(x):
    """ops/ops.py:34 -- forward value of ste_round: round(x) - x + x evaluated left to right."""
    return torch.round(x) - x + x


# ----------------------------------------------------------------------------- GaussianConditional

def lower_bound(x, bound):
    """ops/bound_ops.py:21-22 -- torch.max(x, bound) with bound a 1-element fp32 tensor."""
    return torch.max(x, torch.tensor([float(bound)], dtype=x.dtype))


def gaussian_likelihood(values, scales, means=None, scale_bound=0.11):
    """GaussianConditional._likelihood, entropy_models.py:626-643 (no likelihood floor)."""
    v = values - means if means is not None else values
    s = lower_bound(scales, scale_bound)
    v = torch.abs(v)
    upper = std_normal_cdf((0.5 - v) / s)
    lower = std_normal_cdf((-0.5 - v) / s)
    return upper - lower


class LowerBoundFn(torch.autograd.Function):
    """ops/bound_ops.py:21-65 -- forward max(x, bound); backward passes the gradient where x >= bound or grad < 0."""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, grad_output):
        x, bound = ctx.saved_tensors
        pass_through = (x >= bound) | (grad_output < 0)
        return pass_through.type(grad_output.dtype) * grad_output, None


def lower_bound_train(x, bound):
    return LowerBoundFn.apply(x, torch.tensor([float(bound)], dtype=x.dtype))


def ste_round(x):
    """ops/ops.py:20-34 -- round(x) - x.detach() + x (identity gradient)."""
    return torch.round(x) - x.detach() + x


def gaussian_conditional_train(x, scales, means, noise, scale_bound=0.11):
    """GaussianConditional.forward with training=True, entropy_models.py:645-659 + :131-135 ("noise" ignores the
    means) + :626-643, with the injected noise tensor standing for U(-1/2, 1/2).  Differentiable."""
    out = x + noise
    v = torch.abs(out - means if means is not None else out)
    s = lower_bound_train(scales, scale_bound)
    lik = std_normal_cdf((0.5 - v) / s) - std_normal_cdf((-0.5 - v) / s)
    return out, lower_bound_train(lik, LIKELIHOOD_FLOOR)


def eb_forward_train(p, z, noise):
    """EntropyBottleneck.forward with training=True, entropy_models.py:446-489 (noise quantisation, :131-135;
    the sign of :428-429 is detached).  Differentiable w.r.t. z and the parameters."""
    perm = list(range(z.ndim))
    perm[0], perm[1] = 1, 0
    zc = z.permute(*perm).contiguous()
    shape = zc.shape
    out = zc.reshape(shape[0], 1, -1) + noise.permute(*perm).contiguous().reshape(shape[0], 1, -1)
    lo = eb_logits_cumulative(p, out - 0.5)
    hi = eb_logits_cumulative(p, out + 0.5)
    sgn = -torch.sign(lo + hi).detach()
    lik = lower_bound_train(torch.abs(torch.sigmoid(sgn * hi) - torch.sigmoid(sgn * lo)), LIKELIHOOD_FLOOR)
    return out.reshape(shape).permute(*perm).contiguous(), lik.reshape(shape).permute(*perm).contiguous()


def gaussian_conditional_eval(x, scales, means=None):
    """GaussianConditional.forward with training=False, entropy_models.py:645-659.
    Returns (outputs, likelihood)."""
    out = quantize(x, "dequantize", means)
    lik = gaussian_likelihood(out, scales, means)
    return out, lower_bound(lik, LIKELIHOOD_FLOOR)


def build_indexes(scales, table=None, scale_bound=0.11):
    """GaussianConditional.build_indexes, entropy_models.py:661-666."""
    table = scale_table() if table is None else torch.as_tensor(table, dtype=torch.float32)
    s = lower_bound(scales, scale_bound)
    idx = torch.full(s.shape, len(table) - 1, dtype=torch.int32)
    for t in table[:-1]:
        idx -= (s <= t).int()
    return idx


# ----------------------------------------------------------------------------- EntropyBottleneck

EB_FILTERS = (1, 3, 3, 3, 3, 1)  # entropy_models.py:313,324


def eb_param_names():
    names = []
    for i in range(5):
        names += [f"_matrix{i}", f"_bias{i}"]
        if i < 4:
            names.append(f"_factor{i}")
    return names + ["quantiles"]


def eb_logits_cumulative(p, x):
    """EntropyBottleneck._logits_cumulative, entropy_models.py:400-419.  x: (C,1,L)."""
    h = x
    for i in range(5):
        h = torch.matmul(torch.nn.functional.softplus(p[f"_matrix{i}"]), h)
        h = h + p[f"_bias{i}"]
        if i < 4:
            h = h + torch.tanh(p[f"_factor{i}"]) * torch.tanh(h)
    return h


def eb_likelihood(p, x):
    """EntropyBottleneck._likelihood, entropy_models.py:421-433."""
    lo = eb_logits_cumulative(p, x - 0.5)
    hi = eb_logits_cumulative(p, x + 0.5)
    sgn = -torch.sign(lo + hi)
    return torch.abs(torch.sigmoid(sgn * hi) - torch.sigmoid(sgn * lo))


def eb_medians(p):
    """entropy_models.py:350-352."""
    return p["quantiles"][:, :, 1:2]


def eb_forward_eval(p, z):
    """EntropyBottleneck.forward with training=False, entropy_models.py:446-489. z: (B,C,...)."""
    perm = list(range(z.ndim))
    perm[0], perm[1] = 1, 0
    zc = z.permute(*perm).contiguous()
    shape = zc.shape
    vals = zc.reshape(shape[0], 1, -1)
    out = quantize(vals, "dequantize", eb_medians(p))
    lik = lower_bound(eb_likelihood(p, out), LIKELIHOOD_FLOOR)
    out = out.reshape(shape).permute(*perm).contiguous()
    lik = lik.reshape(shape).permute(*perm).contiguous()
    return out, lik


def eb_tables(p):
    """EntropyBottleneck.update, entropy_models.py:354-393.
    Returns (quantized_cdf, cdf_length, offset) as int32 numpy arrays."""
    q = p["quantiles"]
    med = q[:, 0, 1]
    minima = torch.clamp(torch.ceil(med - q[:, 0, 0]).int(), min=0)
    maxima = torch.clamp(torch.ceil(q[:, 0, 2] - med).int(), min=0)
    start = med - minima
    length = maxima + minima + 1
    max_length = int(length.max())
    samples = torch.arange(max_length)[None, :] + start[:, None, None]
    lo = eb_logits_cumulative(p, samples - 0.5)
    hi = eb_logits_cumulative(p, samples + 0.5)
    sgn = -torch.sign(lo + hi)
    pmf = torch.abs(torch.sigmoid(sgn * hi) - torch.sigmoid(sgn * lo))[:, 0, :]
    tail = torch.sigmoid(lo[:, 0, :1]) + torch.sigmoid(-hi[:, 0, -1:])
    cdf = _rows_to_cdf(pmf, tail, length, max_length)
    return cdf, (length + 2).numpy().astype(np.int32), (-minima).numpy().astype(np.int32)


def eb_indexes(shape):
    """EntropyBottleneck._build_indexes, entropy_models.py:491-502: channel id broadcast."""
    n, c = shape[0], shape[1]
    view = [1] * len(shape)
    view[1] = c
    return torch.arange(c, dtype=torch.int32).view(*view).expand(*shape).contiguous()


# ----------------------------------------------------------------------------- rANS (C oracle)

def _i32(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.int32))
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def rans_encode(symbols, indexes, cdf, cdf_lengths, offsets) -> bytes:
    """RansEncoder.encode_with_indexes, rans_interface.cpp:193-204 (via rans_oracle.c)."""
    s, sp = _i32(symbols)
    ix, ixp = _i32(indexes)
    c, cp = _i32(cdf)
    ln, lnp = _i32(cdf_lengths)
    of, ofp = _i32(offsets)
    cap = 8 * s.size + 64
    out = np.zeros(cap, dtype=np.uint8)
    n = lib().oracle_rans_encode(sp, ixp, s.size, cp, c.shape[1], lnp, ofp,
                                 out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), cap)
    if n < 0:
        raise RuntimeError("oracle_rans_encode failed")
    return out[:n].tobytes()


class RansStreamDecoder:
    """RansDecoder.set_stream + decode_stream, rans_interface.cpp:277-350 (via rans_oracle.c)."""

    def __init__(self, stream: bytes):
        self._buf = np.frombuffer(bytes(stream) + b"\0" * 8, dtype=np.uint8).copy()
        self._state = ctypes.create_string_buffer(16)
        lib().oracle_rans_dec_init(self._state, self._buf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))

    def decode(self, indexes, cdf, cdf_lengths, offsets):
        ix, ixp = _i32(indexes)
        c, cp = _i32(cdf)
        ln, lnp = _i32(cdf_lengths)
        of, ofp = _i32(offsets)
        out = np.zeros(ix.size, dtype=np.int32)
        lib().oracle_rans_dec_run(self._state, ixp, ix.size, cp, c.shape[1], lnp, ofp,
                                  out.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
        return out


def rans_decode(stream, indexes, cdf, cdf_lengths, offsets):
    """RansDecoder.decode_with_indexes, rans_interface.cpp:206-275."""
    return RansStreamDecoder(stream).decode(indexes, cdf, cdf_lengths, offsets)
