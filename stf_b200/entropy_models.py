"""Host-side mirror of `compressai.entropy_models` running on libstf_b200 kernels.

Same classes, method names, argument meaning, buffer / parameter names and error behaviour as the
reference (compressai/entropy_models/entropy_models.py:70-666):

  EntropyModel          quantize / dequantize / compress / decompress, _offset/_quantized_cdf/_cdf_length
  EntropyBottleneck     forward (eval), compress, decompress, update, loss, _get_medians, _build_indexes
  GaussianConditional   forward (eval), build_indexes, update_scale_table, update
  LowerBound            compressai/ops/bound_ops.py:21-65 (module kept for its `bound` buffer / keys)

What runs where:
  * per-element work (quantize, dequantize, Gaussian / logistic likelihoods, build_indexes) = one
    CUDA kernel each (stf_b200/csrc/entropy_kernels.cu) -- the reference spends 4-190 launches on each;
  * entropy coding = the host rANS codec of the same library (csrc/rans_host.cpp), fed with int32
    buffers copied once from the device, one stream per image on its own host thread;
  * update() (CDF table construction, once per model) = host-side set-up in torch CPU fp32, in the
    reference's operation order so the tables are bit-identical, quantised by stf_pmf_to_quantized_cdf.
Training-mode forward (additive noise, ste_round, LowerBound gradient rule): GaussianConditional through
stf_b200/autograd.py's fused likelihood kernels, EntropyBottleneck through torch autograd over the reference's op
sequence (an 18 k-element tensor); see DESIGN.md section 4.5.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import _C, ans, ops

_GAUSS_TAIL = {1e-9: 6.1094102048693975}   # -scipy.stats.norm.ppf(tail_mass / 2), entropy_models.py:600


def _std_quantile_multiplier(tail_mass):
    if tail_mass in _GAUSS_TAIL:
        return _GAUSS_TAIL[tail_mass]
    import scipy.stats
    return float(-scipy.stats.norm.ppf(tail_mass / 2))


class LowerBound(nn.Module):
    """max(x, bound); holds the 1-element `bound` buffer (checkpoint key `*.bound`)."""

    def __init__(self, bound):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x):
        if x.requires_grad and torch.is_grad_enabled():
            from .autograd import LowerBoundFunction
            return LowerBoundFunction.apply(x, self.bound)
        return torch.max(x, self.bound)


class EntropyModel(nn.Module):
    """Entropy model base class (entropy_models.py:70-290)."""

    def __init__(self, likelihood_bound=1e-9, entropy_coder=None, entropy_coder_precision=16):
        super().__init__()
        if entropy_coder not in (None, "ans"):
            raise ValueError(f'Unknown entropy coder "{entropy_coder}" (available: ans)')
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.use_likelihood_bound = likelihood_bound > 0
        self._likelihood_bound = float(likelihood_bound)
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())
        self._rans_table = None   # (key, ans.RansTable)

    # -- table access ------------------------------------------------------------------------
    @property
    def offset(self):
        return self._offset

    @property
    def quantized_cdf(self):
        return self._quantized_cdf

    @property
    def cdf_length(self):
        return self._cdf_length

    def rans_table(self):
        """Prepared coder tables (encoder reciprocals + decoder LUT), rebuilt when the buffers change."""
        self._check_cdf_size()
        self._check_cdf_length()
        self._check_offsets_size()
        key = (self._quantized_cdf.data_ptr(), self._quantized_cdf._version, tuple(self._quantized_cdf.shape))
        if self._rans_table is None or self._rans_table[0] != key:
            self._rans_table = (key, ans.RansTable(self._quantized_cdf.cpu().numpy(),
                                                   self._cdf_length.reshape(-1).int().cpu().numpy(),
                                                   self._offset.reshape(-1).int().cpu().numpy()))
        return self._rans_table[1]

    # -- quantisation ------------------------------------------------------------------------
    def quantize(self, inputs, mode, means=None):
        if mode not in ("noise", "dequantize", "symbols"):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        if mode == "noise":
            return inputs + torch.empty_like(inputs).uniform_(-0.5, 0.5)
        x = inputs.contiguous()
        if mode == "dequantize":
            return ops.quantize_dequantize(x, means)
        return ops.quantize_symbols(x, means)

    @staticmethod
    def dequantize(inputs, means=None):
        if means is None:
            return inputs.float()
        if not inputs.dtype.is_floating_point and means.is_cuda and means.shape == inputs.shape:
            sym = inputs.to(device=means.device, dtype=torch.int32).contiguous()
            m = means.contiguous()
            return ops.dequantize(sym.reshape(1, -1), 0, m.reshape(1, 1, -1)).reshape(means.shape)
        outputs = inputs.type_as(means)
        outputs = outputs + means
        return outputs

    def _pmf_to_cdf(self, pmf, tail_mass, pmf_length, max_length):
        cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
        for i, p in enumerate(pmf):
            prob = torch.cat((p[: int(pmf_length[i])], tail_mass[i]), dim=0)
            row = ans.pmf_to_quantized_cdf(prob.tolist(), self.entropy_coder_precision)
            cdf[i, : len(row)] = torch.tensor(row, dtype=torch.int32)
        return cdf

    def _check_cdf_size(self):
        if self._quantized_cdf.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")
        if self._quantized_cdf.dim() != 2:
            raise ValueError(f"Invalid CDF size {self._quantized_cdf.size()}")

    def _check_offsets_size(self):
        if self._offset.numel() == 0:
            raise ValueError("Uninitialized offsets. Run update() first")
        if self._offset.dim() != 1:
            raise ValueError(f"Invalid offsets size {self._offset.size()}")

    def _check_cdf_length(self):
        if self._cdf_length.numel() == 0:
            raise ValueError("Uninitialized CDF lengths. Run update() first")
        if self._cdf_length.dim() != 1:
            raise ValueError(f"Invalid offsets size {self._cdf_length.size()}")

    # -- entropy coding ----------------------------------------------------------------------
    def compress(self, inputs, indexes, means=None, flag=1):
        """inputs (B, ...) fp32, indexes same shape int32 -> list of B byte strings."""
        if inputs.dim() < 2:
            raise ValueError("Invalid `inputs` size. Expected a tensor with at least 2 dimensions.")
        if inputs.size() != indexes.size():
            raise ValueError("`inputs` and `indexes` should have the same size.")
        table = self.rans_table()
        symbols = self.quantize(inputs, "symbols", means)
        B = symbols.shape[0]
        sym = symbols.reshape(B, -1).cpu().numpy()
        idx = indexes.reshape(B, -1).int().cpu().numpy()
        return ans.encode_batch(table, list(sym), list(idx))

    def decompress(self, strings, indexes, means=None, flag=1):
        if not isinstance(strings, (tuple, list)):
            raise ValueError("Invalid `strings` parameter type.")
        if not len(strings) == indexes.size(0):
            raise ValueError("Invalid strings or indexes parameters")
        if indexes.dim() < 2:
            raise ValueError("Invalid `indexes` size. Expected a tensor with at least 2 dimensions.")
        table = self.rans_table()
        if means is not None:
            if means.size()[:2] != indexes.size()[:2]:
                raise ValueError("Invalid means or indexes parameters")
            if means.size() != indexes.size():
                for i in range(2, indexes.dim()):
                    if means.size(i) != 1:
                        raise ValueError("Invalid means parameters")
        B = len(strings)
        idx = indexes.reshape(B, -1).int().cpu().numpy()
        decs = []
        for s in strings:
            d = ans.RansDecoder()
            d.set_stream(s)
            decs.append(d)
        outs = ans.decode_batch(decs, table, list(idx))
        device = self._quantized_cdf.device
        sym = torch.from_numpy(np.stack(outs)).reshape(indexes.shape).to(device)
        if means is not None and means.is_cuda:
            return self.dequantize(sym, means.expand_as(sym).contiguous())
        return self.dequantize(sym, means)


class EntropyBottleneck(EntropyModel):
    """Factorised-prior entropy bottleneck (entropy_models.py:293-522)."""

    def __init__(self, channels, *args, tail_mass=1e-9, init_scale=10, filters=(3, 3, 3, 3), **kwargs):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        if self.filters != (3, 3, 3, 3):
            raise ValueError("stf_b200: the logistic-CDF kernel is specialised for filters=(3,3,3,3) "
                             "(the only configuration the reference models use)")
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        f = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        for i in range(len(self.filters) + 1):
            init = math.log(math.expm1(1 / scale / f[i + 1]))
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(torch.full((channels, f[i + 1], f[i]), init)))
            self.register_parameter(f"_bias{i:d}", nn.Parameter(torch.empty(channels, f[i + 1], 1).uniform_(-0.5, 0.5)))
            if i < len(self.filters):
                self.register_parameter(f"_factor{i:d}", nn.Parameter(torch.zeros(channels, f[i + 1], 1)))
        self.quantiles = nn.Parameter(torch.tensor([-self.init_scale, 0.0, self.init_scale]).repeat(channels, 1, 1))
        target = math.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))
        self._packed = None

    def _get_medians(self):
        return self.quantiles[:, :, 1:2]

    def _param_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def packed_params(self):
        """(C, 60) device table: softplus(matrix) / bias / tanh(factor) per layer + median (header layout)."""
        key = self._param_key()
        if self._packed is None or self._packed[0] != key:
            with torch.no_grad():
                C = self.channels
                cols = []
                for i in range(5):
                    cols.append(torch.nn.functional.softplus(getattr(self, f"_matrix{i}")).reshape(C, -1))
                    cols.append(getattr(self, f"_bias{i}").reshape(C, -1))
                    if i < 4:
                        cols.append(torch.tanh(getattr(self, f"_factor{i}")).reshape(C, -1))
                cols.append(self.quantiles[:, 0, 1:2])
                cols.append(torch.zeros(C, 1, device=self.quantiles.device))
                packed = torch.cat(cols, dim=1).float().contiguous()
                assert packed.shape[1] == _C.STF_EB_PARAMS
            self._packed = (key, packed)
        return self._packed[1]

    def _logits_cumulative(self, inputs, stop_gradient=True, params=None):
        """entropy_models.py:400-419 in plain torch: only used by update() / loss() (set-up, not the hot
        path).  `params` optionally substitutes host copies of the parameters."""
        get = (lambda n: params[n]) if params is not None else (lambda n: getattr(self, n))
        logits = inputs
        for i in range(len(self.filters) + 1):
            matrix, bias = get(f"_matrix{i:d}"), get(f"_bias{i:d}")
            if stop_gradient:
                matrix, bias = matrix.detach(), bias.detach()
            logits = torch.matmul(torch.nn.functional.softplus(matrix), logits) + bias
            if i < len(self.filters):
                factor = get(f"_factor{i:d}")
                if stop_gradient:
                    factor = factor.detach()
                logits = logits + torch.tanh(factor) * torch.tanh(logits)
        return logits

    def update(self, force=False):
        if self._offset.numel() > 0 and not force:
            return False
        device = self.quantiles.device
        with torch.no_grad():
            host = {n: p.detach().float().cpu() for n, p in self.named_parameters(recurse=False)}
            q = host["quantiles"]
            medians = q[:, 0, 1]
            minima = torch.clamp(torch.ceil(medians - q[:, 0, 0]).int(), min=0)
            maxima = torch.clamp(torch.ceil(q[:, 0, 2] - medians).int(), min=0)
            pmf_start = medians - minima
            pmf_length = maxima + minima + 1
            max_length = int(pmf_length.max().item())
            samples = torch.arange(max_length)[None, :] + pmf_start[:, None, None]
            lower = self._logits_cumulative(samples - 0.5, params=host)
            upper = self._logits_cumulative(samples + 0.5, params=host)
            sign = -torch.sign(lower + upper)
            pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
            tail_mass = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
            self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length).to(device)
            self._offset = (-minima).to(device)
            self._cdf_length = (pmf_length + 2).to(device)
        return True

    def loss(self):
        logits = self._logits_cumulative(self.quantiles, stop_gradient=True)
        return torch.abs(logits - self.target).sum()

    def forward(self, x, training=None, noise=None):
        if training is None:
            training = self.training
        if x.dim() < 2 or x.shape[1] != self.channels:
            raise ValueError(f"expected (B, {self.channels}, ...) input, got {tuple(x.shape)}")
        if training:      # noise quantisation is selected by `training` alone (entropy_models.py:131-135, 456)
            return self._forward_train(x, noise)
        z_hat, lik, _ = ops.entropy_bottleneck(x.contiguous(), self.packed_params(),
                                               lik_bound=self._likelihood_bound if self.use_likelihood_bound else 0.0)
        return z_hat, lik

    def _forward_train(self, x, noise=None):
        """Training forward (entropy_models.py:446-489 with "noise" quantisation), differentiable w.r.t. x and the 58
        parameters per channel.  z is 18 k - 50 k elements: latency-bound either way, so this path is the reference's own
        op sequence on torch's autograd (the eval kernel has no backward); the LowerBound keeps its custom gradient."""
        perm = [1, 0] + list(range(2, x.dim()))
        v = x.permute(*perm).contiguous()
        shape = v.size()
        v = v.reshape(v.size(0), 1, -1)
        n = torch.empty_like(v).uniform_(-0.5, 0.5) if noise is None else \
            noise.permute(*perm).contiguous().reshape(v.shape)
        out = v + n
        lower = self._logits_cumulative(out - 0.5, stop_gradient=False)
        upper = self._logits_cumulative(out + 0.5, stop_gradient=False)
        sign = -torch.sign(lower + upper).detach()
        lik = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))
        if self.use_likelihood_bound:
            lik = self.likelihood_lower_bound(lik)
        inv = [1, 0] + list(range(2, x.dim()))
        return out.reshape(shape).permute(*inv).contiguous(), lik.reshape(shape).permute(*inv).contiguous()

    @staticmethod
    def _build_indexes(size):
        N, C = size[0], size[1]
        view = [1] * len(size)
        view[1] = C
        return torch.arange(C, dtype=torch.int32).view(*view).repeat(N, 1, *size[2:])

    @staticmethod
    def _extend_ndims(tensor, n):
        return tensor.reshape(-1, *([1] * n)) if n > 0 else tensor.reshape(-1)

    def compress(self, x):
        """One string per batch element (entropy_models.py:508-515)."""
        table = self.rans_table()
        if x.dim() < 2:
            raise ValueError("Invalid `inputs` size. Expected a tensor with at least 2 dimensions.")
        _, _, sym = ops.entropy_bottleneck(x.contiguous(), self.packed_params(), want_z_hat=False, want_lik=False,
                                           want_symbols=True)
        return self.encode_symbols(sym)

    def encode_symbols(self, sym):
        """int32 symbols (B, C, ...) on the device -> one rANS string per image."""
        table = self.rans_table()
        B = sym.shape[0]
        idx = self._build_indexes(sym.size()).reshape(B, -1).numpy()
        return ans.encode_batch(table, list(sym.reshape(B, -1).cpu().numpy()), list(idx))

    def decompress(self, strings, size):
        output_size = (len(strings), self._quantized_cdf.size(0), *size)
        indexes = self._build_indexes(output_size)
        medians = self._extend_ndims(self._get_medians().detach(), len(size))
        medians = medians.expand(len(strings), *([-1] * (len(size) + 1)))
        return super().decompress(strings, indexes, medians, 0)


class GaussianConditional(EntropyModel):
    """Gaussian conditional layer (entropy_models.py:525-666)."""

    def __init__(self, scale_table, *args, scale_bound=0.11, tail_mass=1e-9, **kwargs):
        super().__init__(*args, **kwargs)
        if not isinstance(scale_table, (type(None), list, tuple)):
            raise ValueError(f'Invalid type for scale_table "{type(scale_table)}"')
        if isinstance(scale_table, (list, tuple)) and len(scale_table) < 1:
            raise ValueError(f'Invalid scale_table length "{len(scale_table)}"')
        if scale_table and (scale_table != sorted(scale_table) or any(s <= 0 for s in scale_table)):
            raise ValueError(f'Invalid scale_table "({scale_table})"')
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            scale_bound = scale_table[0]
        if scale_bound <= 0:
            raise ValueError("Invalid parameters")
        self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer("scale_table", self._prepare_scale_table(scale_table) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]))
        self._table_host = None

    @staticmethod
    def _prepare_scale_table(scale_table):
        return torch.Tensor(tuple(float(s) for s in scale_table))

    def _standardized_cumulative(self, inputs):
        return 0.5 * torch.erfc(float(-(2 ** -0.5)) * inputs)

    def update_scale_table(self, scale_table, force=False):
        if self._offset.numel() > 0 and not force:
            return False
        device = self.scale_table.device
        self.scale_table = self._prepare_scale_table(scale_table).to(device)
        self.update()
        return True

    def update(self):
        """Host-side table construction (entropy_models.py:599-624), torch CPU fp32 in the reference's order."""
        device = self.scale_table.device
        table = self.scale_table.detach().float().cpu()
        multiplier = _std_quantile_multiplier(self.tail_mass)
        pmf_center = torch.ceil(table * multiplier).int()
        pmf_length = 2 * pmf_center + 1
        max_length = int(torch.max(pmf_length).item())
        samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
        scale = table.unsqueeze(1).float()
        upper = self._standardized_cumulative((0.5 - samples) / scale)
        lower = self._standardized_cumulative((-0.5 - samples) / scale)
        pmf = upper - lower
        tail_mass = 2 * lower[:, :1]
        self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length).to(device)
        self._offset = (-pmf_center).to(device)
        self._cdf_length = (pmf_length + 2).to(device)

    def host_scale_table(self):
        key = (self.scale_table.data_ptr(), self.scale_table._version, self.scale_table.numel())
        if self._table_host is None or self._table_host[0] != key:
            if self.scale_table.numel() == 0:
                raise ValueError("Uninitialized scale table. Run update_scale_table() first")
            self._table_host = (key, self.scale_table.detach().float().cpu().numpy().copy())
        return self._table_host[1]

    def forward(self, inputs, scales, means=None, training=None, noise=None):
        if training is None:
            training = self.training
        if training:
            from .autograd import GaussianLikelihoodTrain
            n = torch.empty_like(inputs).uniform_(-0.5, 0.5) if noise is None else noise
            lik = GaussianLikelihoodTrain.apply(inputs, scales.expand_as(inputs), None if means is None else
                                                means.expand_as(inputs), n, self.scale_bound_value(),
                                                self._likelihood_bound if self.use_likelihood_bound else 0.0)
            return inputs + n, lik
        x = inputs.contiguous()
        x4 = x.reshape(x.shape[0], x.shape[1] if x.dim() > 1 else 1, -1)
        s = scales.expand_as(inputs).contiguous().reshape(x4.shape)
        m = None if means is None else means.expand_as(inputs).contiguous().reshape(x4.shape)
        out, lik = ops.gaussian_likelihood(x4, 0, s, m, scale_bound=self.scale_bound_value(),
                                           lik_bound=self._likelihood_bound if self.use_likelihood_bound else 0.0)
        return out.reshape(inputs.shape), lik.reshape(inputs.shape)

    def scale_bound_value(self):
        """fp32 value of the lower bound (0.11 -> 0.10999999940395355), cached off the device once."""
        if getattr(self, "_sb_cache", None) is None:
            self._sb_cache = float(self.lower_bound_scale.bound.detach().cpu().item())
        return self._sb_cache

    def build_indexes(self, scales):
        return ops.build_indexes(scales.contiguous(), self.host_scale_table(), self.scale_bound_value())
