"""Host-side mirror of the reference's two codecs on libstf_b200 kernels.

  SymmetricalTransFormer   compressai/models/stf.py:384-788   (zoo name "stf")
  WACNN                    compressai/models/cnn.py:23-332     (zoo name "cnn")
  CompressionModel         compressai/models/base.py:13-70

Constructor arguments, module tree / checkpoint keys, and forward / compress / decompress / update /
load_state_dict / aux_loss signatures are the reference's.  What runs underneath in eval mode:
  * every Swin block, PatchMerging / PatchSplit and WinBasedAttention -> the fused tcgen05 linear kernel +
    the window-attention core (stf_b200/layers.py);
  * EntropyBottleneck / GaussianConditional quantize, likelihood, build_indexes, dequantize -> one
    kernel per slice step (stf_b200/entropy_models.py, ops.gaussian_compress_step);
  * symbols and indexes of ALL slices are written by the kernels straight into one (B, M*h*w) int32
    device buffer in the reference's coding order and cross PCIe once, instead of 24 `.tolist()` syncs;
  * rANS runs on the host codec of the same library, one stream per image per host thread.
  * the convolution stacks between the transforms (hyperprior h_a / h_mean_s / h_scale_s, cc_mean / cc_scale / lrp,
    STF's PatchEmbed and end_conv[0]) -> stf_conv2d / stf_patch_embed: implicit-GEMM tcgen05 kernels on NHWC data moved
    by tensor-map TMA, with the channel concat, bias + GELU, PixelShuffle and the LRP tail fused.
Still cuDNN through torch: WACNN's g_a / g_s convolutions + GDN (8f rank 3) and the 3-channel end_conv[2] of STF
(x_hat only).  For STF nothing upstream of a bitstream runs on cuDNN, so strings do not depend on the batch geometry.

Batch semantics: the reference's compress() concatenates a whole batch into one y-string that its own
decompress() cannot decode.  Here strings[0] holds ONE y-string PER IMAGE, identical to what a
batch-1 call on that image gives (asserted in tests/test_gpu_codec.py); for batch 1 the return value has exactly the
reference's structure.
"""
import math
import os
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _C, ans, graphs, ops
from .entropy_models import EntropyBottleneck, GaussianConditional, LowerBound
from .layers import (BasicLayer, PatchEmbed, PatchMerging, PatchSplit, Win_noShift_Attention, conv3x3,
                     subpel_conv3x3)

SCALES_MIN, SCALES_MAX, SCALES_LEVELS = 0.11, 256, 64   # stf.py:16-18

_GRAPHS_DEFAULT = os.environ.get("STF_B200_CUDA_GRAPHS", "1") != "0"
_EB_MEDIAN = _C.STF_EB_MEDIAN_SLOT
# Batches of at least this many images are coded as two pipelined halves (host rANS of one half overlaps the
# device work of the other; measured +4 % at batch 16, +8 % at batch 32).  Below it the per-slice device segments
# are latency bound (~80 dependent small kernels), so halving the batch does not halve their time.
_PIPELINE_MIN_BATCH = int(os.environ.get("STF_B200_PIPELINE_MIN_BATCH", "16"))
_PIPELINE_PARTS = int(os.environ.get("STF_B200_PIPELINE_PARTS", "0"))   # 0 = auto: 3 sub-batches from 48 images, else 2
_DEC_PARTS = int(os.environ.get("STF_B200_DEC_PARTS", "0")) or None     # decompress(): sub-batches (default: same as compress)
# decompress(): rANS decoding of the y-strings on the device (one warp lane per image, no host round trip per slice) instead of
# the host thread pool: "1" / "0", default "1"
_DEVICE_DECODE = os.environ.get("STF_B200_DEVICE_DECODE", "0") != "0"
# symbols / indexes cross PCIe as int16 / uint8 (3 instead of 8 bytes per symbol; D2H measured at 8.8 GiB/s per rank with
# eight ranks copying at once, against 26 GiB/s H2D) -- "0" keeps int32 / int32
_NARROW = os.environ.get("STF_B200_NARROW_TRANSFER", "1") != "0"
def _split_env(name):
    v = os.environ.get(name, "")
    return [float(t) for t in v.split(",")] if v else None


# explicit relative sub-batch sizes, first to last, e.g. "3,5,4" (overrides the tapered default below)
_ENC_SPLIT, _DEC_SPLIT = _split_env("STF_B200_ENC_SPLIT"), _split_env("STF_B200_DEC_SPLIT")
# CTA cap of each of the two conv stacks of a slice (parallel graph branches): half the SMs each, so that the latency-bound
# small layers of the two stacks overlap instead of queueing behind each other (batch 64: -1.3 ms per step; 0 = no cap)
_SLICE_CTAS = int(os.environ.get("STF_B200_SLICE_CTAS", "74"))
_PART_TAPER = float(os.environ.get("STF_B200_PART_TAPER", "0.75"))    # size of the last sub-batch relative to the first
_DEC_LEAD = int(os.environ.get("STF_B200_DEC_LEAD", "3"))               # decompress(): slices a sub-batch may lead the next one by
TRACE = None         # set to a list to collect a host / device timeline of the pipelined step (tools/timeline.py)
PHASE_TIMES = None   # set to a dict to collect a synchronised wall-clock breakdown (tools/phase_breakdown.py)


class _trace:
    """`with _trace("name", part, slice, device=bool):` -- no-op unless TRACE is a list.  Host spans are perf_counter pairs;
    device spans are CUDA event pairs recorded on the current stream (resolved by tools/timeline.py after a sync)."""

    def __init__(self, name, part=0, step=0, device=False):
        self.key, self.device = (name, part, step), device

    def __enter__(self):
        if TRACE is not None:
            if self.device:
                self.e0 = torch.cuda.Event(enable_timing=True)
                self.e0.record()
            self.t0 = time.perf_counter()
        return self

    def __exit__(self, *exc):
        if TRACE is not None:
            t1 = time.perf_counter()
            TRACE.append(("host",) + self.key + (self.t0, t1))
            if self.device:
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record()
                TRACE.append(("dev",) + self.key + (self.e0, e1))
        return False


class _phase:
    """`with _phase("name"):` -- no-op unless PHASE_TIMES is a dict (then: device sync on both sides)."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if PHASE_TIMES is not None:
            import time
            torch.cuda.synchronize()
            self.t0 = time.perf_counter()

    def __exit__(self, *exc):
        if PHASE_TIMES is not None:
            import time
            torch.cuda.synchronize()
            PHASE_TIMES[self.name] = PHASE_TIMES.get(self.name, 0.0) + (time.perf_counter() - self.t0) * 1e3
        return False


def get_scale_table(min=SCALES_MIN, max=SCALES_MAX, levels=SCALES_LEVELS):
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


def conv(in_channels, out_channels, kernel_size=5, stride=2):
    return nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=kernel_size // 2)


def deconv(in_channels, out_channels, kernel_size=5, stride=2):
    return nn.ConvTranspose2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                              output_padding=stride - 1, padding=kernel_size // 2)


def _resize_buffers(module, prefix, names, state_dict):
    """Entropy-model tables have data-dependent sizes: give the registered (empty) buffers the
    checkpoint's shapes before nn.Module.load_state_dict (models/utils.py:46-111, policy resize_if_empty)."""
    for name in names:
        key = f"{prefix}.{name}"
        if key not in state_dict:
            continue
        buf = getattr(module, name)
        if buf.numel() == 0:
            setattr(module, name, buf.new_zeros(state_dict[key].shape))


class CompressionModel(nn.Module):
    def __init__(self, init_weights=True):
        super().__init__()
        # The reference runs its kaiming pass here, before any sub-module exists, so it is a no-op
        # (SURVEY.md section 8d); layers keep PyTorch's default init.  Kept a no-op on purpose.

    def aux_loss(self):
        return sum(m.loss() for m in self.modules() if isinstance(m, EntropyBottleneck))

    # ------------------------------------------------------------------ CUDA-graph plan validity
    # A captured plan bakes in raw pointers to (and, for the scale table, values of) buffers DERIVED from the
    # parameters: packed weight images, the EntropyBottleneck's pre-activated parameter block, medians, the host scale
    # table passed by value.  They are rebuilt -- and the old ones freed -- when a parameter changes (optimizer.step(),
    # load_state_dict(), update(force=True), .to(), set_precision()).  Every plan cache is therefore tagged with a
    # "weights epoch": storage pointer + in-place version counter of every parameter and buffer, plus the GEMM precision
    # mode.  A mismatch drops all plans; the next call re-captures from the live weights.
    def _weights_epoch(self):
        ts = self.__dict__.get("_epoch_tensors")
        if ts is None:   # the module tree walk is the slow part: cached until the next _drop_plans()
            ts = self.__dict__["_epoch_tensors"] = list(self.parameters()) + list(self.buffers())
        return hash((ops.precision_code(), tuple([t.data_ptr() for t in ts]), tuple([t._version for t in ts])))

    def _plans(self, name):
        """Plan cache `name` for the current weights epoch (all caches are dropped when the epoch changed)."""
        d = self.__dict__
        epoch = self._weights_epoch()
        if d.get("_plan_epoch") != epoch:
            self._drop_plans()
            d["_plan_epoch"] = epoch
        return d.setdefault(name, {})

    @staticmethod
    def _plan_slot(plans, key, cap):
        """LRU bookkeeping for a plan cache: returns True when `key` is cached (and marks it most recently used),
        else evicts the least recently used entries down to cap - 1 (mixed-size workloads keep their hot shapes
        instead of re-capturing everything whenever a fifth shape shows up)."""
        if key in plans:
            plans[key] = plans.pop(key)
            return True
        while len(plans) >= cap:
            plans.pop(next(iter(plans)))
        return False

    def _drop_plans(self):
        for k in ("_fwd_plans", "_enc_plans", "_dec_plans", "_plan_epoch", "_epoch_tensors"):
            self.__dict__.pop(k, None)

    def train(self, mode=True):
        if mode != self.training:
            self._drop_plans()      # an optimizer step between two eval phases changes the weights in place
        return super().train(mode)

    def _apply(self, fn, *args, **kwargs):
        self._drop_plans()          # .to() / .cuda() / .float(): new storages
        return super()._apply(fn, *args, **kwargs)

    def update(self, force=False):
        self._drop_plans()
        updated = False
        for m in self.children():
            if isinstance(m, EntropyBottleneck):
                updated |= m.update(force=force)
        return updated

    def load_state_dict(self, state_dict, strict=True):
        self._drop_plans()
        _resize_buffers(self.entropy_bottleneck, "entropy_bottleneck", ["_quantized_cdf", "_offset", "_cdf_length"],
                        state_dict)
        return super().load_state_dict(state_dict, strict=strict)


def _stack5(c_in):
    """The five-layer 3x3 stack shared by cc_mean / cc_scale / lrp transforms (stf.py:510-548)."""
    return nn.Sequential(conv(c_in, 224, stride=1, kernel_size=3), nn.GELU(), conv(224, 176, stride=1, kernel_size=3),
                         nn.GELU(), conv(176, 128, stride=1, kernel_size=3), nn.GELU(),
                         conv(128, 64, stride=1, kernel_size=3), nn.GELU(), conv(64, 32, stride=1, kernel_size=3))


class _ConvPacks:
    """Packed weight images of the convolutions that run on stf_conv2d, one per (conv module, source split, precision),
    rebuilt when the module's weight / bias changed (storage pointer or in-place version counter)."""

    def __init__(self):
        self.cache = {}

    def get(self, conv, src_channels, shuffle):
        prec = ops.conv_precision_code()
        key = (id(conv), tuple(src_channels), shuffle, prec)
        w, b = conv.weight, conv.bias
        ver = (w.data_ptr(), w._version, None if b is None else (b.data_ptr(), b._version))
        hit = self.cache.get(key)
        if hit is None or hit[0] != ver:
            if conv.kernel_size[0] != conv.kernel_size[1] or conv.padding[0] != conv.kernel_size[0] // 2 or \
                    conv.stride[0] != conv.stride[1] or conv.groups != 1 or conv.dilation[0] != 1:
                raise ValueError(f"stf_conv2d: unsupported convolution {conv}")
            hit = (ver, ops.PackedConv(w, b, src_channels, stride=conv.stride[0], pixel_shuffle=shuffle, prec=prec))
            self.cache[key] = hit
        return hit[1]


def _stack_layers(seq):
    """nn.Sequential of Conv2d / GELU / Sequential(Conv2d, PixelShuffle) -> [(conv, pixel_shuffle_factor, gelu_follows)]."""
    mods, out, i = list(seq), [], 0
    while i < len(mods):
        m = mods[i]
        shuffle = 0
        if isinstance(m, nn.Sequential) and len(m) == 2 and isinstance(m[0], nn.Conv2d) and isinstance(m[1], nn.PixelShuffle):
            conv, shuffle = m[0], m[1].upscale_factor
        elif isinstance(m, nn.Conv2d):
            conv = m
        else:
            raise ValueError(f"unexpected module in a conv stack: {type(m).__name__}")
        gelu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.GELU)
        out.append((conv, shuffle, gelu))
        i += 2 if gelu else 1
    return out


class _SliceCodec(CompressionModel):
    """Hyperprior + channel-conditional slice loop shared by STF and WACNN
    (stf.py:600-636, 687-735, 737-779 == cnn.py:144-183, 223-267, 289-327)."""

    slice_channels = 32

    # supplied by subclasses -----------------------------------------------------------------
    def _analysis(self, x):
        raise NotImplementedError

    def _synthesis(self, y_hat):
        raise NotImplementedError

    def _analysis_nhwc(self, x):
        """Inference: the latent as (B, h, w, M) NHWC."""
        return self._analysis(x).permute(0, 2, 3, 1).contiguous()

    def _synthesis_nhwc(self, y_hat):
        """Inference: x_hat from the NHWC (B, h, w, M) y_hat buffer."""
        return self._synthesis(y_hat.permute(0, 3, 1, 2))

    # shared ---------------------------------------------------------------------------------
    def update(self, scale_table=None, force=False):
        if scale_table is None:
            scale_table = get_scale_table()
        updated = self.gaussian_conditional.update_scale_table(scale_table, force=force)
        updated |= super().update(force=force)
        return updated

    def load_state_dict(self, state_dict, strict=True):
        _resize_buffers(self.gaussian_conditional, "gaussian_conditional",
                        ["_quantized_cdf", "_offset", "_cdf_length", "scale_table"], state_dict)
        return super().load_state_dict(state_dict, strict=strict)

    @classmethod
    def from_state_dict(cls, state_dict):
        net = cls()
        net.load_state_dict(state_dict)
        return net

    # ------------------------------------------------------------------ slice-loop building blocks
    # Everything between the analysis and the synthesis transform is NHWC (pixel-major) and runs on our own kernels:
    # stf_conv2d for every convolution (the torch.cat in front of a stack is a multi-source gather, bias + GELU, the pixel
    # shuffle and the LRP tail are its epilogues) and stf_slice_step_nhwc for the entropy steps.  All y_hat slices live in
    # ONE (B, h, w, M) buffer: slice i's slot is channels [32i, 32i+32), the support of slice i is the channel prefix
    # [0, 32 * min(i, max_support)) of the same buffer -- no concatenation, no layout copies, and after the last slice the
    # buffer IS the synthesis transform's token-major input.  No cuDNN in here: per-image results do not depend on the
    # batch they were computed in (stf.py:767 needs the decoder's indexes to equal the encoder's bit for bit).
    _CL = torch.channels_last

    def _prepare_inference(self):
        """One-time: conv weights to channels_last (values unchanged, strides only) for the cuDNN convolutions that
        remain (training path, WACNN g_a / g_s, the 3-channel image-side convolutions)."""
        if self.__dict__.get("_prepared"):
            return
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)) and m.weight.dim() == 4:
                m.weight.data = m.weight.data.contiguous(memory_format=self._CL)
        self.__dict__["_prepared"] = True
        self._drop_plans()

    def _conv_stack(self, seq, srcs, out=None, last_act=None):
        """Run a conv stack on NHWC sources through stf_conv2d.  srcs: list of (B, h, w, C_s) views = the channel concat
        the reference builds with torch.cat; out: optional NHWC destination view of the last layer; last_act: "lrp" makes
        the last layer compute out <- out + 0.5 * tanh(conv) in place (stf.py:631-633)."""
        packs = self.__dict__.setdefault("_conv_packs", _ConvPacks())
        layers = _stack_layers(seq)
        x = list(srcs)
        for li, (conv, shuffle, gelu) in enumerate(layers):
            last = li == len(layers) - 1
            pc = packs.get(conv, [t.shape[3] for t in x], shuffle)
            act = True if gelu else (last_act if last else False)
            x = [ops.conv2d(x, pc, act=act, out=out if last else None)]
        return x[0]

    def _support(self, y_hat, i):
        """Channel prefix of the y_hat buffer that slice i is conditioned on (stf.py:608-611); None for slice 0."""
        ns = i if self.max_support_slices < 0 else min(i, self.max_support_slices)
        return y_hat[..., : self.slice_channels * ns] if ns else None

    def _slice_params(self, i, latent_means, latent_scales, y_hat):
        """mu and scale of slice i, NHWC (B, h, w, 32).  The two five-layer conv stacks are independent (stf.py:615-621):
        the scale stack runs on a side stream, forked and joined with events (parallel branches of the CUDA graph)."""
        sup = self._support(y_hat, i)
        extra = [] if sup is None else [sup]
        main = torch.cuda.current_stream()
        side = self._side_stream(main.device)
        side.wait_stream(main)
        old_cap, ops.MAX_CTAS = ops.MAX_CTAS, _SLICE_CTAS or ops.MAX_CTAS
        try:
            with torch.cuda.stream(side):
                scale = self._conv_stack(self.cc_scale_transforms[i], [latent_scales] + extra)
            mu = self._conv_stack(self.cc_mean_transforms[i], [latent_means] + extra)
        finally:
            ops.MAX_CTAS = old_cap
        main.wait_stream(side)
        scale.record_stream(main)
        return mu, scale

    def _side_stream(self, device):
        streams = self.__dict__.setdefault("_side_streams", {})
        key = str(device)
        if key not in streams:
            streams[key] = torch.cuda.Stream(device=device)
        return streams[key]

    def _lrp(self, i, latent_means, y_hat):
        """y_hat slot i <- y_hat_i + 0.5 * tanh(lrp_i(cat([latent_means, support, y_hat_i])))  in place."""
        Cs = self.slice_channels
        slot = y_hat[..., Cs * i: Cs * (i + 1)]
        ns = i if self.max_support_slices < 0 else min(i, self.max_support_slices)
        if ns == i:      # the slot directly follows the support: one contiguous channel range
            srcs = [latent_means, y_hat[..., : Cs * (i + 1)]]
        else:
            srcs = [latent_means, y_hat[..., : Cs * ns], slot]
        self._conv_stack(self.lrp_transforms[i], srcs, out=slot, last_act="lrp")

    def _needed_as_support(self, i):
        return self.max_support_slices < 0 or i < self.max_support_slices

    def _hyper_analysis(self, y):
        """h_a on the NHWC latent -> z as (B, C, zh, zw) NCHW (18 k elements per image: the layout the bottleneck kernel and
        the z coding order use)."""
        return self._conv_stack(self.h_a, [y]).permute(0, 3, 1, 2).contiguous()

    def _hyper_synthesis(self, z_hat):
        """z_hat (B, C, zh, zw) NCHW -> (latent_scales, latent_means), NHWC (B, h, w, M)."""
        z = z_hat.permute(0, 2, 3, 1).contiguous()
        return self._conv_stack(self.h_scale_s, [z]), self._conv_stack(self.h_mean_s, [z])

    def _check_latent(self, latent, y_shape):
        if tuple(latent.shape) != tuple(y_shape):
            raise ValueError(f"hyper-synthesis output {tuple(latent.shape)} does not match the latent {tuple(y_shape)}: "
                             "pad the image to a multiple of 64 (the reference's torch.cat fails the same way, stf.py:612)")

    def forward(self, x, noise=None):
        """eval / no_grad: the fused inference path.  train() with grad enabled: the training forward of
        stf.py:584-648 / cnn.py:141-189 ("noise" quantisation for the likelihoods, ste_round for y_hat / z_hat) with
        backward through stf_b200/autograd.py.  `noise` (optional dict {"y": (B,M,h,w), "z": (B,192,h',w')} of
        U(-1/2, 1/2) tensors) injects the quantisation noise for parity runs."""
        if self.training:     # chosen on `training` alone, like the reference (noise + DropPath also under no_grad)
            return self._forward_train(x, noise)
        if torch.is_grad_enabled() and x.requires_grad:
            raise RuntimeError("eval-mode forward is the inference path (fused kernels, no autograd graph): call "
                               ".train() for a differentiable forward, or detach the input")
        with torch.no_grad():
            if not (x.is_cuda and self._graphs_enabled()):
                return self._forward_eval(x)
            # one CUDA graph per input shape (~340 launches of ours + cuDNN, launch-bound when issued eagerly); the
            # returned tensors are clones: the graph's static outputs are overwritten by the next call
            self._prepare_inference()          # (before the plan lookup: it re-lays-out conv weights, i.e. changes the epoch)
            plans = self._plans("_fwd_plans")
            key = tuple(x.shape)
            if not self._plan_slot(plans, key, 8):

                def fn(t):
                    o = self._forward_eval(t)
                    return o["x_hat"], o["likelihoods"]["y"], o["likelihoods"]["z"], o.get("y")
                plans[key] = graphs.Segment(lambda t: tuple(v for v in fn(t) if v is not None), [x.contiguous()])
            outs = [v.clone() for v in plans[key](x.contiguous())]
            res = {"x_hat": outs[0], "likelihoods": {"y": outs[1], "z": outs[2]}}
            if hasattr(self, "is_teacher"):
                res["y"] = outs[3] if self.is_teacher and len(outs) > 3 else None
            return res

    def _forward_train(self, x, noise=None):
        from . import autograd as AG
        self._prepare_inference()          # conv weights in channels_last (values unchanged): cuDNN runs NHWC fwd and bwd
        y = self._analysis(x).contiguous(memory_format=self._CL)
        hw = y.shape[2:]
        z = self.h_a(y)
        eb, gc = self.entropy_bottleneck, self.gaussian_conditional
        _, z_likelihoods = eb(z, noise=None if noise is None else noise.get("z"))
        z_offset = eb._get_medians().reshape(1, -1, 1, 1)
        z_hat = AG.ste_round(z - z_offset) + z_offset
        latent_scales, latent_means = self.h_scale_s(z_hat), self.h_mean_s(z_hat)
        Cs = self.slice_channels
        y_hat_slices, y_likelihood = [], []
        for i, y_slice in enumerate(y.chunk(self.num_slices, 1)):
            support = y_hat_slices if self.max_support_slices < 0 else y_hat_slices[: self.max_support_slices]
            mean_support = torch.cat([latent_means] + support, dim=1)
            mu = self.cc_mean_transforms[i](mean_support)[:, :, : hw[0], : hw[1]]
            scale = self.cc_scale_transforms[i](torch.cat([latent_scales] + support, dim=1))[:, :, : hw[0], : hw[1]]
            n_i = (torch.empty_like(y_slice).uniform_(-0.5, 0.5) if noise is None
                   else noise["y"][:, i * Cs:(i + 1) * Cs])
            lik = AG.GaussianLikelihoodTrain.apply(y_slice, scale, mu, n_i, gc.scale_bound_value(),
                                                   gc._likelihood_bound if gc.use_likelihood_bound else 0.0)
            y_likelihood.append(lik)
            y_hat_slice = AG.ste_round(y_slice - mu) + mu
            lrp = self.lrp_transforms[i](torch.cat([mean_support, y_hat_slice], dim=1))
            y_hat_slices.append(y_hat_slice + 0.5 * torch.tanh(lrp))
        y_hat = torch.cat(y_hat_slices, dim=1)
        out = {"x_hat": self._synthesis(y_hat),
               "likelihoods": {"y": torch.cat(y_likelihood, dim=1), "z": z_likelihoods}}
        if hasattr(self, "is_teacher"):
            out["y"] = y if self.is_teacher else None
        return out

    def _forward_eval(self, x):
        self._prepare_inference()
        y = self._analysis_nhwc(x)                                    # (B, h, w, M)
        B, h, w, M = y.shape
        z = self._hyper_analysis(y)
        eb = self.entropy_bottleneck
        z_hat, z_likelihoods, _ = ops.entropy_bottleneck(z, eb.packed_params(), lik_bound=eb._likelihood_bound,
                                                         ste_round=True)
        latent_scales, latent_means = self._hyper_synthesis(z_hat)
        self._check_latent(latent_means, y.shape)
        gc = self.gaussian_conditional
        Cs = self.slice_channels
        y_hat = torch.empty_like(y)
        y_lik = torch.empty((B, M, h, w), dtype=torch.float32, device=y.device)
        for i in range(self.num_slices):
            mu, scale = self._slice_params(i, latent_means, latent_scales, y_hat)
            ops.slice_step_nhwc(y=y[..., Cs * i: Cs * (i + 1)], scales=scale, means=mu, y_hat=y_hat[..., Cs * i: Cs * (i + 1)],
                                likelihood=y_lik, lik_offset=Cs * i, scale_bound=gc.scale_bound_value(),
                                lik_bound=gc._likelihood_bound, ste_round=True)
            self._lrp(i, latent_means, y_hat)
        out = {"x_hat": self._synthesis_nhwc(y_hat), "likelihoods": {"y": y_lik, "z": z_likelihoods}}
        if hasattr(self, "is_teacher"):
            out["y"] = y.permute(0, 3, 1, 2).contiguous() if self.is_teacher else None
        return out

    # ------------------------------------------------------------------ encoder
    def _encode_gpu(self, x, keep=None, narrow=False):
        """All device work of compress(): x -> (y symbols, y indexes, z symbols, overflow flag); y symbols / indexes are
        (B, n) int32 / int32, or int16 / uint8 with `narrow` (the flag, one int32, is then set when a value did not fit
        and the caller repeats the call wide); z symbols (B, C, h, w) int32.
        No host synchronisation inside: capturable as one CUDA graph."""
        gc, eb = self.gaussian_conditional, self.entropy_bottleneck
        with _phase("enc.analysis"):
            y = self._analysis_nhwc(x)
        B, h, w, M = y.shape
        with _phase("enc.hyper"):
            z = self._hyper_analysis(y)
            # EntropyBottleneck.compress + decompress (stf.py:688-689): decompress(z_strings) is
            # dequantize(symbols, medians), which the same kernel emits -- no need to decode our own stream
            z_hat, _, z_sym = ops.entropy_bottleneck(z, eb.packed_params(), want_lik=False, want_symbols=True)
            latent_scales, latent_means = self._hyper_synthesis(z_hat)
            self._check_latent(latent_means, y.shape)
        Cs, plane = self.slice_channels, h * w
        sym = torch.empty((B, M * plane), dtype=torch.int16 if narrow else torch.int32, device=y.device)
        idx = torch.empty((B, M * plane), dtype=torch.uint8 if narrow else torch.int32, device=y.device)
        ovf = torch.zeros(1, dtype=torch.int32, device=y.device)
        table = gc.host_scale_table()
        y_hat = torch.empty_like(y)
        with _phase("enc.slices"):
            for i in range(self.num_slices):
                mu, scale = self._slice_params(i, latent_means, latent_scales, y_hat)
                need = self._needed_as_support(i) or keep is not None
                ops.slice_step_nhwc(y=y[..., Cs * i: Cs * (i + 1)], scales=scale, means=mu, symbols_out=sym, indexes_out=idx,
                                    out_offset=i * Cs * plane, y_hat=y_hat[..., Cs * i: Cs * (i + 1)] if need else None,
                                    table=table, scale_bound=gc.scale_bound_value(), overflow=ovf)
                if need:   # later slices are never read again in compress(): their LRP stacks are dead work
                    self._lrp(i, latent_means, y_hat)
        if keep is not None:
            keep.update(y=y.permute(0, 3, 1, 2).contiguous(), z=z, y_hat=y_hat.permute(0, 3, 1, 2).contiguous())
        return sym, idx, z_sym, ovf

    def _graphs_enabled(self):
        return self.__dict__.get("cuda_graphs", _GRAPHS_DEFAULT) and PHASE_TIMES is None

    @staticmethod
    def _parts(B, pipelined, n_parts=None, weights=None):
        """Image ranges coded as independent sub-batches.  Two parts let the host rANS work of one part overlap
        the device work of the other (the device runs part B while the host codes part A, and vice versa); from 48
        images three parts shorten the exposed rANS tail of the last part further while each part still fills the GPU
        (batch 64 on one B200, same box: 108.8 -> 115.8 Mpixel/s; four parts of 16: 109.8)."""
        if not pipelined or B < _PIPELINE_MIN_BATCH:
            return [(0, B)]
        n = max(2, min(n_parts or _PIPELINE_PARTS or (3 if B >= 48 else 2), B // 8 if B >= 16 else 2))
        # tapered sizes (weights 1 .. _PART_TAPER from the first to the last part): the host work of the LAST part is the
        # exposed tail of compress(), the device work of the FIRST part runs with an idle host
        wts = [1.0 + (_PART_TAPER - 1.0) * i / (n - 1) for i in range(n)]
        if weights and B >= 8 * len(weights):
            wts = list(weights)
        acc, bounds = 0.0, [0]
        for w_ in wts:
            acc += w_
            bounds.append(round(B * acc / sum(wts)))
        return [(lo, hi) for lo, hi in zip(bounds, bounds[1:]) if hi > lo]

    def _enc_parts(self, B, pipelined=True):
        """compress(): 7 : 6 : 3 from 48 images (batch 64 -> 28 / 24 / 12).  The host codes part k while the device runs part
        k + 1, so only the LAST part's rANS is exposed: a small last part, not so small that its device work stops filling the
        GPU (measured at batch 64, 16 / 4 host threads: 24/22/18 82.9 / 86.2 ms, 28/24/12 78.6 / 84.2 ms, 12/20/20/12 80.9 / 83.4)."""
        return self._parts(B, pipelined, weights=_ENC_SPLIT or ((7.0, 6.0, 3.0) if B >= 48 else None))

    def _dec_parts(self, B, pipelined=True):
        """decompress(): tapered thirds; with few host threads (eight ranks sharing a host) the slice loop is bound by the
        host decoder, and four parts of 4 : 5 : 4 : 3 give the device its synthesis work earlier and leave a shorter tail
        (4 threads, batch 64: 102.7 -> 97.5 ms; 16 threads: no difference)."""
        w = _DEC_SPLIT
        if w is None and _DEC_PARTS is None and B >= 56 and ans.default_threads() < 8:
            w = (4.0, 5.0, 4.0, 3.0)
        return self._parts(B, pipelined, len(w) if w else _DEC_PARTS, weights=w)

    @torch.no_grad()
    def compress(self, x, debug=None, _wide=False):
        gc, eb = self.gaussian_conditional, self.entropy_bottleneck
        y_table, z_table = gc.rans_table(), eb.rans_table()
        self._prepare_inference()
        x = x.contiguous()
        use_graphs = debug is None and self._graphs_enabled()
        narrow = _NARROW and not _wide and debug is None and gc.host_scale_table().size <= 256
        parts = self._enc_parts(x.shape[0], use_graphs)
        stream = torch.cuda.current_stream()
        pending = []
        plans = self._plans("_enc_plans") if use_graphs else None      # (validated against the weights epoch once per call)
        for slot, (lo, hi) in enumerate(parts):
            xp = x[lo:hi]
            with _trace("enc.gpu", slot, 0, device=True):
                if use_graphs:
                    key = tuple(xp.shape) + (narrow,)
                    if not self._plan_slot(plans, key, 8):
                        plans[key] = graphs.Segment(lambda t: self._encode_gpu(t, narrow=narrow), [xp])
                    sym, idx, z_sym, ovf = plans[key](xp)
                else:
                    sym, idx, z_sym, ovf = self._encode_gpu(xp, keep=debug, narrow=narrow)
            Bp, total = sym.shape
            with _phase("enc.d2h"), _trace("enc.d2h", slot, 0, device=True):
                sym_h, idx_h = self._host_buffers(("y", slot), Bp, total, (sym.dtype, idx.dtype))
                zsym_h, ovf_h = self._host_buffers(("z", slot), Bp, z_sym[0].numel())
                sym_h.copy_(sym, non_blocking=True)      # stream-ordered before the next part's replay
                idx_h.copy_(idx, non_blocking=True)      # overwrites the plan's static outputs
                zsym_h.copy_(z_sym.reshape(Bp, -1), non_blocking=True)
                ovf_h[0, :1].copy_(ovf, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(stream)
            pending.append((ev, Bp, sym_h, idx_h, zsym_h, tuple(z_sym.shape), ovf_h))
        y_strings, z_strings = [], []
        for slot, (ev, Bp, sym_h, idx_h, zsym_h, zshape, ovf_h) in enumerate(pending):
            with _phase("enc.d2h"), _trace("enc.wait", slot):
                ev.synchronize()
            if narrow and int(ovf_h[0, 0]) != 0:
                # a symbol beyond int16 (latents of natural images are O(100)): code this call again on int32 buffers
                stream.synchronize()
                return self.compress(x, _wide=True)
            if debug is not None:
                debug.update(symbols=sym_h.clone(), indexes=idx_h.clone(), z_symbols=zsym_h.clone())
            with _phase("enc.rans"), _trace("enc.rans", slot):
                z_idx = self._z_indexes(tuple(zshape))
                scratch = self.__dict__.setdefault("_enc_scratch", {})
                y_strings += ans.encode_rows(y_table, sym_h.numpy(), idx_h.numpy(), scratch.setdefault(("y", slot), {}))
                z_strings += ans.encode_rows(z_table, zsym_h.numpy(), z_idx, scratch.setdefault(("z", slot), {}))
        return {"strings": [y_strings, z_strings], "shape": torch.Size(zshape[-2:])}

    def _z_indexes(self, zshape):
        """EntropyBottleneck._build_indexes(size) as a cached (B, C*h*w) int32 numpy array (channel ids; entropy_models.py:517-522)."""
        cache = self.__dict__.setdefault("_z_idx_cache", {})
        if zshape not in cache:
            if len(cache) >= 8:
                cache.clear()
            cache[zshape] = self.entropy_bottleneck._build_indexes(zshape).reshape(zshape[0], -1).numpy().copy()
        return cache[zshape]

    def _host_buffers(self, tag, B, n, dtypes=(torch.int32, torch.int32)):
        """Pinned staging buffer pair (int32 unless `dtypes` says otherwise), cached per (tag, shape, dtypes): cudaHostAlloc
        is slow (1.8 ms per call with eight ranks on one host), so every user has its own tag -- compress() and decompress()
        shared one in round 1 and evicted each other's buffers every call."""
        cache = self.__dict__.setdefault("_pinned", {})
        key = (tag, B, n, tuple(dtypes))
        if key not in cache:
            for k in [k for k in cache if k[0] == tag]:
                del cache[k]
            cache[key] = (torch.empty((B, n), dtype=dtypes[0], pin_memory=True),
                          torch.empty((B, n), dtype=dtypes[1], pin_memory=True))
        return cache[key]

    # ------------------------------------------------------------------ decoder
    # The decoder alternates device segments and host rANS decodes (12 hand-offs, stf.py:757-779):
    #   seg 0        : z symbols -> z_hat -> hyper synthesis -> slice-0 parameters -> indexes 0
    #   seg i (1..S-1): symbols i-1 -> y_hat i-1 (+LRP) -> slice-i parameters -> indexes i
    #   seg S        : symbols S-1 -> y_hat S-1 (+LRP) -> synthesis transform -> x_hat
    def _dec_first(self, st, z_sym):
        eb = self.entropy_bottleneck
        med = eb.packed_params()[:, _EB_MEDIAN]
        B, C = z_sym.shape[0], z_sym.shape[1]
        medians = med.reshape(1, C, 1).expand(B, C, z_sym[0, 0].numel()).contiguous()
        z_hat = ops.dequantize(z_sym.reshape(B, -1), 0, medians).reshape(z_sym.shape)
        st["scales"], st["means"] = self._hyper_synthesis(z_hat)
        st["y_hat"] = torch.empty_like(st["means"])
        st["ovf"] = torch.zeros(1, dtype=torch.int32, device=z_hat.device)   # (never set: <= 256 scale levels, checked by the caller)
        return self._dec_params(st, 0)

    def _dec_params(self, st, i):
        """Parameters of slice i -> its indexes (B, 32 * h * w) in coding order; mu is kept for the dequantize step."""
        gc = self.gaussian_conditional
        # per-slice entries keep every segment idempotent on `st` (graph warm-up runs a segment several times)
        st["mu", i], scale = self._slice_params(i, st["means"], st["scales"], st["y_hat"])
        B, h, w, Cs = scale.shape
        idx = torch.empty((B, Cs * h * w), dtype=torch.uint8 if st.get("narrow") else torch.int32, device=scale.device)
        ops.slice_step_nhwc(scales=scale, indexes_out=idx, table=gc.host_scale_table(), scale_bound=gc.scale_bound_value(),
                            overflow=st["ovf"])
        return idx

    def _dec_slice(self, st, i, sym_prev):
        """Finish slice i-1: y_hat = symbols + mu into its slot, then the LRP update in place.  Idempotent on `st`: the slot
        is rewritten from the symbols first (graph warm-up runs a segment more than once)."""
        Cs = self.slice_channels
        ops.slice_step_nhwc(symbols_in=sym_prev, means=st["mu", i - 1], y_hat=st["y_hat"][..., Cs * (i - 1): Cs * i])
        self._lrp(i - 1, st["means"], st["y_hat"])

    def _dec_mid(self, st, i, sym_prev):
        self._dec_slice(st, i, sym_prev)
        return self._dec_params(st, i)

    def _dec_last(self, st, sym_prev):
        self._dec_slice(st, self.num_slices, sym_prev)
        return self._synthesis_nhwc(st["y_hat"]).clamp_(0, 1)

    def _decode_plan(self, plans, slot, B, C, zh, zw, device, narrow=False):
        """13 CUDA-graph segments on one shared memory pool, captured once per (slot, B, z shape)."""
        key = (slot, B, zh, zw, narrow)
        if self._plan_slot(plans, key, 12):
            return plans[key]
        h, w = zh * 4, zw * 4
        n = self.slice_channels * h * w
        st = {"hw": (h, w), "narrow": narrow}
        z0 = torch.zeros((B, C, zh, zw), dtype=torch.int32, device=device)
        s0 = torch.zeros((B, n), dtype=torch.int32, device=device)
        segs = [graphs.Segment(lambda t: self._dec_first(st, t), [z0])]
        pool = segs[0].pool()
        for i in range(1, self.num_slices):
            segs.append(graphs.Segment(lambda t, i=i: self._dec_mid(st, i, t), [s0], pool=pool))
        segs.append(graphs.Segment(lambda t: self._dec_last(st, t), [s0], pool=pool))
        plans[key] = (segs, st)
        return plans[key]

    @torch.no_grad()
    def decompress(self, strings, shape):
        assert isinstance(strings, list) and len(strings) == 2
        gc, eb = self.gaussian_conditional, self.entropy_bottleneck
        y_table, z_table = gc.rans_table(), eb.rans_table()
        self._prepare_inference()
        B = len(strings[1])
        if len(strings[0]) != B:
            raise ValueError(f"{len(strings[0])} y-strings for {B} z-strings (one y-string per image expected)")
        device = eb._quantized_cdf.device
        C = eb._quantized_cdf.size(0)
        zh, zw = int(shape[0]), int(shape[1])
        h, w = zh * 4, zw * 4
        n = self.slice_channels * h * w
        S = self.num_slices
        main = torch.cuda.current_stream()
        use_graphs = self._graphs_enabled()
        narrow = _NARROW and gc.host_scale_table().size <= 256       # indexes come back as uint8
        if _DEVICE_DECODE and PHASE_TIMES is None:
            out = self._decompress_device(strings, B, C, zh, zw, device, use_graphs)
            if out is not None:
                return out

        class Part:
            pass

        # Every sub-batch decodes on its own CUDA stream, and the host serves them depth-first: the most advanced part
        # whose device segment has finished gets its next slice decoded and its next segment launched; only when no part
        # is ready does the host block (on the most advanced one).  The parts therefore drift apart instead of marching in
        # lockstep, and the long last segment of an early part (slice 11 + the whole synthesis transform) runs on the
        # device while the host is still decoding the slices of the later parts -- in lockstep the three synthesis
        # transforms ran back to back at the end with the host idle.
        parts = []
        plans = self._plans("_dec_plans") if use_graphs else None
        streams = self.__dict__.setdefault("_part_streams", {})
        for slot, (lo, hi) in enumerate(self._dec_parts(B, use_graphs)):
            p = Part()
            p.B = hi - lo
            p.stream = main if not use_graphs else streams.setdefault((str(device), slot), torch.cuda.Stream(device=device))
            p.segs, p.st = self._decode_plan(plans, slot, p.B, C, zh, zw, device, narrow) if use_graphs else \
                (None, {"hw": (h, w), "narrow": narrow})
            p.decoders = _decoders(strings[0][lo:hi])
            p.sym_h, p.idx_h = self._host_buffers(("y_dec", slot), p.B, n, (torch.int32, torch.uint8 if narrow else torch.int32))
            p.zsym_h, _ = self._host_buffers(("z_dec", slot), p.B, C * zh * zw)
            p.ev = torch.cuda.Event()
            p.plan = ans.DecodePlan(p.decoders, y_table, p.idx_h.numpy(), p.sym_h.numpy())
            p.next = 1                      # next segment to launch (segment i consumes the symbols of slice i - 1)
            with _phase("dec.hyper"), _trace("dec.first", slot):
                z_np = p.zsym_h.numpy()
                z_idx = self._z_indexes((p.B, C, zh, zw))
                ans.decode_batch(_decoders(strings[1][lo:hi]), z_table, [z_idx[b] for b in range(p.B)],
                                 outs=[z_np[b] for b in range(p.B)])
                p.stream.wait_stream(main)
                with torch.cuda.stream(p.stream):
                    if use_graphs:
                        idx = p.segs[0](p.zsym_h.reshape(p.B, C, zh, zw))
                    else:
                        idx = self._dec_first(p.st, p.zsym_h.to(device, non_blocking=True).reshape(p.B, C, zh, zw))
                    p.idx_h.copy_(idx.reshape(p.B, n), non_blocking=True)
                    p.ev.record(p.stream)
            parts.append(p)
        outs = [None] * len(parts)
        todo = len(parts)
        def may_lead(k):
            """Part k may run at most _DEC_LEAD slices ahead of the next unfinished part: enough stagger for its synthesis
            to overlap the later parts' slices, not so much that the last part ends up alone (host and device would then
            alternate with nothing to overlap: measured with 4 host threads per rank)."""
            later = next((q for q in parts[k + 1:] if q.next <= S), None)
            return later is None or parts[k].next - later.next < _DEC_LEAD

        while todo:
            live = [k for k, p in enumerate(parts) if p.next <= S]
            ready = next((parts[k] for k in live if may_lead(k) and parts[k].ev.query()), None)
            if ready is None:
                ready = next((parts[k] for k in live if parts[k].ev.query()), None)
            if ready is None:               # nothing has landed yet: wait for the most advanced part that may proceed
                ready = next((parts[k] for k in live if may_lead(k)), parts[live[0]])
                with _phase("dec.slices.gpu"), _trace("dec.wait", parts.index(ready), ready.next):
                    ready.ev.synchronize()
            p, i = ready, ready.next
            last = i == S
            with _phase("dec.slices.rans"), _trace("dec.rans", parts.index(p), i):
                p.plan.run()                # decoders advance by one slice: pinned indexes -> pinned symbols
            with _phase("dec.synthesis" if last else "dec.slices.gpu"), torch.cuda.stream(p.stream), \
                    _trace("dec.seg", parts.index(p), i, device=True):
                if use_graphs:
                    out = p.segs[i](p.sym_h)          # pinned host buffer -> the segment's static input
                else:
                    sym = p.sym_h.to(device, non_blocking=True)
                    out = self._dec_last(p.st, sym) if last else self._dec_mid(p.st, i, sym)
                if last:
                    outs[parts.index(p)] = out
                    todo -= 1
                else:
                    p.idx_h.copy_(out.reshape(p.B, n), non_blocking=True)
                    p.ev.record(p.stream)
            p.next += 1
        for p in parts:
            main.wait_stream(p.stream)
        x_hat = torch.cat(outs, dim=0) if (len(outs) > 1 or use_graphs) else outs[0]
        return {"x_hat": x_hat}


    # ------------------------------------------------------------------ decoder, rANS on the device
    def _dec_all(self, st, ds, z_sym, reserve_sms=0):
        """The whole decode of one sub-batch with NO host round trip: hyper-synthesis, then per slice parameters -> indexes
        -> device rANS decode (lane b of one warp = image b) -> y_hat + LRP, then the synthesis transform.  The persistent
        kernels leave `reserve_sms` SMs to the decoding warps of the sub-batches that run beside this one."""
        y_table = self.gaussian_conditional.rans_table()
        old_cap, ops.MAX_CTAS = ops.MAX_CTAS, ops.NUM_SMS - reserve_sms     # (baked into the captured launches)
        try:
            ds.upload()
            idx = self._dec_first(st, z_sym)
            for i in range(1, self.num_slices + 1):
                sym = torch.empty_like(idx)
                ans.decode_device(y_table, ds, idx, sym, first=(i == 1))
                if i < self.num_slices:
                    idx = self._dec_mid(st, i, sym)
                else:
                    return self._dec_last(st, sym), ds.status
        finally:
            ops.MAX_CTAS = old_cap

    def _decompress_device(self, strings, B, C, zh, zw, device, use_graphs):
        """decompress() with the y-strings decoded on the device.  Every sub-batch is ONE CUDA graph on its own stream (z is
        decoded on the host first: 18 k symbols per image); the single decoding warp of a sub-batch runs concurrently with
        the other sub-batches' convolutions, and no host thread is needed during the slice loop -- which is what bounded
        the 8-GPU runs (4 host threads per rank).  Returns None when the strings do not fit the staging buffers of the cached
        plan shape (the caller then takes the host path)."""
        eb = self.entropy_bottleneck
        z_table = eb.rans_table()
        main = torch.cuda.current_stream()
        plans = self._plans("_dec_plans")
        streams = self.__dict__.setdefault("_part_streams", {})
        jobs = []
        part_list = self._dec_parts(B, use_graphs)
        reserve = len(part_list) if len(part_list) > 1 else 0
        for slot, (lo, hi) in enumerate(part_list):
            ys = strings[0][lo:hi]
            if any(len(s_) % 4 or len(s_) < 8 for s_ in ys):
                return None
            Bp = hi - lo
            need = sum(len(s_) for s_ in ys) // 4
            cap = 1 << max(16, (need + need // 4).bit_length())          # words: next power of two above 1.25 x this call
            key = ("dev", slot, Bp, zh, zw, cap, reserve)
            if not self._plan_slot(plans, key, 12):
                st = {"hw": (zh * 4, zw * 4)}
                ds = ans.DeviceStreams(Bp, cap, device)
                seg = None
                if use_graphs:
                    ds.load(ys)
                    z0 = torch.zeros((Bp, C, zh, zw), dtype=torch.int32, device=device)
                    seg = graphs.Segment(lambda t, st=st, ds=ds: self._dec_all(st, ds, t, reserve), [z0])
                plans[key] = (seg, st, ds)
            seg, st, ds = plans[key]
            ds.load(ys)
            zsym_h, _ = self._host_buffers(("z_dec", slot), Bp, C * zh * zw)
            z_np = zsym_h.numpy()
            z_idx = self._z_indexes((Bp, C, zh, zw))
            ans.decode_batch(_decoders(strings[1][lo:hi]), z_table, [z_idx[b] for b in range(Bp)],
                             outs=[z_np[b] for b in range(Bp)])
            stream = streams.setdefault((str(device), slot), torch.cuda.Stream(device=device)) if use_graphs else main
            stream.wait_stream(main)
            with torch.cuda.stream(stream):
                z_in = zsym_h.reshape(Bp, C, zh, zw)
                out, status = seg(z_in) if seg is not None else self._dec_all(st, ds, z_in.to(device, non_blocking=True), reserve)
                status_h = self.__dict__.setdefault("_dec_status", {}).setdefault(
                    (slot, Bp), torch.empty(Bp, dtype=torch.int32, pin_memory=True))
                status_h.copy_(status, non_blocking=True)
            jobs.append((stream, out, status_h))
        for stream, _, _ in jobs:
            main.wait_stream(stream)
        x_hat = torch.cat([o for _, o, _ in jobs], dim=0) if (len(jobs) > 1 or use_graphs) else jobs[0][1]
        main.synchronize()
        for _, _, status_h in jobs:
            bad = int(status_h.min())
            if bad != 0:
                raise ValueError("corrupt or truncated y bitstream" if bad == -6 else f"rANS device decoder failed ({bad})")
        return {"x_hat": x_hat}


def _decoders(strings):
    out = []
    for s in strings:
        d = ans.RansDecoder()
        d.set_stream(s)
        out.append(d)
    return out


class SymmetricalTransFormer(_SliceCodec):
    """STF: PatchEmbed -> 4 Swin stages -> hyperprior + 12-slice entropy model -> 4 Swin stages -> end_conv."""

    def __init__(self, pretrain_img_size=256, patch_size=2, in_chans=3, embed_dim=48, depths=[2, 2, 6, 2],
                 num_heads=[3, 6, 12, 24], window_size=4, num_slices=12, mlp_ratio=4.0, qkv_bias=True, qk_scale=None,
                 drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.2, norm_layer=nn.LayerNorm, patch_norm=True,
                 frozen_stages=-1, use_checkpoint=False, is_teacher=False):
        super().__init__()
        self.pretrain_img_size = pretrain_img_size
        self.num_layers = len(depths)
        self.embed_dim = embed_dim
        self.patch_norm = patch_norm
        self.frozen_stages = frozen_stages
        self.num_slices = num_slices
        self.max_support_slices = num_slices // 2
        self.is_teacher = is_teacher
        self.patch_embed = PatchEmbed(patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                      norm_layer=norm_layer if patch_norm else None)
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, sum(depths))]

        def stage(dim, i, depths_, heads_, resample, inverse):
            return BasicLayer(dim=dim, depth=depths_[i], num_heads=heads_[i], window_size=window_size,
                              mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate,
                              attn_drop=attn_drop_rate, drop_path=dpr[sum(depths_[:i]):sum(depths_[:i + 1])],
                              norm_layer=norm_layer, downsample=resample if i < self.num_layers - 1 else None,
                              use_checkpoint=use_checkpoint, inverse=inverse)

        self.layers = nn.ModuleList(stage(int(embed_dim * 2 ** i), i, depths, num_heads, PatchMerging, False)
                                    for i in range(self.num_layers))
        rd, rh = depths[::-1], num_heads[::-1]
        self.syn_layers = nn.ModuleList(stage(int(embed_dim * 2 ** (3 - i)), i, rd, rh, PatchSplit, True)
                                        for i in range(self.num_layers))
        self.end_conv = nn.Sequential(
            nn.Conv2d(embed_dim, embed_dim * patch_size ** 2, kernel_size=5, stride=1, padding=2),
            nn.PixelShuffle(patch_size), nn.Conv2d(embed_dim, 3, kernel_size=3, stride=1, padding=1))
        self.num_features = [int(embed_dim * 2 ** i) for i in range(self.num_layers)]
        self.g_a = None
        self.g_s = None
        M = embed_dim * 8
        self.h_a = nn.Sequential(conv3x3(M, 384), nn.GELU(), conv3x3(384, 336), nn.GELU(),
                                 conv3x3(336, 288, stride=2), nn.GELU(), conv3x3(288, 240), nn.GELU(),
                                 conv3x3(240, 192, stride=2))

        def hyper_synthesis():
            return nn.Sequential(conv3x3(192, 240), nn.GELU(), subpel_conv3x3(240, 288, 2), nn.GELU(),
                                 conv3x3(288, 336), nn.GELU(), subpel_conv3x3(336, 384, 2), nn.GELU(),
                                 conv3x3(384, M))

        self.h_mean_s = hyper_synthesis()
        self.h_scale_s = hyper_synthesis()
        ms = self.max_support_slices
        self.cc_mean_transforms = nn.ModuleList(_stack5(M + 32 * min(i, ms)) for i in range(num_slices))
        self.cc_scale_transforms = nn.ModuleList(_stack5(M + 32 * min(i, ms)) for i in range(num_slices))
        self.lrp_transforms = nn.ModuleList(_stack5(M + 32 * min(i + 1, ms + 1)) for i in range(num_slices))
        self.entropy_bottleneck = EntropyBottleneck(embed_dim * 4)
        self.gaussian_conditional = GaussianConditional(None)
        self._freeze_stages()

    def _freeze_stages(self):
        if self.frozen_stages >= 0:
            self.patch_embed.eval()
            for p in self.patch_embed.parameters():
                p.requires_grad = False
        if self.frozen_stages >= 2:
            self.pos_drop.eval()
            for i in range(self.frozen_stages - 1):
                self.layers[i].eval()
                for p in self.layers[i].parameters():
                    p.requires_grad = False

    def init_weights(self):
        """stf.py:570-582 (never called by the constructor, kept for API parity)."""
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.kaiming_normal_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)

    def _analysis(self, x):
        if self.training:
            t = self.patch_embed(x)
            Wh, Ww = t.shape[2], t.shape[3]
            t = t.flatten(2).transpose(1, 2).contiguous()
        else:
            t, Wh, Ww = self.patch_embed.tokens(x)
        for layer in self.layers:
            t, Wh, Ww = layer(t, Wh, Ww)
        C = self.embed_dim * 8
        return t.reshape(-1, Wh, Ww, C).permute(0, 3, 1, 2).contiguous()

    def _analysis_nhwc(self, x):
        """Inference: the token-major output of the last Swin stage IS the NHWC latent (no layout copy)."""
        t, Wh, Ww = self.patch_embed.tokens(x)
        for layer in self.layers:
            t, Wh, Ww = layer(t, Wh, Ww)
        return t.reshape(-1, Wh, Ww, self.embed_dim * 8)

    def _synthesis(self, y_hat):
        B, C, Wh, Ww = y_hat.shape
        t = y_hat.permute(0, 2, 3, 1).contiguous().reshape(B, Wh * Ww, C)
        for layer in self.syn_layers:
            t, Wh, Ww = layer(t, Wh, Ww)
        return self.end_conv(t.reshape(B, Wh, Ww, self.embed_dim).permute(0, 3, 1, 2).contiguous())

    def _synthesis_nhwc(self, y_hat):
        """Inference: NHWC end to end (stf.py:466-469 makes three full-size layout copies around the pixel shuffle).  The
        y_hat buffer is the token-major input of the first synthesis stage; end_conv[0] + PixelShuffle is one stf_conv2d
        launch on the token-major output of the last stage; the 3-channel end_conv[2] is a cuDNN call (x_hat only: nothing
        downstream of the bitstream)."""
        B, Wh, Ww, C = y_hat.shape
        t = y_hat.reshape(B, Wh * Ww, C)
        for layer in self.syn_layers:
            t, Wh, Ww = layer(t, Wh, Ww)
        E = self.embed_dim
        c0, ps, c2 = self.end_conv[0], self.end_conv[1], self.end_conv[2]
        packs = self.__dict__.setdefault("_conv_packs", _ConvPacks())
        u = ops.conv2d([t.reshape(B, Wh, Ww, E)], packs.get(c0, [E], ps.upscale_factor))     # (B, 2Wh, 2Ww, E) NHWC
        return c2(u.permute(0, 3, 1, 2)).contiguous()


class _NonNegativeParametrizer(nn.Module):
    """compressai/ops/parametrizers.py:23-49."""

    def __init__(self, minimum=0.0, reparam_offset=2 ** -18):
        super().__init__()
        self.minimum, self.reparam_offset = float(minimum), float(reparam_offset)
        self.register_buffer("pedestal", torch.Tensor([self.reparam_offset ** 2]))
        self.lower_bound = LowerBound((self.minimum + self.reparam_offset ** 2) ** 0.5)

    def init(self, x):
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))

    def forward(self, x):
        return self.lower_bound(x) ** 2 - self.pedestal


class GDN(nn.Module):
    """Generalised divisive normalisation (layers/gdn.py:26-75): a 1x1 convolution on x^2 (cuDNN) + rsqrt.
    Adjacent to the hot path (out of scope for custom kernels)."""

    def __init__(self, in_channels, inverse=False, beta_min=1e-6, gamma_init=0.1):
        super().__init__()
        self.inverse = bool(inverse)
        self.beta_reparam = _NonNegativeParametrizer(minimum=float(beta_min))
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(in_channels)))
        self.gamma_reparam = _NonNegativeParametrizer()
        self.gamma = nn.Parameter(self.gamma_reparam.init(float(gamma_init) * torch.eye(in_channels)))

    def forward(self, x):
        C = x.shape[1]
        norm = F.conv2d(x ** 2, self.gamma_reparam(self.gamma).reshape(C, C, 1, 1), self.beta_reparam(self.beta))
        return x * (torch.sqrt(norm) if self.inverse else torch.rsqrt(norm))


class WACNN(_SliceCodec):
    """CNN codec with window-attention blocks at 1/4 and 1/16 resolution (cnn.py:23-332)."""

    def __init__(self, N=192, M=320, **kwargs):
        super().__init__(**kwargs)
        self.num_slices = 10
        self.max_support_slices = 5
        self.g_a = nn.Sequential(
            conv(3, N), GDN(N), conv(N, N), GDN(N),
            Win_noShift_Attention(dim=N, num_heads=8, window_size=8, shift_size=4),
            conv(N, N), GDN(N), conv(N, M),
            Win_noShift_Attention(dim=M, num_heads=8, window_size=4, shift_size=2))
        self.g_s = nn.Sequential(
            Win_noShift_Attention(dim=M, num_heads=8, window_size=4, shift_size=2),
            deconv(M, N), GDN(N, inverse=True), deconv(N, N), GDN(N, inverse=True),
            Win_noShift_Attention(dim=N, num_heads=8, window_size=8, shift_size=4),
            deconv(N, N), GDN(N, inverse=True), deconv(N, 3))
        self.h_a = nn.Sequential(conv3x3(M, 320), nn.GELU(), conv3x3(320, 288), nn.GELU(),
                                 conv3x3(288, 256, stride=2), nn.GELU(), conv3x3(256, 224), nn.GELU(),
                                 conv3x3(224, 192, stride=2))

        def hyper_synthesis():
            return nn.Sequential(conv3x3(192, 192), nn.GELU(), subpel_conv3x3(192, 224, 2), nn.GELU(),
                                 conv3x3(224, 256), nn.GELU(), subpel_conv3x3(256, 288, 2), nn.GELU(),
                                 conv3x3(288, M))

        self.h_mean_s = hyper_synthesis()
        self.h_scale_s = hyper_synthesis()
        ms = self.max_support_slices
        self.cc_mean_transforms = nn.ModuleList(_stack5(M + 32 * min(i, ms)) for i in range(self.num_slices))
        self.cc_scale_transforms = nn.ModuleList(_stack5(M + 32 * min(i, ms)) for i in range(self.num_slices))
        self.lrp_transforms = nn.ModuleList(_stack5(M + 32 * min(i + 1, ms + 1)) for i in range(self.num_slices))
        self.entropy_bottleneck = EntropyBottleneck(N)
        self.gaussian_conditional = GaussianConditional(None)

    def _analysis(self, x):
        return self.g_a(x).contiguous()

    def _synthesis(self, y_hat):
        return self.g_s(y_hat)

    def _synthesis_nhwc(self, y_hat):
        return self.g_s(y_hat.permute(0, 3, 1, 2).contiguous())      # (g_s starts with a token-major attention block)


models = {"stf": SymmetricalTransFormer, "cnn": WACNN}   # compressai/zoo/__init__.py:20-27 (in-scope entries)
