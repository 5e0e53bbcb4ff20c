"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's two codecs, end to end.

`StfOracle` / `WacnnOracle` hold a reference-format state_dict and restate
forward / compress / decompress of

  SymmetricalTransFormer   compressai/models/stf.py:584-648, 674-735, 737-788
  WACNN                    compressai/models/cnn.py:141-189, 210-267, 289-332

with plain torch CPU ops in the reference's order (same .tolist() hand-off to the entropy coder,
same slice-major symbol order).  The rANS back end is selectable: "oracle" = rans_oracle.c,
"ref" = the reference's own compiled extension from oracle/_ref (when present).
Used as the parity checker and as bench.py's cpu_baseline / `--impl reference` arm.
"""
import torch
import torch.nn.functional as F

from . import entropy as E
from . import swin as S


def _conv(sd, name, x, stride=1, padding=None):
    w = sd[name + ".weight"]
    pad = w.shape[-1] // 2 if padding is None else padding
    return F.conv2d(x, w, sd[name + ".bias"], stride=stride, padding=pad)


def _conv_stack(sd, pfx, x, strides=(1, 1, 1, 1, 1)):
    """nn.Sequential(conv, GELU, conv, GELU, ..., conv) with module indices 0,2,4,6,8."""
    for j, st in enumerate(strides):
        x = _conv(sd, f"{pfx}{2 * j}", x, stride=st)
        if j + 1 < len(strides):
            x = F.gelu(x)
    return x


def _hyper_synthesis(sd, pfx, x):
    """h_mean_s / h_scale_s: conv3x3, GELU, subpel(2), GELU, conv3x3, GELU, subpel(2), GELU, conv3x3
    (stf.py:488-509, cnn.py:66-88); subpel = conv3x3 + PixelShuffle (layers/layers.py:34-38)."""
    x = F.gelu(_conv(sd, pfx + "0", x))
    x = F.gelu(F.pixel_shuffle(_conv(sd, pfx + "2.0", x), 2))
    x = F.gelu(_conv(sd, pfx + "4", x))
    x = F.gelu(F.pixel_shuffle(_conv(sd, pfx + "6.0", x), 2))
    return _conv(sd, pfx + "8", x)


def train_noise(seed, B, M, h, w, Cz, hz, wz, num_slices=12):
    """The reference's quantisation-noise stream under torch.manual_seed(seed) in train() mode: entropy_models.py:131-135
    draws torch.empty_like(inputs).uniform_(-1/2, 1/2) first for z in the (C, 1, B*h*w) layout of :455-461, then for each
    y slice (B, M/num_slices, h, w) in order (stf.py:613-623).  Returned in model layout: {"z": (B,Cz,hz,wz), "y": (B,M,h,w)}."""
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    nz = torch.empty(Cz, 1, B * hz * wz).uniform_(-0.5, 0.5)
    ny = [torch.empty(B, M // num_slices, h, w).uniform_(-0.5, 0.5) for _ in range(num_slices)]
    torch.random.set_rng_state(state)
    return {"z": nz.reshape(Cz, B, hz, wz).permute(1, 0, 2, 3).contiguous(), "y": torch.cat(ny, dim=1)}


def _eb_params(sd, pfx="entropy_bottleneck."):
    return {k: sd[pfx + k] for k in E.eb_param_names()}


class _RansBackend:
    def __init__(self, kind):
        self.kind = kind
        if kind == "ref":
            from .ref_import import load_ref_ans
            self.ans = load_ref_ans()

    def encode(self, symbols, indexes, cdf, lens, offs):
        if self.kind == "ref":   # same Python-list hand-off as stf.py:721-722,730
            return self.ans.RansEncoder().encode_with_indexes(
                symbols.reshape(-1).tolist(), indexes.reshape(-1).tolist(), cdf.tolist(), lens.tolist(), offs.tolist())
        return E.rans_encode(symbols.reshape(-1).numpy(), indexes.reshape(-1).numpy(), cdf, lens, offs)

    def decoder(self, stream, cdf, lens, offs):
        if self.kind == "ref":
            dec = self.ans.RansDecoder()
            dec.set_stream(stream)
            cl, ll, ol = cdf.tolist(), lens.tolist(), offs.tolist()
            return lambda idx: torch.tensor(dec.decode_stream(idx.reshape(-1).tolist(), cl, ll, ol), dtype=torch.int32)
        dec = E.RansStreamDecoder(stream)
        return lambda idx: torch.from_numpy(dec.decode(idx.reshape(-1).numpy(), cdf, lens, offs))


class _SliceCodec:
    """Shared hyperprior + channel-conditional slice loop (stf.py:600-636 == cnn.py:144-183)."""

    num_slices = 0
    max_support = 0

    def __init__(self, state_dict, rans="oracle"):
        self.sd = {k: v.detach().float() if v.is_floating_point() else v.detach() for k, v in state_dict.items()}
        self.eb = _eb_params(self.sd)
        self.rans = _RansBackend(rans)
        self.update()

    def update(self):
        """stf.py:650-655 -> entropy_models.py:588-624 and :354-393."""
        self.table = E.scale_table()
        self.gc_cdf, self.gc_len, self.gc_off = E.gaussian_tables(self.table)
        self.eb_cdf, self.eb_len, self.eb_off = E.eb_tables(self.eb)

    # -- transforms supplied by subclasses
    def analysis(self, x):
        raise NotImplementedError

    def synthesis(self, y_hat):
        raise NotImplementedError

    def _h_a(self, y):
        return _conv_stack(self.sd, "h_a.", y, strides=(1, 1, 2, 1, 2))

    def _slice_params(self, i, means, scales, prev, hw):
        sup = prev[: self.max_support]
        mean_sup = torch.cat([means] + sup, dim=1)
        mu = _conv_stack(self.sd, f"cc_mean_transforms.{i}.", mean_sup)[:, :, : hw[0], : hw[1]]
        scale_sup = torch.cat([scales] + sup, dim=1)
        sc = _conv_stack(self.sd, f"cc_scale_transforms.{i}.", scale_sup)[:, :, : hw[0], : hw[1]]
        return mean_sup, mu, sc

    def _lrp(self, i, mean_sup, y_hat_slice):
        lrp = _conv_stack(self.sd, f"lrp_transforms.{i}.", torch.cat([mean_sup, y_hat_slice], dim=1))
        return y_hat_slice + 0.5 * torch.tanh(lrp)

    @torch.no_grad()
    def forward(self, x):
        y = self.analysis(x)
        hw = y.shape[2:]
        z = self._h_a(y)
        _, z_lik = E.eb_forward_eval(self.eb, z)
        med = E.eb_medians(self.eb).reshape(1, -1, 1, 1)
        z_hat = E.ste_round_value(z - med) + med
        scales = _hyper_synthesis(self.sd, "h_scale_s.", z_hat)
        means = _hyper_synthesis(self.sd, "h_mean_s.", z_hat)
        y_hat_slices, liks = [], []
        for i, y_i in enumerate(y.chunk(self.num_slices, 1)):
            mean_sup, mu, sc = self._slice_params(i, means, scales, y_hat_slices, hw)
            _, lik = E.gaussian_conditional_eval(y_i, sc, mu)
            liks.append(lik)
            y_hat_i = E.ste_round_value(y_i - mu) + mu
            y_hat_slices.append(self._lrp(i, mean_sup, y_hat_i))
        y_hat = torch.cat(y_hat_slices, dim=1)
        return {"x_hat": self.synthesis(y_hat), "likelihoods": {"y": torch.cat(liks, dim=1), "z": z_lik},
                "y": y, "y_hat": y_hat}

    def forward_train(self, x, noise):
        """stf.py:584-648 with self.training == True and drop_path_rate = 0: "noise" quantisation for both
        likelihoods (injected tensors noise["y"], noise["z"] stand for U(-1/2, 1/2)), ste_round for
        z_hat / y_hat.  Differentiable: call with requires_grad tensors in self.sd for the gradient oracle."""
        y = self.analysis(x)
        hw = y.shape[2:]
        z = self._h_a(y)
        _, z_lik = E.eb_forward_train(self.eb, z, noise["z"])
        med = E.eb_medians(self.eb).reshape(1, -1, 1, 1)
        z_hat = E.ste_round(z - med) + med
        scales = _hyper_synthesis(self.sd, "h_scale_s.", z_hat)
        means = _hyper_synthesis(self.sd, "h_mean_s.", z_hat)
        y_hat_slices, liks = [], []
        for i, y_i in enumerate(y.chunk(self.num_slices, 1)):
            mean_sup, mu, sc = self._slice_params(i, means, scales, y_hat_slices, hw)
            _, lik = E.gaussian_conditional_train(y_i, sc, mu, noise["y"].chunk(self.num_slices, 1)[i])
            liks.append(lik)
            y_hat_i = E.ste_round(y_i - mu) + mu
            y_hat_slices.append(self._lrp(i, mean_sup, y_hat_i))
        y_hat = torch.cat(y_hat_slices, dim=1)
        return {"x_hat": self.synthesis(y_hat), "likelihoods": {"y": torch.cat(liks, dim=1), "z": z_lik}}

    @torch.no_grad()
    def compress(self, x, debug=None):
        y = self.analysis(x)
        hw = y.shape[2:]
        z = self._h_a(y)
        med = E.eb_medians(self.eb).reshape(1, -1, 1, 1)
        # EntropyBottleneck.compress, entropy_models.py:508-515 -> :203-238 (one string per image)
        z_sym = E.quantize(z, "symbols", med)
        z_idx = E.eb_indexes(z.shape)
        z_strings = [self.rans.encode(z_sym[b], z_idx[b], self.eb_cdf, self.eb_len, self.eb_off) for b in range(z.shape[0])]
        z_hat = E.dequantize(z_sym, med)      # what decompress(z_strings) returns, :517-522
        scales = _hyper_synthesis(self.sd, "h_scale_s.", z_hat)
        means = _hyper_synthesis(self.sd, "h_mean_s.", z_hat)
        y_hat_slices, syms, idxs = [], [], []
        for i, y_i in enumerate(y.chunk(self.num_slices, 1)):
            mean_sup, mu, sc = self._slice_params(i, means, scales, y_hat_slices, hw)
            idx = E.build_indexes(sc, self.table)
            q = E.quantize(y_i, "symbols", mu)
            syms.append(q.reshape(-1))
            idxs.append(idx.reshape(-1))
            y_hat_slices.append(self._lrp(i, mean_sup, q + mu))
        syms, idxs = torch.cat(syms), torch.cat(idxs)
        if debug is not None:
            debug.update(y=y, z=z, symbols=syms, indexes=idxs, z_symbols=z_sym, y_hat=torch.cat(y_hat_slices, 1))
        y_string = self.rans.encode(syms, idxs, self.gc_cdf, self.gc_len, self.gc_off)
        return {"strings": [[y_string], z_strings], "shape": z.shape[-2:]}

    @torch.no_grad()
    def decompress(self, strings, shape):
        med = E.eb_medians(self.eb).reshape(1, -1, 1, 1)
        C = self.eb_cdf.shape[0]
        z_idx = E.eb_indexes((len(strings[1]), C, *shape))
        z_sym = torch.stack([self.rans.decoder(s, self.eb_cdf, self.eb_len, self.eb_off)(z_idx[b]).reshape(C, *shape)
                             for b, s in enumerate(strings[1])])
        z_hat = E.dequantize(z_sym, med)
        scales = _hyper_synthesis(self.sd, "h_scale_s.", z_hat)
        means = _hyper_synthesis(self.sd, "h_mean_s.", z_hat)
        hw = (z_hat.shape[2] * 4, z_hat.shape[3] * 4)
        dec = self.rans.decoder(strings[0][0], self.gc_cdf, self.gc_len, self.gc_off)
        y_hat_slices = []
        for i in range(self.num_slices):
            mean_sup, mu, sc = self._slice_params(i, means, scales, y_hat_slices, hw)
            idx = E.build_indexes(sc, self.table)
            rv = dec(idx).float().reshape(1, -1, hw[0], hw[1])       # batch 1 hard-coded, stf.py:770
            y_hat_slices.append(self._lrp(i, mean_sup, E.dequantize(rv, mu)))
        y_hat = torch.cat(y_hat_slices, dim=1)
        return {"x_hat": self.synthesis(y_hat).clamp_(0, 1)}


class StfOracle(_SliceCodec):
    """SymmetricalTransFormer with constructor defaults (stf.py:385-404)."""

    num_slices, max_support = 12, 6
    embed_dim, depths, heads, ws = 48, (2, 2, 6, 2), (3, 6, 12, 24), 4

    def analysis(self, x):
        sd = self.sd
        # PatchEmbed, stf.py:365-381 (pad to patch multiple, conv k2 s2, LN over channels)
        if x.shape[3] % 2:
            x = F.pad(x, (0, 1))
        if x.shape[2] % 2:
            x = F.pad(x, (0, 0, 0, 1))
        t = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=2)
        H, W = t.shape[2], t.shape[3]
        t = S.layer_norm(sd, "patch_embed.norm.", t.flatten(2).transpose(1, 2))
        for i in range(4):
            t, H, W = S.basic_layer(sd, f"layers.{i}.", t, H, W, self.depths[i], self.heads[i], self.ws,
                                    "merge" if i < 3 else None)
        C = self.embed_dim * 8
        return t.reshape(-1, H, W, C).permute(0, 3, 1, 2).contiguous()

    def synthesis(self, y_hat):
        sd = self.sd
        B, C, H, W = y_hat.shape
        t = y_hat.permute(0, 2, 3, 1).reshape(B, H * W, C)
        depths, heads = self.depths[::-1], self.heads[::-1]
        for i in range(4):
            t, H, W = S.basic_layer(sd, f"syn_layers.{i}.", t, H, W, depths[i], heads[i], self.ws,
                                    "split" if i < 3 else None)
        t = t.reshape(B, H, W, self.embed_dim).permute(0, 3, 1, 2).contiguous()
        t = F.pixel_shuffle(_conv(sd, "end_conv.0", t), 2)     # stf.py:466-469
        return _conv(sd, "end_conv.2", t)


class WacnnOracle(_SliceCodec):
    """WACNN(N=192, M=320), cnn.py:26-52."""

    num_slices, max_support = 10, 5

    def _gdn(self, pfx, x, inverse):
        """layers/gdn.py:62-75 with NonNegativeParametrizer, ops/parametrizers.py:23-49."""
        sd = self.sd

        def reparam(v, minimum):
            ped = (2 ** -18) ** 2
            bound = (minimum + ped) ** 0.5
            return E.lower_bound(v, bound) ** 2 - ped
        beta = reparam(sd[pfx + "beta"], 1e-6)
        gamma = reparam(sd[pfx + "gamma"], 0.0)
        C = x.shape[1]
        norm = F.conv2d(x ** 2, gamma.reshape(C, C, 1, 1), beta)
        norm = torch.sqrt(norm) if inverse else torch.rsqrt(norm)
        return x * norm

    def _res_unit(self, pfx, x):
        """layers/layers.py:52-71."""
        h = F.gelu(_conv(self.sd, pfx + "conv.0", x))
        h = F.gelu(_conv(self.sd, pfx + "conv.2", h))
        h = _conv(self.sd, pfx + "conv.4", h)
        return F.gelu(h + x)

    def _win_attention_block(self, pfx, x, ws, shift):
        """Win_noShift_Attention.forward, layers/layers.py:83-89."""
        a = x
        for j in range(3):
            a = self._res_unit(f"{pfx}conv_a.{j}.", a)
        b = S.win_based_attention(self.sd, pfx + "conv_b.0.", x, 8, ws, shift)
        for j in (1, 2, 3):
            b = self._res_unit(f"{pfx}conv_b.{j}.", b)
        b = _conv(self.sd, pfx + "conv_b.4", b)
        return a * torch.sigmoid(b) + x

    def analysis(self, x):
        sd = self.sd
        x = self._gdn("g_a.1.", _conv(sd, "g_a.0", x, stride=2), False)
        x = self._gdn("g_a.3.", _conv(sd, "g_a.2", x, stride=2), False)
        x = self._win_attention_block("g_a.4.", x, 8, 4)
        x = self._gdn("g_a.6.", _conv(sd, "g_a.5", x, stride=2), False)
        x = _conv(sd, "g_a.7", x, stride=2)
        return self._win_attention_block("g_a.8.", x, 4, 2)

    def _deconv(self, name, x):
        """models/utils.py:124-132."""
        w = self.sd[name + ".weight"]
        return F.conv_transpose2d(x, w, self.sd[name + ".bias"], stride=2, padding=w.shape[-1] // 2, output_padding=1)

    def synthesis(self, y_hat):
        x = self._win_attention_block("g_s.0.", y_hat, 4, 2)
        x = self._gdn("g_s.2.", self._deconv("g_s.1", x), True)
        x = self._gdn("g_s.4.", self._deconv("g_s.3", x), True)
        x = self._win_attention_block("g_s.5.", x, 8, 4)
        x = self._gdn("g_s.7.", self._deconv("g_s.6", x), True)
        return self._deconv("g_s.8", x)
