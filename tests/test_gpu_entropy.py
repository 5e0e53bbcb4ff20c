"""GPU parity: the entropy-model kernels (called through the C ABI via stf_b200.ops / entropy_models)
against the CPU oracle and the golden vectors recorded from the reference.
Integer outputs (indexes, symbols) are bit-exact; fp32 outputs are checked with
|d| <= 1e-3 * |ref| + tiny absolute floor (BASELINE.json north_star tolerance), and in practice match
to a few ulp."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import entropy as OE

pytestmark = pytest.mark.gpu

RTOL = 1e-3   # north-star fp32 tolerance


@pytest.fixture(scope="module")
def ent(golden_dir):
    return np.load(os.path.join(golden_dir, "entropy_ops.npz"))


@pytest.fixture(scope="module")
def kat(golden_dir):
    return json.load(open(os.path.join(golden_dir, "kat.json")))


def cu(a):
    return torch.as_tensor(a).cuda()


def test_build_indexes_kat_and_golden(ent, kat):
    from stf_b200 import ops
    table = OE.scale_table()
    s = torch.tensor([float(v) for v in kat["build_indexes_in"]])
    assert ops.build_indexes(s.cuda(), table).cpu().tolist() == kat["build_indexes_out"]
    idx = ops.build_indexes(cu(ent["bi_scales"]), table)
    assert np.array_equal(idx.cpu().numpy(), ent["bi_indexes"])


@pytest.mark.parametrize("n", [1, 3, 4, 1023, 49152, 131072, 360448, 1 << 22])
def test_build_indexes_vs_oracle_sizes(n):
    """Natural slice sizes of the BASELINE configs, ragged sizes (scalar tail path) and a large one."""
    from stf_b200 import ops
    g = torch.Generator().manual_seed(n)
    table = OE.scale_table()
    s = torch.exp(torch.rand(n, generator=g) * (np.log(400) - np.log(0.01)) + np.log(0.01))
    k = min(n, 64)
    s[:k] = table[:k]                                       # exact table hits (<= comparison edge)
    if n > 200:
        s[64:128] = torch.nextafter(table, torch.tensor(float("inf")))
        s[128:192] = torch.nextafter(table, torch.tensor(0.0))
        s[192:196] = torch.tensor([float("nan"), float("inf"), -float("inf"), 0.0])
    ref = OE.build_indexes(s)
    out = ops.build_indexes(s.cuda(), table)
    assert out.dtype == torch.int32 and torch.equal(out.cpu(), ref)


def test_build_indexes_non_monotone_table():
    from stf_b200 import ops
    table = torch.tensor([0.5, 0.2, 0.9, 0.3, 2.0])        # reference semantics = linear count
    s = torch.rand(5000) * 3
    ref = OE.build_indexes(s, table, scale_bound=0.11)
    assert torch.equal(ops.build_indexes(s.cuda(), table, 0.11).cpu(), ref)


def test_quantize_kats(ent, kat):
    from stf_b200.entropy_models import GaussianConditional
    gc = GaussianConditional(None).cuda()
    assert gc.quantize(cu(kat["quantize_in"]).float(), "symbols").cpu().tolist() == kat["quantize_out"]
    y, mu = cu(ent["gc_y"]), cu(ent["gc_mu"])
    q = gc.quantize(y, "symbols", mu)
    assert q.dtype == torch.int32 and np.array_equal(q.cpu().numpy(), ent["gc_symbols"])
    assert np.array_equal(gc.dequantize(q, mu).cpu().numpy(), ent["gc_dequant"])
    assert np.array_equal(gc.quantize(y, "dequantize", mu).cpu().numpy(), ent["gc_out"])
    with pytest.raises(ValueError):
        gc.quantize(y, "bogus")


def test_gaussian_conditional_forward_golden(ent, kat):
    from stf_b200.entropy_models import GaussianConditional
    gc = GaussianConditional(None).cuda().eval()
    f = kat["gc_forward"]
    out, lik = gc(cu(f["y"]).float(), cu(f["scale"]).float(), cu(f["mu"]).float())
    assert out.cpu().tolist() == f["out"]
    assert np.allclose(lik.cpu().numpy(), np.array(f["lik"], dtype=np.float32), rtol=RTOL, atol=1e-12)
    out, lik = gc(cu(ent["gc_y"]), cu(ent["gc_scale"]), cu(ent["gc_mu"]))
    assert np.array_equal(out.cpu().numpy(), ent["gc_out"])
    ref = ent["gc_lik"]
    err = np.abs(lik.cpu().numpy() - ref)
    assert np.all(err <= RTOL * ref + 1e-9)
    assert np.median(err / ref) < 1e-6          # in practice: ulp-level agreement


@pytest.mark.parametrize("shape", [(1, 32, 32, 48), (16, 32, 16, 16), (2, 32, 5, 7), (1, 32, 88, 128)])
def test_compress_step_and_likelihood_slices(shape):
    """Slice addressing inside a larger latent + the fused compress step, vs the oracle."""
    from stf_b200 import ops
    B, Cs, h, w = shape
    M = 3 * Cs
    g = torch.Generator().manual_seed(h * w)
    y = 6 * torch.randn(B, M, h, w, generator=g)
    mu = 2 * torch.randn(B, Cs, h, w, generator=g)
    sc = torch.exp(torch.rand(B, Cs, h, w, generator=g) * 9 - 4.5)
    y[0, Cs:2 * Cs].view(-1)[:16] = (mu[0].view(-1)[:16] + torch.arange(-8, 8) + 0.5)   # ties
    table = OE.scale_table()
    total = M * h * w
    sym = torch.full((B, total), -7, dtype=torch.int32).cuda()
    idx = torch.full((B, total), -7, dtype=torch.int32).cuda()
    off = Cs * h * w
    yh = ops.gaussian_compress_step(y.cuda(), Cs, sc.cuda(), mu.cuda(), table, sym, idx, off)
    ys = y[:, Cs:2 * Cs]
    q = OE.quantize(ys, "symbols", mu)
    assert torch.equal(sym[:, off:2 * off].cpu().reshape(B, Cs, h, w), q)
    assert torch.equal(idx[:, off:2 * off].cpu().reshape(B, Cs, h, w), OE.build_indexes(sc))
    assert torch.equal(yh.cpu(), OE.dequantize(q, mu))
    assert int((sym[:, :off] != -7).sum()) == 0 and int((sym[:, 2 * off:] != -7).sum()) == 0   # no stray writes
    # decode side: dequantize from the strided symbol buffer
    assert torch.equal(ops.dequantize(sym, off, mu.cuda()).cpu(), OE.dequantize(q, mu))
    # likelihood + ste_round value
    yhat, lik = ops.gaussian_likelihood(y.cuda(), Cs, sc.cuda(), mu.cuda(), ste_round=True)
    o_out, o_lik = OE.gaussian_conditional_eval(ys, sc, mu)
    assert torch.equal(yhat.cpu(), OE.ste_round_value(ys - mu) + mu)
    err = (lik.cpu() - o_lik).abs()
    assert bool((err <= RTOL * o_lik + 1e-9).all())
    yq, _ = ops.gaussian_likelihood(y.cuda(), Cs, sc.cuda(), mu.cuda(), ste_round=False)
    assert torch.equal(yq.cpu(), o_out)


def test_empty_inputs():
    from stf_b200 import ops
    table = OE.scale_table()
    e = torch.empty(0, device="cuda")
    assert ops.build_indexes(e, table).numel() == 0
    assert ops.quantize_symbols(e).numel() == 0
    with pytest.raises(RuntimeError):
        ops.build_indexes(torch.ones(4), table)              # CPU tensor: no fallback


def _eb(seed, golden=True):
    from stf_b200.entropy_models import EntropyBottleneck
    from stf_b200.synth import synthetic_state_dict
    eb = EntropyBottleneck(192)
    spec = {"entropy_bottleneck." + k: (tuple(v.shape), v.dtype) for k, v in eb.state_dict().items()}
    sd = synthetic_state_dict(spec, seed)
    eb.load_state_dict({k[len("entropy_bottleneck."):]: v for k, v in sd.items()}, strict=False)
    return eb.cuda().eval()


def test_entropy_bottleneck_golden(ent):
    eb = _eb(int(ent["eb_seed"]))
    assert eb.update(force=True)
    assert np.array_equal(eb.quantized_cdf.cpu().numpy(), ent["eb_cdf"])
    z = cu(ent["eb_z"])
    out, lik = eb(z)
    assert np.allclose(out.cpu().numpy(), ent["eb_out"], rtol=1e-6, atol=1e-6)
    ref = ent["eb_lik"]
    assert np.all(np.abs(lik.cpu().numpy() - ref) <= RTOL * ref + 1e-9)
    strings = eb.compress(z)
    assert [s.hex() for s in strings] == [str(s) for s in ent["eb_strings_hex"]]      # bit-exact bitstream
    assert np.array_equal(eb.decompress(strings, z.shape[-2:]).cpu().numpy(), ent["eb_zhat"])
    with pytest.raises(ValueError):
        _eb(3).compress(z)                                                             # update() not run


@pytest.mark.parametrize("shape", [(16, 192, 4, 4), (1, 192, 8, 12), (1, 192, 22, 32)])
def test_entropy_bottleneck_vs_oracle(shape):
    eb = _eb(5)
    p = {k: v.detach().cpu() for k, v in eb.state_dict().items() if k in OE.eb_param_names()}
    g = torch.Generator().manual_seed(shape[2])
    z = 5 * torch.randn(*shape, generator=g)
    out, lik = eb(z.cuda())
    o_out, o_lik = OE.eb_forward_eval(p, z)
    assert torch.allclose(out.cpu(), o_out, rtol=1e-6, atol=1e-6)
    assert bool(((lik.cpu() - o_lik).abs() <= RTOL * o_lik + 1e-9).all())


def test_gaussian_conditional_compress_decompress_api():
    """EntropyModel.compress / decompress (entropy_models.py:203-290) on CUDA tensors: strings equal the oracle
    coder's for the same symbols / indexes, and decompress inverts compress."""
    from stf_b200.entropy_models import GaussianConditional
    gc = GaussianConditional(None).cuda().eval()
    gc.update_scale_table(OE.scale_table())
    g = torch.Generator().manual_seed(21)
    B = 3
    sc = torch.exp(torch.rand(B, 8, 5, 7, generator=g) * 7 - 3)
    mu = torch.randn(B, 8, 5, 7, generator=g)
    y = mu + sc * torch.randn(B, 8, 5, 7, generator=g) * 1.5
    idx = gc.build_indexes(sc.cuda())
    strings = gc.compress(y.cuda(), idx, mu.cuda())
    assert len(strings) == B
    cdf, lens, offs = OE.gaussian_tables()
    o_idx = OE.build_indexes(sc)
    o_sym = OE.quantize(y, "symbols", mu)
    for b in range(B):
        assert strings[b] == OE.rans_encode(o_sym[b].numpy(), o_idx[b].numpy(), cdf, lens, offs)
    y_hat = gc.decompress(strings, idx, mu.cuda())
    assert torch.equal(y_hat.cpu(), OE.dequantize(o_sym, mu))
    with pytest.raises(ValueError):
        gc.compress(y.cuda(), idx[:, :4], mu.cuda())
    with pytest.raises(ValueError):
        gc.decompress(strings[:2], idx, mu.cuda())
    with pytest.raises(ValueError):
        gc.decompress("notalist", idx)


def test_build_indexes_lut_equals_count_on_random_bit_patterns():
    """The bucket-LUT index (csrc/entropy_kernels.cu scale_index_lut) against the definition
    idx = #{i < 63 : table[i] < max(scale, bound)} (NaN -> 63) on uniformly random float BIT PATTERNS
    (negatives, denormals, infinities, NaNs, every exponent) plus every table value and its +-1-ulp neighbours."""
    from stf_b200 import ops
    table = OE.scale_table()
    g = torch.Generator().manual_seed(7)
    bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (1 << 20,), generator=g, dtype=torch.int64).to(torch.int32)
    t = torch.as_tensor(table, dtype=torch.float32)
    ti = t.view(torch.int32)
    near = torch.cat([ti - 1, ti, ti + 1, ti - 2, ti + 2])
    s = torch.cat([bits, near]).view(torch.float32)
    sig = torch.where(s < 0.11, torch.full_like(s, 0.11), s)          # torch.max(x, bound) with NaN propagation
    sig = torch.where(torch.isnan(s), s, sig)
    ref = (t[:-1][None, :] < sig[:, None]).sum(1).to(torch.int32)
    ref = torch.where(torch.isnan(sig), torch.full_like(ref, 63), ref)
    got = ops.build_indexes(s.cuda(), table).cpu()
    assert torch.equal(got, ref)
    # a table whose thresholds crowd one bucket falls back to the binary search: same definition
    crowded = torch.cat([torch.linspace(0.11, 0.1101, 40), torch.linspace(0.2, 256.0, 24)]).float()
    sig2 = torch.where(s < 0.11, torch.full_like(s, 0.11), s)
    sig2 = torch.where(torch.isnan(s), s, sig2)
    ref2 = (crowded[:-1][None, :] < sig2[:, None]).sum(1).to(torch.int32)
    ref2 = torch.where(torch.isnan(sig2), torch.full_like(ref2, 63), ref2)
    assert torch.equal(ops.build_indexes(s.cuda(), crowded.numpy()).cpu(), ref2)


def test_conv_glue_kernels_bias_gelu_and_layernorm():
    """stf_bias_act == torch's `conv_out + bias` followed by nn.GELU() on channels_last data (same fp32 add, same GELU
    formula: bit-identical or 1 ulp), stf_layernorm_fwd == F.layer_norm to fp32 round-off."""
    from stf_b200 import ops
    g = torch.Generator().manual_seed(2)
    for (B, C, H, W) in ((3, 224, 8, 12), (2, 32, 5, 7), (1, 48, 16, 24)):
        x = (3 * torch.randn(B, C, H, W, generator=g)).cuda().contiguous(memory_format=torch.channels_last)
        b = torch.randn(C, generator=g).cuda()
        ref = torch.nn.functional.gelu(x + b.reshape(1, -1, 1, 1))
        got = ops.bias_act_(x.clone(memory_format=torch.channels_last), b, gelu=True)
        assert float((got - ref).abs().max()) <= 2e-7 * float(ref.abs().max())
        got2 = ops.bias_act_(x.clone(memory_format=torch.channels_last), b, gelu=False)
        assert torch.equal(got2, x + b.reshape(1, -1, 1, 1))
    with pytest.raises(ValueError):
        ops.bias_act_(torch.randn(2, 8, 4, 4, device="cuda"), torch.zeros(8, device="cuda"), True)   # NCHW-contiguous input
    for (M, C) in ((1000, 48), (77, 384), (5, 768)):
        x = (2 * torch.randn(M, C, generator=g) + 1).cuda()
        w, b = (1 + 0.1 * torch.randn(C, generator=g)).cuda(), (0.1 * torch.randn(C, generator=g)).cuda()
        ref = torch.nn.functional.layer_norm(x, (C,), w, b, 1e-5)
        assert float((ops.layernorm(x, w, b, 1e-5) - ref).abs().max()) <= 2e-6 * float(ref.abs().max())


def test_device_rans_decoder_matches_host_decoder():
    """stf_rans_decode_device (one warp lane per stream, state carried across calls) returns exactly the symbols that were
    encoded -- i.e. what the host decoder / the reference's RansDecoder return -- incl. escape-coded outliers, ragged stream
    lengths, more than 32 streams (two CTAs) and a truncated stream (status STF_E_STREAM, other streams unaffected)."""
    from stf_b200 import ans
    cdf, lens, offs = OE.gaussian_tables()
    tab = ans.RansTable(cdf, lens, offs)
    table = OE.scale_table().numpy()
    rng = np.random.default_rng(11)
    B, n, slices = 37, 6000, 3
    ix = rng.integers(0, 64, size=(B, n * slices)).astype(np.int32)
    sy = np.rint(rng.standard_normal((B, n * slices)) * table[ix] * 1.5).astype(np.int32)
    sy[:, ::97] += 5000                                                   # escapes (nibble-coded bypass values)
    strings = ans.encode_rows(tab, sy, ix)
    strings[5] = strings[5][: len(strings[5]) // 2 // 4 * 4]              # truncated stream
    ds = ans.DeviceStreams(B, sum(len(s) for s in strings) // 4 + 4 * B, "cuda")
    assert ds.fits(strings)
    ds.load(strings)
    ds.upload()
    idx_d = torch.from_numpy(ix).cuda()
    out = torch.empty((B, n * slices), dtype=torch.int32, device="cuda")
    for k in range(slices):
        ans.decode_device(tab, ds, idx_d[:, k * n:(k + 1) * n], out[:, k * n:(k + 1) * n], first=(k == 0))
    got, status = out.cpu().numpy(), ds.status.cpu().numpy()
    ok = [b for b in range(B) if b != 5]
    assert np.array_equal(got[ok], sy[ok])
    assert (status[ok] == 0).all() and status[5] == -6


def test_slice_step_narrow_outputs_equal_wide_and_flag_overflow():
    """stf_slice_step_nhwc with int16 symbols / uint8 indexes (the 3-byte-per-symbol transfer format): same values as the
    int32 outputs, at a non-zero output offset; a symbol beyond int16 sets the overflow flag (and only then)."""
    from stf_b200 import ops
    from stf_b200.models import get_scale_table
    torch.manual_seed(3)
    B, h, w, C = 3, 9, 7, 32                       # ragged: plane 63 is not a multiple of the 32-pixel tile
    y = (torch.randn(B, h, w, C, device="cuda") * 40).contiguous()
    mu = torch.randn(B, h, w, C, device="cuda")
    sc = (torch.rand(B, h, w, C, device="cuda") * 60 + 0.05).contiguous()
    table = get_scale_table()
    total, off = 2 * C * h * w, C * h * w
    sym32 = torch.zeros((B, total), dtype=torch.int32, device="cuda")
    idx32 = torch.zeros_like(sym32)
    ops.slice_step_nhwc(y=y, scales=sc, means=mu, symbols_out=sym32, indexes_out=idx32, out_offset=off, table=table)
    sym16 = torch.zeros((B, total), dtype=torch.int16, device="cuda")
    idx8 = torch.zeros((B, total), dtype=torch.uint8, device="cuda")
    ovf = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.slice_step_nhwc(y=y, scales=sc, means=mu, symbols_out=sym16, indexes_out=idx8, out_offset=off, table=table, overflow=ovf)
    assert int(ovf.item()) == 0
    assert torch.equal(sym16.int(), sym32) and torch.equal(idx8.int(), idx32)
    assert int(sym32[:, off:].abs().max()) > 50 and int(idx32.max()) > 40 and not sym32[:, :off].any()
    # indexes only (the decoder's step)
    idx8b = torch.zeros((B, C * h * w), dtype=torch.uint8, device="cuda")
    ops.slice_step_nhwc(scales=sc, indexes_out=idx8b, table=table, overflow=ovf)
    assert int(ovf.item()) == 0 and torch.equal(idx8b.int(), idx32[:, off:])
    # overflow: one latent beyond int16
    y2 = y.clone()
    y2[1, 4, 3, 17] = 40000.0
    ops.slice_step_nhwc(y=y2, scales=sc, means=mu, symbols_out=sym16, indexes_out=idx8, out_offset=off, table=table, overflow=ovf)
    assert int(ovf.item()) == 1
    with pytest.raises(ValueError):
        ops.slice_step_nhwc(y=y, scales=sc, means=mu, symbols_out=sym16, indexes_out=idx8, out_offset=off, table=table)
