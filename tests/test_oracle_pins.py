"""CPU: the oracle (oracle/) against the golden vectors recorded from the live reference
(tests/golden/, written by oracle/gen_golden.py) and, when /root/reference is present, against the
live reference itself.  These pins are what lets the GPU parity tests trust the oracle."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import codec as OC
from oracle import entropy as OE
from oracle import swin as OS
from oracle.ref_import import reference_available
from stf_b200.synth import synthetic_image, synthetic_state_dict


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def kat(golden_dir):
    return json.load(open(os.path.join(golden_dir, "kat.json")))


@pytest.fixture(scope="module")
def ent(golden_dir):
    return np.load(os.path.join(golden_dir, "entropy_ops.npz"))


@pytest.fixture(scope="module")
def sw(golden_dir):
    return np.load(os.path.join(golden_dir, "swin_ops.npz"))


def test_scale_table_and_gaussian_cdf(kat):
    t = OE.scale_table()
    assert sha(t.numpy()) == kat["scale_table_sha256"]
    assert [float(v) for v in t[:4]] + [float(t[63])] == kat["scale_table_head"]
    cdf, lens, offs = OE.gaussian_tables()
    assert list(cdf.shape) == kat["gc_cdf_shape"]
    assert sha(cdf.astype(np.int32)) == kat["gc_cdf_sha256"]
    assert lens.tolist() == kat["gc_cdf_length"] and offs.tolist() == kat["gc_offset"]


def test_build_indexes_quantize_forward_kats(kat):
    s = torch.tensor([float(v) for v in kat["build_indexes_in"]])
    assert OE.build_indexes(s).tolist() == kat["build_indexes_out"]
    assert OE.quantize(torch.tensor(kat["quantize_in"]), "symbols").tolist() == kat["quantize_out"]
    f = kat["gc_forward"]
    out, lik = OE.gaussian_conditional_eval(torch.tensor(f["y"]), torch.tensor(f["scale"]), torch.tensor(f["mu"]))
    assert out.tolist() == f["out"]
    assert [float(v) for v in lik] == f["lik"]


def test_rans_oracle_kat_and_streams(kat):
    cdf, lens, offs = OE.gaussian_tables()
    r = kat["rans"]
    assert OE.rans_encode(r["symbols"], r["indexes"], cdf, lens, offs).hex() == r["bytes_hex"]
    assert OE.rans_decode(bytes.fromhex(r["bytes_hex"]), r["indexes"], cdf, lens, offs).tolist() == r["symbols"]
    rng = np.random.default_rng(kat["rans_streams_seed"])
    table = OE.scale_table().numpy()
    for rec in kat["rans_streams"]:
        n, spread = rec["n"], rec["spread"]
        ix = rng.integers(0, 64, size=n).astype(np.int32)
        sy = np.rint(rng.standard_normal(n) * table[ix] * spread).astype(np.int32)
        if n < 4:
            sy = np.array([70000], dtype=np.int32)[:n]
        assert sha(sy) == rec["symbols_sha256"] and sha(ix) == rec["indexes_sha256"]
        b = OE.rans_encode(sy, ix, cdf, lens, offs)
        assert len(b) == rec["nbytes"] and hashlib.sha256(b).hexdigest() == rec["sha256"]
        assert np.array_equal(OE.rans_decode(b, ix, cdf, lens, offs), sy)


def test_entropy_op_vectors(ent):
    assert np.array_equal(OE.build_indexes(torch.from_numpy(ent["bi_scales"])).numpy(), ent["bi_indexes"])
    y, sc, mu = (torch.from_numpy(ent[k]) for k in ("gc_y", "gc_scale", "gc_mu"))
    out, lik = OE.gaussian_conditional_eval(y, sc, mu)
    assert np.array_equal(out.numpy(), ent["gc_out"]) and np.array_equal(lik.numpy(), ent["gc_lik"])
    q = OE.quantize(y, "symbols", mu)
    assert np.array_equal(q.numpy(), ent["gc_symbols"])
    assert np.array_equal(OE.dequantize(q, mu).numpy(), ent["gc_dequant"])


def _eb_params(seed):
    spec = {"entropy_bottleneck." + k: (s, torch.float32) for k, s in (
        [(f"_matrix{i}", (192, f1, f0)) for i, (f0, f1) in enumerate(zip(OE.EB_FILTERS[:-1], OE.EB_FILTERS[1:]))]
        + [(f"_bias{i}", (192, f1, 1)) for i, f1 in enumerate(OE.EB_FILTERS[1:])]
        + [(f"_factor{i}", (192, f1, 1)) for i, f1 in enumerate(OE.EB_FILTERS[1:-1])]
        + [("quantiles", (192, 1, 3))])}
    sd = synthetic_state_dict(spec, seed)
    return {k[len("entropy_bottleneck."):]: v for k, v in sd.items()}


def test_entropy_bottleneck_vectors(ent):
    p = _eb_params(int(ent["eb_seed"]))
    out, lik = OE.eb_forward_eval(p, torch.from_numpy(ent["eb_z"]))
    assert np.allclose(out.numpy(), ent["eb_out"], rtol=1e-6, atol=1e-6)
    assert np.allclose(lik.numpy(), ent["eb_lik"], rtol=1e-5, atol=1e-9)
    cdf, lens, offs = OE.eb_tables(p)
    assert np.array_equal(cdf, ent["eb_cdf"]) and np.array_equal(lens, ent["eb_len"]) and np.array_equal(offs, ent["eb_off"])
    z = torch.from_numpy(ent["eb_z"])
    sym = OE.quantize(z, "symbols", OE.eb_medians(p).reshape(1, -1, 1, 1))
    idx = OE.eb_indexes(z.shape)
    for b in range(z.shape[0]):
        assert OE.rans_encode(sym[b].numpy(), idx[b].numpy(), cdf, lens, offs).hex() == str(ent["eb_strings_hex"][b])


def _spec(module):
    return {k: (tuple(v.shape), v.dtype) for k, v in module.state_dict().items()}


def test_swin_vectors_against_oracle(sw):
    """Oracle restatement of the Swin ops reproduces the recorded reference outputs (weights are
    regenerated from key names, so the module classes are only used for their state_dict spec)."""
    from stf_b200 import layers as L
    for (C, nh, ws, H, W, B) in ((48, 3, 4, 8, 12, 2), (96, 6, 4, 4, 4, 1), (384, 24, 4, 8, 8, 1)):
        for shift in (0, ws // 2):
            sd = synthetic_state_dict(_spec(L.SwinTransformerBlock(C, nh, ws, shift)), 11)
            tag = f"blk_C{C}_H{H}_W{W}_s{shift}"
            mask = OS.shift_mask(H, W, ws, ws // 2)
            y = OS.swin_block(sd, "", torch.from_numpy(sw[tag + "_x"]), H, W, nh, ws, shift, mask)
            assert np.allclose(y.numpy(), sw[tag + "_y"], rtol=2e-5, atol=2e-5), tag
    sd = synthetic_state_dict(_spec(L.SwinTransformerBlock(48, 3, 4, 2)), 12)
    y = OS.swin_block(sd, "", torch.from_numpy(sw["blkpad_x"]), 6, 10, 3, 4, 2, OS.shift_mask(8, 12, 4, 2))
    assert np.allclose(y.numpy(), sw["blkpad_y"], rtol=2e-5, atol=2e-5)
    for kind, Ds in (("merge", L.PatchMerging), ("split", L.PatchSplit)):
        sd = synthetic_state_dict(_spec(L.BasicLayer(96, 2, 6, window_size=4, downsample=Ds)), 13)
        y, _, _ = OS.basic_layer(sd, "", torch.from_numpy(sw[f"layer_{kind}_x"]), 8, 8, 2, 6, 4, kind)
        assert np.allclose(y.numpy(), sw[f"layer_{kind}_y"], rtol=2e-5, atol=2e-5), kind
    assert np.array_equal(OS.shift_mask(8, 12, 4, 2).numpy(), sw["mask_8_12_4_2"])
    assert np.array_equal(OS.relative_position_index(4).numpy(), sw["relidx_4"])
    for (C, ws, H, W) in ((192, 8, 16, 24), (320, 4, 8, 12)):
        sd = synthetic_state_dict(_spec(L.WinBasedAttention(C, 8, ws, ws // 2)), 15)
        y = OS.win_based_attention(sd, "", torch.from_numpy(sw[f"wba_C{C}_x"]), 8, ws, ws // 2)
        assert np.allclose(y.numpy(), sw[f"wba_C{C}_y"], rtol=2e-5, atol=2e-5), C


@pytest.mark.parametrize("name,Ora", [("stf", OC.StfOracle), ("cnn", OC.WacnnOracle)])
def test_codec_bitstreams_match_golden(golden_dir, name, Ora):
    """End to end: the oracle codec reproduces the reference's recorded bitstreams byte for byte."""
    e2e = json.load(open(os.path.join(golden_dir, "e2e.json")))[name]
    spec = {k: (tuple(s), getattr(torch, d.split(".")[-1])) for k, (s, d) in
            json.load(open(os.path.join(golden_dir, f"{name}_spec.json"))).items()}
    sd = synthetic_state_dict(spec, e2e["weights_seed"])
    assert hashlib.sha256(b"".join(sd[k].numpy().tobytes() for k in sorted(sd))).hexdigest() == e2e["weights_sha256"]
    ora = Ora(sd)
    case = e2e["cases"][0]
    x = synthetic_image(1, case["H"], case["W"], seed=case["image_seed"])
    dbg = {}
    enc = ora.compress(x, debug=dbg)
    assert sha(dbg["symbols"].numpy().astype(np.int32)) == case["symbols_sha256"]
    assert sha(dbg["indexes"].numpy().astype(np.int32)) == case["indexes_sha256"]
    assert enc["strings"][0][0].hex() == case["y_string_hex"]
    assert [s.hex() for s in enc["strings"][1]] == case["z_strings_hex"]
    dec = ora.decompress(enc["strings"], enc["shape"])
    probe = dec["x_hat"].reshape(-1)[:: max(1, dec["x_hat"].numel() // 64)][:64]
    ref_probe = torch.tensor(case["x_hat_probe"]).clamp(0, 1)
    assert torch.allclose(probe, ref_probe, rtol=1e-4, atol=1e-4)


@pytest.mark.live_reference
@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
def test_oracle_vs_live_reference_random_ops():
    """Fresh random inputs (not the recorded ones) through the live reference and the oracle."""
    from oracle.ref_import import import_reference
    import_reference()
    from compressai.entropy_models import GaussianConditional
    from compressai.models import stf as RS
    gc = GaussianConditional(None)
    gc.update_scale_table(RS.get_scale_table())
    gc.eval()
    g = torch.Generator().manual_seed(99)
    sc = torch.exp(torch.rand(20000, generator=g) * 9 - 4.5)
    mu = torch.randn(20000, generator=g)
    y = mu + sc * torch.randn(20000, generator=g)
    assert torch.equal(gc.build_indexes(sc), OE.build_indexes(sc))
    out, lik = gc(y, sc, mu)
    o_out, o_lik = OE.gaussian_conditional_eval(y, sc, mu)
    assert torch.equal(out, o_out) and torch.equal(lik, o_lik)
    blk = RS.SwinTransformerBlock(dim=96, num_heads=6, window_size=4, shift_size=2).eval()
    sd = {k: v.clone() for k, v in blk.state_dict().items()}
    x = torch.randn(2, 12 * 8, 96, generator=g)
    blk.H, blk.W = 12, 8
    mask = OS.shift_mask(12, 8, 4, 2)
    with torch.no_grad():
        assert torch.allclose(blk(x, mask), OS.swin_block(sd, "", x, 12, 8, 6, 4, 2, mask), rtol=2e-5, atol=2e-5)


def test_training_restatement_matches_recorded_reference(golden_dir):
    """oracle forward_train (+ torch autograd) reproduces the loss and every gradient norm the live reference produced in
    train() mode (tests/golden/train_kat.json, recorded by oracle/gen_golden_train.py)."""
    import math
    from oracle import codec as OC
    from stf_b200.synth import synthetic_image
    kat = json.load(open(os.path.join(golden_dir, "train_kat.json")))
    spec = {k: (tuple(s), getattr(torch, d.split(".")[-1])) for k, (s, d) in
            json.load(open(os.path.join(golden_dir, "stf_spec.json"))).items()}
    sd = synthetic_state_dict(spec, kat["weights_seed"])
    ora = OC.StfOracle(sd)
    for k, v in ora.sd.items():
        if k in kat["grad_norm"]:
            v.requires_grad_(True)
    im = kat["image"]
    x = synthetic_image(im["B"], im["H"], im["W"], seed=im["seed"])
    noise = OC.train_noise(kat["noise_seed"], im["B"], 384, im["H"] // 16, im["W"] // 16, 192, im["H"] // 64, im["W"] // 64)
    out = ora.forward_train(x, noise)
    n_px = im["B"] * im["H"] * im["W"]
    bpp = sum(torch.log(l).sum() / (-math.log(2) * n_px) for l in out["likelihoods"].values())
    loss = kat["lmbda"] * 255 ** 2 * torch.nn.functional.mse_loss(out["x_hat"], x) + bpp
    loss.backward()
    assert abs(float(loss.detach()) - kat["loss"]) <= 1e-5 * abs(kat["loss"])
    for n, ref in kat["grad_norm"].items():
        got = float(ora.sd[n].grad.norm())
        assert abs(got - ref) <= 1e-4 * ref + 1e-9, (n, got, ref)


def test_wacnn_training_restatement_matches_recorded_reference(golden_dir):
    """Same pin for WACNN (cnn.py) in train() mode: tests/golden/train_kat_cnn.json."""
    import math
    from oracle import codec as OC
    from stf_b200.synth import synthetic_image
    kat = json.load(open(os.path.join(golden_dir, "train_kat_cnn.json")))
    spec = {k: (tuple(s), getattr(torch, d.split(".")[-1])) for k, (s, d) in
            json.load(open(os.path.join(golden_dir, "cnn_spec.json"))).items()}
    ora = OC.WacnnOracle(synthetic_state_dict(spec, kat["weights_seed"]))
    for k, v in ora.sd.items():
        if k in kat["grad_norm"]:
            v.requires_grad_(True)
    im = kat["image"]
    x = synthetic_image(im["B"], im["H"], im["W"], seed=im["seed"])
    B, M, h, w, Cz, hz, wz, slices = kat["latent"]
    noise = OC.train_noise(kat["noise_seed"], B, M, h, w, Cz, hz, wz, num_slices=slices)
    out = ora.forward_train(x, noise)
    n_px = im["B"] * im["H"] * im["W"]
    bpp = sum(torch.log(l).sum() / (-math.log(2) * n_px) for l in out["likelihoods"].values())
    loss = kat["lmbda"] * 255 ** 2 * torch.nn.functional.mse_loss(out["x_hat"], x) + bpp
    loss.backward()
    assert abs(float(loss.detach()) - kat["loss"]) <= 1e-5 * abs(kat["loss"])
    for n, ref in kat["grad_norm"].items():
        got = float(ora.sd[n].grad.norm())
        assert abs(got - ref) <= 1e-4 * ref + 1e-9, (n, got, ref)
