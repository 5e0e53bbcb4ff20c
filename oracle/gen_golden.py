"""TEST INFRASTRUCTURE ONLY -- records tests/golden/* from the LIVE, UNMODIFIED reference.

Run in the build container (needs /root/reference and `make -C oracle`):

    python -m oracle.gen_golden

The reference ships no tests and no golden vectors, so these fixtures are
the pin: every value written here comes out of the reference's own classes / compiled extension,
driven with deterministic synthetic weights (stf_b200/synth.py) and seeded inputs.  While
recording, the script also checks the CPU restatement in oracle/ against the same live outputs
and aborts on any disagreement, so a committed fixture implies "oracle == reference" at
generation time.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import codec as OC  # noqa: E402
from oracle import entropy as OE  # noqa: E402
from oracle import swin as OS  # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402
from stf_b200.synth import synthetic_image, synthetic_state_dict  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def spec_of(module):
    return {k: (tuple(v.shape), v.dtype) for k, v in module.state_dict().items()}


def load_synth(module, seed):
    sd = synthetic_state_dict(spec_of(module), seed)
    missing, unexpected = module.load_state_dict(sd, strict=False) if not hasattr(module, "entropy_bottleneck") \
        else torch.nn.Module.load_state_dict(module, sd, strict=False)
    assert not unexpected, unexpected
    return {k: v.clone() for k, v in module.state_dict().items()}


def close(a, b, rtol=1e-5, atol=1e-6, what=""):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    ok = torch.allclose(a.float(), b.float(), rtol=rtol, atol=atol)
    if not ok:
        raise SystemExit(f"oracle != reference for {what}: max abs diff {(a.float() - b.float()).abs().max().item()}")


def edge_scales(table):
    t = table.numpy()
    vals = [t, np.nextafter(t, np.float32(np.inf), dtype=np.float32), np.nextafter(t, np.float32(-np.inf), dtype=np.float32),
            np.array([-1.0, 0.0, 0.05, 0.11, 0.1100001, 0.12, 0.5, 1.0, 100.0, 256.0, 300.0, np.inf, -np.inf, np.nan], dtype=np.float32)]
    return torch.from_numpy(np.concatenate(vals).astype(np.float32))


def main():
    os.makedirs(GOLD, exist_ok=True)
    C = import_reference()
    from compressai.entropy_models import EntropyBottleneck, GaussianConditional
    from compressai.models import stf as RS
    from compressai.layers import win_attention as RW
    from compressai.zoo import models as zoo
    import compressai.ans as ans

    kat = {}
    torch.manual_seed(0)

    # ------------------------------------------------------------------ Gaussian tables + KATs
    table = RS.get_scale_table()
    gc = GaussianConditional(None)
    gc.update_scale_table(table)
    gc.eval()
    cdf, lens, offs = gc.quantized_cdf.numpy(), gc.cdf_length.numpy(), gc.offset.numpy()
    o_cdf, o_len, o_off = OE.gaussian_tables()
    assert np.array_equal(cdf, o_cdf) and np.array_equal(lens, o_len) and np.array_equal(offs, o_off), "gaussian tables"
    assert torch.equal(table, OE.scale_table())
    kat["scale_table_head"] = [float(v) for v in table[:4]] + [float(table[63])]
    kat["scale_table_sha256"] = sha(table.numpy())
    kat["gc_cdf_shape"] = list(cdf.shape)
    kat["gc_cdf_sha256"] = sha(cdf.astype(np.int32))
    kat["gc_cdf_row0_head"] = cdf[0, :6].tolist()
    kat["gc_cdf_length"] = lens.tolist()
    kat["gc_offset"] = offs.tolist()

    s_in = [-1, 0, .05, .11, .1100001, .12, .5, 1, 100, 256, 300, float("inf"), float("nan")]
    kat["build_indexes_in"] = [repr(float(v)) for v in s_in]
    kat["build_indexes_out"] = gc.build_indexes(torch.tensor(s_in)).tolist()
    kat["quantize_in"] = [0.5, 1.5, 2.5, -0.5, -1.5]
    kat["quantize_out"] = gc.quantize(torch.tensor(kat["quantize_in"]), "symbols").tolist()
    y = torch.tensor([0, .3, -.7, 2.5, 10, -100.0])
    sg = torch.tensor([.05, .11, .5, 1, 3, .11])
    mu = torch.tensor([0, .1, -.2, .4, -1, 0.0])
    out, lik = gc(y, sg, mu)
    kat["gc_forward"] = {"y": y.tolist(), "scale": sg.tolist(), "mu": mu.tolist(), "out": out.tolist(), "lik": [float(v) for v in lik]}
    o_out, o_lik = OE.gaussian_conditional_eval(y, sg, mu)
    assert torch.equal(out, o_out) and torch.equal(lik, o_lik)

    sym = [0, 1, -1, 2, -3, 0, 0, 5, -40, 100]
    idx = [0, 5, 10, 20, 30, 40, 50, 63, 0, 1]
    bs = ans.RansEncoder().encode_with_indexes(sym, idx, cdf.tolist(), lens.tolist(), offs.tolist())
    kat["rans"] = {"symbols": sym, "indexes": idx, "bytes_hex": bs.hex()}
    assert OE.rans_encode(sym, idx, cdf, lens, offs) == bs
    assert ans.RansDecoder().decode_with_indexes(bs, idx, cdf.tolist(), lens.tolist(), offs.tolist()) == sym
    assert OE.rans_decode(bs, idx, cdf, lens, offs).tolist() == sym

    # larger random rANS streams incl. many escapes: reference bytes recorded by hash
    rng = np.random.default_rng(1234)
    streams = []
    for n, spread in ((1, 3.0), (7, 0.5), (5000, 1.0), (60000, 4.0)):
        ix = rng.integers(0, 64, size=n).astype(np.int32)
        sc = table.numpy()[ix]
        sy = np.rint(rng.standard_normal(n) * sc * spread).astype(np.int32)
        if n < 4:
            sy = np.array([70000], dtype=np.int32)[:n]       # multi-nibble escape on a tiny stream
        b = ans.RansEncoder().encode_with_indexes(sy.tolist(), ix.tolist(), cdf.tolist(), lens.tolist(), offs.tolist())
        assert OE.rans_encode(sy, ix, cdf, lens, offs) == b, f"rans n={n}"
        assert OE.rans_decode(b, ix, cdf, lens, offs).tolist() == sy.tolist()
        streams.append({"n": n, "spread": spread, "nbytes": len(b), "sha256": hashlib.sha256(b).hexdigest(),
                        "symbols_sha256": sha(sy), "indexes_sha256": sha(ix)})
    kat["rans_streams_seed"] = 1234
    kat["rans_streams"] = streams

    # ------------------------------------------------------------------ entropy op vectors
    ent = {}
    g = torch.Generator().manual_seed(7)
    scales = torch.cat([edge_scales(table), torch.exp(torch.rand(4096, generator=g) * (np.log(400) - np.log(0.01)) + np.log(0.01))])
    ent["bi_scales"] = scales.numpy()
    ent["bi_indexes"] = gc.build_indexes(scales).numpy()
    assert torch.equal(gc.build_indexes(scales), OE.build_indexes(scales))

    n = 8192
    sc = torch.exp(torch.rand(n, generator=g) * (np.log(400) - np.log(0.01)) + np.log(0.01))
    mu = 2 * torch.randn(n, generator=g)
    yy = mu + sc * torch.randn(n, generator=g)
    ties = torch.arange(-8, 8).float() + 0.5
    yy[:16] = mu[:16] + ties                      # exact k+1/2 ties (where fp32 keeps them exact)
    mu[16:32] = 0.0
    yy[16:32] = ties
    ent["gc_y"], ent["gc_scale"], ent["gc_mu"] = yy.numpy(), sc.numpy(), mu.numpy()
    out, lik = gc(yy, sc, mu)
    ent["gc_out"], ent["gc_lik"] = out.numpy(), lik.numpy()
    ent["gc_symbols"] = gc.quantize(yy, "symbols", mu).numpy()
    ent["gc_dequant"] = gc.dequantize(gc.quantize(yy, "symbols", mu), mu).numpy()
    o_out, o_lik = OE.gaussian_conditional_eval(yy, sc, mu)
    assert torch.equal(out, o_out) and torch.equal(lik, o_lik)
    assert torch.equal(gc.quantize(yy, "symbols", mu), OE.quantize(yy, "symbols", mu))

    eb = EntropyBottleneck(192)
    sd = synthetic_state_dict({("entropy_bottleneck." + k): v for k, v in spec_of(eb).items()}, 3)
    eb.load_state_dict({k[len("entropy_bottleneck."):]: v for k, v in sd.items()}, strict=False)
    eb.eval()
    eb.update(force=True)
    z = 4 * torch.randn(2, 192, 3, 5, generator=g)
    with torch.no_grad():
        z_out, z_lik = eb(z)
    p = {k: v.detach() for k, v in eb.state_dict().items() if k in OE.eb_param_names()}
    o_out, o_lik = OE.eb_forward_eval(p, z)
    close(z_out, o_out, what="eb out")
    close(z_lik, o_lik, rtol=1e-5, atol=1e-9, what="eb lik")
    e_cdf, e_len, e_off = OE.eb_tables(p)
    assert np.array_equal(e_cdf, eb.quantized_cdf.numpy()) and np.array_equal(e_len, eb.cdf_length.numpy()) \
        and np.array_equal(e_off, eb.offset.numpy()), "eb tables"
    ent["eb_seed"] = np.array(3)
    ent["eb_z"], ent["eb_out"], ent["eb_lik"] = z.numpy(), z_out.numpy(), z_lik.numpy()
    ent["eb_cdf"], ent["eb_len"], ent["eb_off"] = e_cdf, e_len, e_off
    zs = eb.compress(z)
    ent["eb_strings_hex"] = np.array([s.hex() for s in zs])
    with torch.no_grad():
        ent["eb_zhat"] = eb.decompress(zs, z.shape[-2:]).numpy()
    np.savez_compressed(os.path.join(GOLD, "entropy_ops.npz"), **ent)

    # ------------------------------------------------------------------ window-attention op vectors
    sw = {}

    def rec(name, ref_out, ora_out, tol=2e-5):
        close(ref_out, ora_out, rtol=tol, atol=tol, what=name)
        sw[name] = ref_out.detach().numpy()

    with torch.no_grad():
        for (C_, nh, ws_, Hh, Ww, B_) in ((48, 3, 4, 8, 12, 2), (96, 6, 4, 4, 4, 1), (384, 24, 4, 8, 8, 1)):
            for shift in (0, ws_ // 2):
                blk = RS.SwinTransformerBlock(dim=C_, num_heads=nh, window_size=ws_, shift_size=shift).eval()
                sdm = load_synth(blk, 11)
                x = torch.randn(B_, Hh * Ww, C_, generator=g)
                blk.H, blk.W = Hh, Ww
                mask = OS.shift_mask(Hh, Ww, ws_, ws_ // 2)
                tag = f"blk_C{C_}_H{Hh}_W{Ww}_s{shift}"
                sw[tag + "_x"] = x.numpy()
                rec(tag + "_y", blk(x, mask), OS.swin_block(sdm, "", x, Hh, Ww, nh, ws_, shift, mask))
        # padded (non window-aligned) block: H=6, W=10 with ws=4
        blk = RS.SwinTransformerBlock(dim=48, num_heads=3, window_size=4, shift_size=2).eval()
        sdm = load_synth(blk, 12)
        x = torch.randn(1, 60, 48, generator=g)
        blk.H, blk.W = 6, 10
        mask = OS.shift_mask(8, 12, 4, 2)
        sw["blkpad_x"] = x.numpy()
        rec("blkpad_y", blk(x, mask), OS.swin_block(sdm, "", x, 6, 10, 3, 4, 2, mask))

        # BasicLayer with PatchMerging / PatchSplit
        for kind, Ds in (("merge", RS.PatchMerging), ("split", RS.PatchSplit)):
            layer = RS.BasicLayer(dim=96, depth=2, num_heads=6, window_size=4, downsample=Ds).eval()
            sdm = load_synth(layer, 13)
            x = torch.randn(2, 8 * 8, 96, generator=g)
            yref, h2, w2 = layer(x, 8, 8)
            yora, h3, w3 = OS.basic_layer(sdm, "", x, 8, 8, 2, 6, 4, kind)
            assert (h2, w2) == (h3, w3)
            sw[f"layer_{kind}_x"] = x.numpy()
            rec(f"layer_{kind}_y", yref, yora)

        # raw WindowAttention with explicit mask
        wa = RS.WindowAttention(48, (4, 4), 3).eval()
        sdm = load_synth(wa, 14)
        x = torch.randn(12, 16, 48, generator=g)
        mask = OS.shift_mask(8, 12, 4, 2)
        sw["wa_x"] = x.numpy()
        rec("wa_y_nomask", wa(x), OS.window_attention(sdm, "", x, 3, 4, None))
        rec("wa_y_mask", wa(x, mask), OS.window_attention(sdm, "", x, 3, 4, mask))
        sw["mask_8_12_4_2"] = mask.numpy()
        sw["relidx_4"] = wa.relative_position_index.numpy()
        assert torch.equal(wa.relative_position_index, OS.relative_position_index(4))

        # WACNN attention variants (NCHW, always shifted): ws=8,d=24 and ws=4,d=40
        for (C_, ws_, Hh, Ww) in ((192, 8, 16, 24), (320, 4, 8, 12)):
            m = RW.WinBasedAttention(dim=C_, num_heads=8, window_size=ws_, shift_size=ws_ // 2).eval()
            sdm = load_synth(m, 15)
            x = torch.randn(2, C_, Hh, Ww, generator=g)
            sw[f"wba_C{C_}_x"] = x.numpy()
            rec(f"wba_C{C_}_y", m(x), OS.win_based_attention(sdm, "", x, 8, ws_, ws_ // 2))
    np.savez_compressed(os.path.join(GOLD, "swin_ops.npz"), **sw)

    # ------------------------------------------------------------------ full models
    e2e = {}
    for name, Ora, sizes in (("stf", OC.StfOracle, ((64, 64), (128, 192))), ("cnn", OC.WacnnOracle, ((64, 128),))):
        torch.manual_seed(0)
        net = zoo[name]().eval()
        spec = spec_of(net)
        with open(os.path.join(GOLD, f"{name}_spec.json"), "w") as f:
            json.dump({k: [list(s), str(d)] for k, (s, d) in spec.items()}, f, indent=0)
        sd = synthetic_state_dict(spec, 0)
        torch.nn.Module.load_state_dict(net, sd, strict=False)
        net.update(force=True)
        full_sd = {k: v.clone() for k, v in net.state_dict().items()}
        ora = Ora(full_sd, rans="oracle")
        assert np.array_equal(ora.eb_cdf, net.entropy_bottleneck.quantized_cdf.numpy())
        e2e[name] = {"weights_seed": 0, "n_keys": len(spec),
                     "n_params": int(sum(p.numel() for p in net.parameters())),
                     "weights_sha256": hashlib.sha256(b"".join(sd[k].numpy().tobytes() for k in sorted(sd))).hexdigest(),
                     "cases": []}
        for (Hh, Ww) in sizes:
            x = synthetic_image(1, Hh, Ww, seed=5)
            with torch.no_grad():
                fwd = net(x)
                enc = net.compress(x)
                dec = net.decompress(enc["strings"], enc["shape"])
            dbg = {}
            o_fwd = ora.forward(x)
            o_enc = ora.compress(x, debug=dbg)
            o_dec = ora.decompress(o_enc["strings"], o_enc["shape"])
            close(fwd["x_hat"], o_fwd["x_hat"], rtol=1e-4, atol=1e-4, what=f"{name} x_hat")
            close(fwd["likelihoods"]["y"], o_fwd["likelihoods"]["y"], rtol=1e-4, atol=1e-7, what=f"{name} y lik")
            close(fwd["likelihoods"]["z"], o_fwd["likelihoods"]["z"], rtol=1e-4, atol=1e-7, what=f"{name} z lik")
            assert o_enc["strings"][0][0] == enc["strings"][0][0], f"{name} y string"
            assert o_enc["strings"][1] == enc["strings"][1], f"{name} z strings"
            close(dec["x_hat"], o_dec["x_hat"], rtol=1e-4, atol=1e-4, what=f"{name} dec x_hat")
            bpp_y = float(-torch.log2(fwd["likelihoods"]["y"]).sum() / (Hh * Ww))
            bpp_z = float(-torch.log2(fwd["likelihoods"]["z"]).sum() / (Hh * Ww))
            idx_hist = np.bincount(dbg["indexes"].numpy(), minlength=64).tolist()
            e2e[name]["cases"].append({
                "H": Hh, "W": Ww, "image_seed": 5,
                "y_string_hex": enc["strings"][0][0].hex(),
                "z_strings_hex": [s.hex() for s in enc["strings"][1]],
                "z_shape": list(enc["shape"]),
                "symbols_sha256": sha(dbg["symbols"].numpy().astype(np.int32)),
                "indexes_sha256": sha(dbg["indexes"].numpy().astype(np.int32)),
                "index_histogram": idx_hist,
                "symbol_absmax": int(dbg["symbols"].abs().max()),
                "bpp_y_est": bpp_y, "bpp_z_est": bpp_z,
                "x_hat_mean": float(fwd["x_hat"].mean()), "x_hat_std": float(fwd["x_hat"].std()),
                "x_hat_probe": fwd["x_hat"].reshape(-1)[:: max(1, fwd["x_hat"].numel() // 64)][:64].tolist(),
                "dec_equals_fwd_clamped_maxabs": float((dec["x_hat"] - fwd["x_hat"].clamp(0, 1)).abs().max()),
            })
            print(name, Hh, Ww, "y bytes", len(enc["strings"][0][0]), "idx used", sum(1 for c in idx_hist if c),
                  "absmax", e2e[name]["cases"][-1]["symbol_absmax"], "bpp", bpp_y + bpp_z)
    with open(os.path.join(GOLD, "e2e.json"), "w") as f:
        json.dump(e2e, f, indent=1)
    with open(os.path.join(GOLD, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("golden written to", GOLD)


if __name__ == "__main__":
    main()
