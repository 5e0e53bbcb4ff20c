"""CPU: host-side logic of the module mirrors that needs no GPU -- checkpoint key space, table
construction in update(), dynamic-buffer loading, error behaviour, and the 'no CPU fallback' rule."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import entropy as OE


def _spec(golden_dir, name):
    return {k: (tuple(s), d) for k, (s, d) in json.load(open(os.path.join(golden_dir, f"{name}_spec.json"))).items()}


@pytest.mark.parametrize("name", ["stf", "cnn"])
def test_state_dict_keys_and_shapes_match_reference(golden_dir, name):
    from stf_b200.models import models
    net = models[name]()
    spec = _spec(golden_dir, name)
    ours = {k: (tuple(v.shape), str(v.dtype)) for k, v in net.state_dict().items()}
    assert list(ours) == list(spec), "key ORDER differs too (state_dict is ordered)" if set(ours) == set(spec) else \
        sorted(set(ours) ^ set(spec))[:10]
    for k in spec:
        assert ours[k] == spec[k], (k, ours[k], spec[k])
    e2e = json.load(open(os.path.join(golden_dir, "e2e.json")))[name]
    assert sum(p.numel() for p in net.parameters()) == e2e["n_params"]


def test_update_builds_reference_tables_and_reloads(golden_dir):
    from stf_b200.models import SymmetricalTransFormer
    net = SymmetricalTransFormer()
    assert net.update() is True and net.update() is False and net.update(force=True) is True
    cdf, lens, offs = OE.gaussian_tables()
    gc = net.gaussian_conditional
    assert np.array_equal(gc.quantized_cdf.numpy(), cdf) and np.array_equal(gc.cdf_length.numpy(), lens)
    assert np.array_equal(gc.offset.numpy(), offs)
    assert tuple(net.entropy_bottleneck.quantized_cdf.shape) == (192, 23)
    sd = net.state_dict()
    net2 = SymmetricalTransFormer.from_state_dict(sd)                               # dynamic-size buffers
    assert torch.equal(net2.gaussian_conditional.quantized_cdf, gc.quantized_cdf)
    assert float(net.aux_loss()) > 0


def test_no_cpu_fallback():
    from stf_b200 import layers, ops
    from stf_b200.entropy_models import GaussianConditional
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.build_indexes(torch.ones(8), OE.scale_table())
    gc = GaussianConditional(None).eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        gc(torch.zeros(4), torch.ones(4), torch.zeros(4))
    blk = layers.SwinTransformerBlock(48, 3, 4, 0).eval()
    blk.H, blk.W = 4, 4
    with pytest.raises(RuntimeError):
        blk(torch.zeros(1, 16, 48), None)


def test_entropy_model_argument_errors():
    from stf_b200.entropy_models import EntropyBottleneck, GaussianConditional
    with pytest.raises(ValueError):
        GaussianConditional([3.0, 1.0])
    with pytest.raises(ValueError):
        GaussianConditional("abc")
    with pytest.raises(ValueError):
        GaussianConditional(None, scale_bound=0)
    eb = EntropyBottleneck(8)
    with pytest.raises(ValueError, match="Uninitialized"):
        eb.decompress([b"12345678"], (2, 2))
    with pytest.raises(ValueError):
        eb.quantize(torch.zeros(2), "nope")


def test_rate_distortion_loss_and_optimizer_split():
    """train.py:39-59 (loss) and train.py:88-120 (the two Adam parameter sets) -- host logic, CPU only."""
    import math
    from stf_b200.models import SymmetricalTransFormer, WACNN
    from stf_b200.training import RateDistortionLoss, configure_optimizers
    g = torch.Generator().manual_seed(0)
    x = torch.rand(2, 3, 16, 24, generator=g)
    xh = torch.rand(2, 3, 16, 24, generator=g)
    lik = {"y": torch.rand(2, 8, 1, 2, generator=g) * 0.9 + 0.05, "z": torch.rand(2, 4, 1, 1, generator=g) * 0.9 + 0.05}
    out = RateDistortionLoss(0.0035)({"x_hat": xh, "likelihoods": lik}, x)
    bpp = -(np.log2(lik["y"].double().numpy()).sum() + np.log2(lik["z"].double().numpy()).sum()) / (2 * 16 * 24)
    mse = float(((xh.double() - x.double()) ** 2).mean())
    assert abs(float(out["bpp_loss"]) - bpp) <= 1e-5 * bpp
    assert abs(float(out["mse_loss"]) - mse) <= 1e-5 * mse
    assert abs(float(out["loss"]) - (0.0035 * 255 ** 2 * mse + bpp)) <= 1e-5 * float(out["loss"])
    assert math.isfinite(float(out["loss"]))
    for Net in (SymmetricalTransFormer, WACNN):
        net = Net()
        opt, aux = configure_optimizers(net, 1e-4, 1e-3)
        named = dict(net.named_parameters())
        aux_ids = {id(p) for grp in aux.param_groups for p in grp["params"]}
        main_ids = {id(p) for grp in opt.param_groups for p in grp["params"]}
        assert aux_ids == {id(p) for n, p in named.items() if n.endswith(".quantiles")} and len(aux_ids) == 1
        assert not aux_ids & main_ids and len(aux_ids) + len(main_ids) == len(named)
        assert opt.param_groups[0]["lr"] == 1e-4 and aux.param_groups[0]["lr"] == 1e-3
