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
