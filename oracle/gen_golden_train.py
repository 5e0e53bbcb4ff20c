"""TEST INFRASTRUCTURE ONLY.  Record the training-step golden vector (tests/golden/train_kat.json) from the LIVE,
unmodified reference in train() mode and pin the oracle's differentiable restatement (oracle/codec.py forward_train,
oracle/entropy.py *_train) to it.  Run in the build container (needs /root/reference):

    python -m oracle.gen_golden_train

Setting: SymmetricalTransFormer(drop_path_rate=0).train(), synthetic weights
(stf_b200/synth.py seed 0), x = synthetic_image(2, 64, 64, seed 7), RateDistortionLoss(lambda = 0.0035)
(train.py:39-59).  The quantisation noise is the reference's own `torch.empty_like(.).uniform_(-1/2, 1/2)` stream
under torch.manual_seed(NOISE_SEED); oracle.codec.train_noise re-draws the same stream (same order and shapes:
entropy_models.py:131-135 is called first for z as (C, 1, B*h*w), then for the 12 y slices) so that the oracle -- and
the CUDA path in tests/test_gpu_train.py -- can be fed the identical tensors.
"""
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import codec as OC  # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402
from stf_b200.synth import synthetic_image, synthetic_state_dict  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
NOISE_SEED, LMBDA = 1234, 0.0035


def rd_loss(out, x, lmbda=LMBDA):
    """train.py:39-59."""
    N, _, H, W = x.shape
    bpp = sum(torch.log(l).sum() / (-math.log(2) * N * H * W) for l in out["likelihoods"].values())
    mse = torch.nn.functional.mse_loss(out["x_hat"], x)
    return lmbda * 255 ** 2 * mse + bpp, bpp, mse


def record(name, net, Ora, x, latent, probes, out_file):
    """Live reference in train() mode vs the oracle's restatement on the same noise stream -> golden file."""
    spec = {k: (tuple(v.shape), v.dtype) for k, v in net.state_dict().items()}
    sd = synthetic_state_dict(spec, 0)
    torch.nn.Module.load_state_dict(net, sd, strict=False)
    net.train()
    torch.manual_seed(NOISE_SEED)
    out = net(x)
    loss, bpp, mse = rd_loss(out, x)
    loss.backward()
    ref_grads = {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}
    ora = Ora({k: v.clone() for k, v in net.state_dict().items()})
    names = dict(net.named_parameters())
    for k, v in ora.sd.items():
        if v.is_floating_point() and k in names:
            v.requires_grad_(True)
    B, M, h, w, Cz, hz, wz, slices = latent
    noise = OC.train_noise(NOISE_SEED, B, M, h, w, Cz, hz, wz, num_slices=slices)
    o_loss, _, _ = rd_loss(ora.forward_train(x, noise), x)
    o_loss.backward()
    assert abs(float(o_loss.detach()) - float(loss.detach())) <= 1e-5 * abs(float(loss.detach())), (float(o_loss), float(loss))
    worst = 0.0
    for n, g in ref_grads.items():
        og = ora.sd[n].grad
        assert og is not None, n
        err = float((og - g).abs().max()) / (float(g.abs().max()) + 1e-12)
        worst = max(worst, err)
        assert err <= 2e-3, (n, err)
    print(f"{name}: oracle forward_train == live reference: loss {float(loss.detach())}, worst relative gradient error {worst}")
    kat = {"noise_seed": NOISE_SEED, "lmbda": LMBDA, "image": {"B": x.shape[0], "H": x.shape[2], "W": x.shape[3], "seed": 7},
           "weights_seed": 0, "latent": list(latent), "loss": float(loss.detach()), "bpp_loss": float(bpp.detach()),
           "mse_loss": float(mse.detach()),
           "grad_norm": {n: float(g.norm()) for n, g in ref_grads.items()},
           "grad_absmax": {n: float(g.abs().max()) for n, g in ref_grads.items()},
           "grad_probe": {n: ref_grads[n].reshape(-1)[:: max(1, ref_grads[n].numel() // 16)][:16].tolist() for n in probes}}
    with open(os.path.join(GOLD, out_file), "w") as f:
        json.dump(kat, f, indent=0)
    print("wrote", os.path.join(GOLD, out_file), len(kat["grad_norm"]), "parameters")


def main_wacnn():
    import_reference()
    from compressai.models.cnn import WACNN
    torch.manual_seed(0)
    x = synthetic_image(1, 64, 128, seed=7)
    record("cnn", WACNN(), OC.WacnnOracle, x, (1, 320, 4, 8, 192, 1, 2, 10),
           ("g_a.4.conv_b.0.attn.relative_position_bias_table", "g_a.8.conv_b.0.attn.qkv.weight", "g_s.0.conv_b.0.attn.proj.weight",
            "entropy_bottleneck._matrix1", "cc_mean_transforms.2.0.weight"), "train_kat_cnn.json")


def main():
    import_reference()
    from compressai.models.stf import SymmetricalTransFormer
    torch.manual_seed(0)
    net = SymmetricalTransFormer(drop_path_rate=0.0)
    spec = {k: (tuple(v.shape), v.dtype) for k, v in net.state_dict().items()}
    sd = synthetic_state_dict(spec, 0)
    torch.nn.Module.load_state_dict(net, sd, strict=False)
    net.train()
    x = synthetic_image(2, 64, 64, seed=7)
    torch.manual_seed(NOISE_SEED)
    out = net(x)
    loss, bpp, mse = rd_loss(out, x)
    loss.backward()
    ref_grads = {n: p.grad.detach().clone() for n, p in net.named_parameters() if p.grad is not None}

    ora = OC.StfOracle({k: v.clone() for k, v in net.state_dict().items()})
    for k, v in ora.sd.items():
        if v.is_floating_point() and k in dict(net.named_parameters()):
            v.requires_grad_(True)
    noise = OC.train_noise(NOISE_SEED, 2, 384, 4, 4, 192, 1, 1)
    o_out = ora.forward_train(x, noise)
    o_loss, o_bpp, o_mse = rd_loss(o_out, x)
    o_loss.backward()
    assert abs(float(o_loss) - float(loss)) <= 1e-5 * abs(float(loss)), (float(o_loss), float(loss))
    worst = 0.0
    for n, g in ref_grads.items():
        og = ora.sd[n].grad
        assert og is not None, n
        err = float((og - g).abs().max()) / (float(g.abs().max()) + 1e-12)
        worst = max(worst, err)
        assert err <= 2e-3, (n, err)
    print("oracle forward_train == live reference: loss", float(loss), "worst relative gradient error", worst)
    kat = {"noise_seed": NOISE_SEED, "lmbda": LMBDA, "image": {"B": 2, "H": 64, "W": 64, "seed": 7}, "weights_seed": 0,
           "loss": float(loss), "bpp_loss": float(bpp), "mse_loss": float(mse),
           "grad_norm": {n: float(g.norm()) for n, g in ref_grads.items()},
           "grad_absmax": {n: float(g.abs().max()) for n, g in ref_grads.items()},
           "grad_probe": {n: ref_grads[n].reshape(-1)[:: max(1, ref_grads[n].numel() // 16)][:16].tolist()
                          for n in ("layers.0.blocks.0.attn.relative_position_bias_table", "layers.0.blocks.1.attn.qkv.weight",
                                    "layers.2.blocks.3.mlp.fc1.weight", "syn_layers.3.blocks.1.norm1.weight",
                                    "entropy_bottleneck._matrix2", "patch_embed.proj.weight",
                                    "cc_scale_transforms.3.8.bias")}}
    with open(os.path.join(GOLD, "train_kat.json"), "w") as f:
        json.dump(kat, f, indent=0)
    print("wrote", os.path.join(GOLD, "train_kat.json"), len(kat["grad_norm"]), "parameters")


if __name__ == "__main__":
    if "--stf-only" not in sys.argv:
        main_wacnn()
    if "--cnn-only" not in sys.argv:
        main()
