"""Synthetic checkpoints and inputs (no dataset / no trained weights are available offline).

`synthetic_state_dict(spec, seed)` fills a reference-format state_dict deterministically from the
key names alone (one torch CPU generator per key, seeded by a hash of the key), so the reference,
the CPU oracle and the CUDA path can all be given bit-identical weights on any machine without
depending on constructor RNG order.

Why not constructor defaults: with random-init weights every predicted scale is ~0, so
`build_indexes` returns 0 everywhere and the 64-row CDF table / escape coding are never exercised
(SURVEY.md F5).  The synthetic fill therefore spreads the last bias of every
`cc_scale_transforms[i]` log-uniformly over [0.05, 64], gives the entropy bottleneck non-zero
medians and tanh gates, and uses O(1) relative-position biases.
"""
import hashlib
import math

import torch


def _gen(key: str, seed: int) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF)
    return g


def _uniform(shape, lo, hi, g):
    return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo


def synthetic_state_dict(spec, seed: int = 0):
    """spec: mapping key -> (shape tuple, torch.dtype) of a reference-format state_dict
    (e.g. {k: (tuple(v.shape), v.dtype) for k, v in model.state_dict().items()}).
    Buffers that the modules compute themselves (relative_position_index, CDF tables, bounds,
    pedestals, targets) are skipped: they are returned only for float parameters."""
    out = {}
    for key in sorted(spec):
        shape, dtype = spec[key]
        shape = tuple(shape)
        if dtype not in (torch.float32, "torch.float32", "float32"):
            continue
        leaf = key.rsplit(".", 1)[-1]
        if leaf in ("bound", "pedestal", "target", "scale_bound", "scale_table") or 0 in shape:
            continue
        g = _gen(key, seed)
        if leaf == "relative_position_bias_table":
            t = 0.5 * torch.randn(shape, generator=g)
        elif key.startswith("entropy_bottleneck."):
            t = _eb_param(leaf, shape, g)
        elif leaf in ("beta", "gamma"):               # GDN (WACNN): reparametrised values
            ped = (2.0 ** -18) ** 2
            if leaf == "beta":
                t = torch.sqrt(_uniform(shape, 0.5, 1.5, g) + ped)
            else:
                t = torch.sqrt(0.1 * torch.eye(shape[0]) + _uniform(shape, 0.0, 2e-3, g) + ped)
        elif ".norm" in key and leaf == "weight" and len(shape) == 1:
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif ".norm" in key and leaf == "bias":
            t = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "weight":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            b = 1.0 / math.sqrt(fan_in)
            t = _uniform(shape, -b, b, g)
        elif leaf == "bias":
            parts = key.split(".")
            if parts[0] == "cc_scale_transforms" and parts[2] == "8":
                t = torch.exp(_uniform(shape, math.log(0.05), math.log(64.0), g))
            elif parts[0] == "cc_mean_transforms" and parts[2] == "8":
                t = 2.0 * torch.randn(shape, generator=g)
            elif parts[0] == "h_a" and parts[1] == "8":
                t = 3.0 * torch.randn(shape, generator=g)
            else:
                t = _uniform(shape, -0.1, 0.1, g)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
        out[key] = t.to(torch.float32).contiguous()
    return out


def _eb_param(leaf, shape, g):
    """EntropyBottleneck parameters around their constructor init (reference
    entropy_models.py:323-345) but with non-trivial gates and medians."""
    filters = (1, 3, 3, 3, 3, 1)
    scale = 10.0 ** (1 / 5)
    if leaf.startswith("_matrix"):
        i = int(leaf[-1])
        init = math.log(math.expm1(1 / scale / filters[i + 1]))
        return init + 0.2 * torch.randn(shape, generator=g)
    if leaf.startswith("_bias"):
        return _uniform(shape, -0.5, 0.5, g)
    if leaf.startswith("_factor"):
        return 0.3 * torch.randn(shape, generator=g)
    if leaf == "quantiles":
        med = 2.0 * torch.randn((shape[0], 1, 1), generator=g)
        lo = _uniform((shape[0], 1, 1), 4.0, 12.0, g)
        hi = _uniform((shape[0], 1, 1), 4.0, 12.0, g)
        return torch.cat([med - lo, med, med + hi], dim=2)
    raise KeyError(leaf)


def synthetic_image(batch, height, width, seed=0):
    """x ~ U[0,1) like ToTensor() images, CPU fp32, deterministic."""
    g = torch.Generator(device="cpu")
    g.manual_seed(1_000_003 * seed + 17)
    return torch.rand((batch, 3, height, width), generator=g, dtype=torch.float32)
