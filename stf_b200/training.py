"""Rate-distortion training step of the reference's train.py (BASELINE config 5) on the stf_b200 kernels.

  RateDistortionLoss        train.py:39-59
  configure_optimizers      train.py:88-120   (main Adam on everything but *.quantiles, aux Adam on *.quantiles)
  train_step                train.py:135-150  forward, loss.backward(), gradient all-reduce, clip, step; then the aux loss

Data parallelism: one process per GPU, every rank steps on its own images; the only
collective is the gradient all-reduce (mean) -- 99.86 M fp32 = 399 MB per step for STF.  `GradientAllReduce` does what
DDP's reducer does for this model: gradients are packed into a few flat fp32 buckets, each bucket is all-reduced with
NCCL (NVLS / NVLink) asynchronously in reverse registration order, and unpacked -- launched after backward here (the
backward of this model ends with the largest conv stacks, so there is little left to overlap with).
"""
import math

import torch
import torch.distributed as dist
import torch.nn as nn


class RateDistortionLoss(nn.Module):
    """lambda * 255^2 * MSE + bpp  (train.py:39-59)."""

    def __init__(self, lmbda=1e-2):
        super().__init__()
        self.mse = nn.MSELoss()
        self.lmbda = lmbda

    def forward(self, output, target):
        N, _, H, W = target.size()
        num_pixels = N * H * W
        out = {}
        out["bpp_loss"] = sum(torch.log(lik).sum() / (-math.log(2) * num_pixels)
                              for lik in output["likelihoods"].values())
        out["mse_loss"] = self.mse(output["x_hat"], target)
        out["loss"] = self.lmbda * 255 ** 2 * out["mse_loss"] + out["bpp_loss"]
        return out


def configure_optimizers(net, learning_rate=1e-4, aux_learning_rate=1e-3):
    """Two Adam optimizers, parameters split on the `.quantiles` suffix (train.py:88-120)."""
    named = dict(net.named_parameters())
    main = sorted(n for n, p in named.items() if not n.endswith(".quantiles") and p.requires_grad)
    aux = sorted(n for n, p in named.items() if n.endswith(".quantiles") and p.requires_grad)
    assert not set(main) & set(aux) and len(main) + len(aux) == len(named)
    return (torch.optim.Adam((named[n] for n in main), lr=learning_rate),
            torch.optim.Adam((named[n] for n in aux), lr=aux_learning_rate))


class GradientAllReduce:
    """Bucketed gradient averaging over the data-parallel group (the one collective of the training step)."""

    def __init__(self, params, bucket_mb=100.0, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.buckets, cur, size = [], [], 0
        cap = int(bucket_mb * (1 << 20) / 4)
        for p in reversed(self.params):           # reverse registration order ~ the order gradients become ready
            cur.append(p)
            size += p.numel()
            if size >= cap:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self._flat = None

    def world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def __call__(self):
        """Average .grad over the ranks in place.  Returns the number of bytes all-reduced."""
        w = self.world()
        if w == 1:
            return 0
        if self._flat is None:
            dev = self.params[0].device
            self._flat = [torch.empty(sum(p.numel() for p in b), dtype=torch.float32, device=dev) for b in self.buckets]
        works, nbytes = [], 0
        for flat, bucket in zip(self._flat, self.buckets):
            off = 0
            for p in bucket:
                n = p.numel()
                if p.grad is None:
                    flat[off:off + n].zero_()
                else:
                    flat[off:off + n].copy_(p.grad.reshape(-1))
                off += n
            flat.div_(w)
            works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            nbytes += flat.numel() * 4
        for work, flat, bucket in zip(works, self._flat, self.buckets):
            work.wait()
            off = 0
            for p in bucket:
                n = p.numel()
                if p.grad is None:
                    p.grad = flat[off:off + n].reshape(p.shape).clone()
                else:
                    p.grad.copy_(flat[off:off + n].reshape(p.shape))
                off += n
        return nbytes


def train_step(net, x, criterion, optimizer, aux_optimizer, reducer=None, clip_max_norm=1.0, noise=None):
    """One step of train.py:135-150.  Returns the criterion dict (+ "aux_loss")."""
    optimizer.zero_grad(set_to_none=True)
    aux_optimizer.zero_grad(set_to_none=True)
    out = criterion(net(x) if noise is None else net(x, noise=noise), x)
    out["loss"].backward()
    if reducer is not None:
        reducer()
    if clip_max_norm > 0:
        torch.nn.utils.clip_grad_norm_(net.parameters(), clip_max_norm)
    optimizer.step()
    aux = net.aux_loss()
    aux.backward()
    # (the aux loss depends on the replicated parameters only: its gradient is identical on every rank, no collective)
    aux_optimizer.step()
    out["aux_loss"] = aux
    return out
