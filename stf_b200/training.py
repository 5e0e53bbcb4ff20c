"""Rate-distortion training step of the reference's train.py (BASELINE config 5) on the stf_b200 kernels.

  RateDistortionLoss        train.py:39-59
  configure_optimizers      train.py:88-120   (main Adam on everything but *.quantiles, aux Adam on *.quantiles)
  train_step                train.py:135-150  forward, loss.backward(), gradient all-reduce, clip, step; then the aux loss

Data parallelism: one process per GPU, every rank steps on its own images; the only
collective is the gradient all-reduce (mean) -- 99.86 M fp32 = 399 MB per step for STF.  `GradientAllReduce` does what
DDP's reducer does for this model: gradients are packed into a few flat fp32 buckets, each bucket is all-reduced with
NCCL (NVLS / NVLink) asynchronously in reverse registration order -- from post-accumulate hooks inside backward, so the
transfers overlap the remaining backward work -- and unpacked before the optimizer step.
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
    """Bucketed gradient averaging over the data-parallel group (the one collective of the training step).

    `attach()` registers post-accumulate hooks: as soon as the last gradient of a bucket has been produced by backward,
    its all-reduce is launched asynchronously, so the transfers of the early buckets (the synthesis transform and the conv
    stacks, whose gradients come first) run under the rest of the backward pass.  `__call__()` after backward launches
    whatever has not been launched and waits.
    Gradients live IN the buckets (`grad_views=True`, default): every parameter's .grad is a view (with the parameter's own
    strides) into its bucket's flat buffer, autograd accumulates into it in place, and NCCL reduces the flat buffer where it
    lies -- no pack / unpack copies.  Use `zero_grad()` of this object instead of optimizer.zero_grad(set_to_none=True)."""

    def __init__(self, params, bucket_mb=50.0, group=None, grad_views=True):
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
        self._bucket_of = {id(p): k for k, b in enumerate(self.buckets) for p in b}
        self._pending = [len(b) for b in self.buckets]
        self._works = [None] * len(self.buckets)
        self._hooks = []
        self.grad_views = bool(grad_views)

    def zero_grad(self):
        """Zero the gradients in place (one fill per bucket) and (re)attach every .grad to its bucket view."""
        if not self.grad_views:
            for p in self.params:
                p.grad = None
            return
        self._ensure_flat()
        torch._foreach_zero_(self._flat)
        for bucket, views, flat in zip(self.buckets, self._views, self._flat):
            off = 0
            for j, p in enumerate(bucket):
                if views[j].stride() != p.stride():      # the parameter was re-laid-out (channels_last) after the views were made
                    views[j] = self._view(flat[off:off + p.numel()], p)
                if p.grad is None or p.grad.data_ptr() != views[j].data_ptr() or p.grad.stride() != views[j].stride():
                    p.grad = views[j]
                off += p.numel()

    def world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def attach(self):
        """Overlap mode: launch each bucket's all-reduce from inside backward."""
        if self._hooks or self.world() == 1:
            return self
        for p in self.params:
            self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        return self

    def arm(self):
        """Call right before the backward pass whose gradients are to be averaged (other backward passes, e.g. the
        aux loss, leave the counters alone)."""
        self._pending = [len(b) for b in self.buckets]
        self._works = [None] * len(self.buckets)
        self._armed = True

    def _on_grad(self, p):
        if not getattr(self, "_armed", False):
            return
        k = self._bucket_of[id(p)]
        self._pending[k] -= 1
        if self._pending[k] == 0:
            self._launch(k)

    @staticmethod
    def _view(seg, p):
        """View of a flat bucket segment with the parameter's own strides (conv weights are channels_last): autograd
        accumulates into it without a layout copy."""
        dense = p.numel() == 1 + sum((n - 1) * st for n, st in zip(p.shape, p.stride()))
        return seg.as_strided(p.shape, p.stride()) if dense and not p.is_contiguous() else seg.view_as(p)

    def _ensure_flat(self):
        if self._flat is None:
            dev = self.params[0].device
            self._flat = [torch.empty(sum(p.numel() for p in b), dtype=torch.float32, device=dev) for b in self.buckets]
            self._views = []
            for flat, bucket in zip(self._flat, self.buckets):
                views, off = [], 0
                for p in bucket:
                    views.append(self._view(flat[off:off + p.numel()], p))
                    off += p.numel()
                self._views.append(views)

    def _launch(self, k):
        # pack with ONE multi-tensor copy per bucket (743 parameters would otherwise be ~1500 tiny launches per step)
        self._ensure_flat()
        bucket, views, flat = self.buckets[k], self._views[k], self._flat[k]
        stray = [(v, p.grad) for p, v in zip(bucket, views) if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
        if stray:      # gradients that do not live in the bucket (grad_views off, or re-created by the caller): pack them
            torch._foreach_copy_([v for v, _ in stray], [g for _, g in stray])
        for p, v in zip(bucket, views):
            if p.grad is None:        # no gradient this step (or the caller dropped it): contributes zeros
                v.zero_()
                p.grad = v
        flat.div_(self.world())
        self._works[k] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def __call__(self):
        """Average .grad over the ranks in place.  Returns the number of bytes all-reduced."""
        if self.world() == 1:
            return 0
        nbytes = 0
        for k in range(len(self.buckets)):
            if self._works[k] is None:              # parameters without a gradient this step, or no hooks attached
                self._launch(k)
        for k, bucket in enumerate(self.buckets):
            self._works[k].wait()
            stray = [(p.grad, v) for p, v in zip(bucket, self._views[k]) if p.grad.data_ptr() != v.data_ptr()]
            if stray:
                torch._foreach_copy_([g for g, _ in stray], [v for _, v in stray])
            nbytes += self._flat[k].numel() * 4
        self._armed = False
        self._works = [None] * len(self.buckets)
        return nbytes


def train_step(net, x, criterion, optimizer, aux_optimizer, reducer=None, clip_max_norm=1.0, noise=None):
    """One step of train.py:135-150.  Returns the criterion dict (+ "aux_loss").  `net` may be the bare model (with an
    optional GradientAllReduce `reducer`) or the model wrapped in torch DistributedDataParallel, as in train.py:363 (the
    custom autograd functions produce ordinary .grad tensors, so DDP's bucketed, overlapped all-reduce applies)."""
    if reducer is not None and reducer.grad_views:
        reducer.zero_grad()               # gradients stay attached to the all-reduce buckets (no pack / unpack copies)
        aux_optimizer.zero_grad(set_to_none=False)
    else:
        optimizer.zero_grad(set_to_none=True)
        aux_optimizer.zero_grad(set_to_none=True)
    out = criterion(net(x) if noise is None else net(x, noise=noise), x)
    if reducer is not None:
        reducer.arm()
    out["loss"].backward()
    if reducer is not None:
        reducer()
    if clip_max_norm > 0:
        torch.nn.utils.clip_grad_norm_(net.parameters(), clip_max_norm)
    optimizer.step()
    aux = (net.module if hasattr(net, "module") else net).aux_loss()
    aux.backward()
    # (the aux loss depends on the replicated parameters only: its gradient is identical on every rank, no collective)
    aux_optimizer.step()
    out["aux_loss"] = aux
    return out
