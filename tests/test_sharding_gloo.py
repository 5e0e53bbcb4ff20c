"""CPU, world_size 2 on gloo: the N>1 host path -- contiguous batch sharding, per-rank host rANS coding,
gather of the byte strings, max-over-ranks timing -- reproduces the single-process result byte for byte."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stf_b200.sharding import shard_range


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _images(n):
    """Synthetic (symbols, indexes) per image, deterministic."""
    from oracle import entropy as OE
    table = OE.scale_table().numpy()
    out = []
    for i in range(n):
        rng = np.random.default_rng(100 + i)
        m = 3000 + 517 * i
        ix = rng.integers(0, 64, size=m).astype(np.int32)
        out.append((np.rint(rng.standard_normal(m) * table[ix] * 1.5).astype(np.int32), ix))
    return out


def _worker(rank, world, port, n_images, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import entropy as OE
        from stf_b200 import ans
        from stf_b200.sharding import gather_strings, max_over_ranks
        cdf, lens, offs = OE.gaussian_tables()
        tab = ans.RansTable(cdf, lens, offs)
        lo, hi = shard_range(n_images, rank, world)
        mine = _images(n_images)[lo:hi]
        local = ans.encode_batch(tab, [s for s, _ in mine], [i for _, i in mine], threads=2)
        everything = gather_strings(local)
        slowest = max_over_ranks(10.0 + rank)
        if rank == 0:
            q.put((everything, slowest))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [4, 5])
def test_two_rank_sharding_matches_single_process(n_images):
    from oracle import entropy as OE
    from stf_b200 import ans
    cdf, lens, offs = OE.gaussian_tables()
    tab = ans.RansTable(cdf, lens, offs)
    single = [ans.encode_array(tab, s, i) for s, i in _images(n_images)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_images, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, slowest = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert gathered == single                     # per-image strings identical to the 1-process run, in image order
    assert slowest == 11.0


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 64):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _train_worker(rank, world, port, q, views=False):
    """Data-parallel training step on gloo: each rank has its own half batch; after GradientAllReduce every rank must hold
    the gradient of the full-batch mean loss, and the optimizer steps must keep the replicas identical."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from stf_b200.training import GradientAllReduce
        torch.manual_seed(0)                                   # identical replicas
        net = torch.nn.Sequential(torch.nn.Linear(12, 32), torch.nn.GELU(), torch.nn.Linear(32, 3))
        unused = torch.nn.Parameter(torch.ones(5))             # a parameter that never receives a gradient
        params = list(net.parameters()) + [unused]
        g = torch.Generator().manual_seed(1)
        x, y = torch.randn(8, 12, generator=g), torch.randn(8, 3, generator=g)
        opt = torch.optim.Adam(net.parameters(), lr=1e-2)
        red = GradientAllReduce(params, bucket_mb=0.0005)      # ~130 floats per bucket: several buckets
        red.attach()                                           # overlap mode: buckets launched from backward hooks
        assert len(red.buckets) >= 2
        lo, hi = rank * 4, rank * 4 + 4
        for _ in range(2):
            if views:
                red.zero_grad()        # gradients live in the all-reduce buckets: zeroed in place, reduced where they lie
                assert all(p.grad.data_ptr() == v.data_ptr() for b, vs in zip(red.buckets, red._views) for p, v in zip(b, vs))
            else:
                opt.zero_grad()        # gradients re-created by autograd each step: packed into the buckets
            red.arm()
            torch.nn.functional.mse_loss(net(x[lo:hi]), y[lo:hi]).backward()
            nbytes = red()
            (net[0].weight.sum() * 0.0).backward()              # an unrelated backward (cf. the aux loss) must not disturb it
            opt.step()
        # (numpy, not tensors: tensors travel through the queue as shared-memory handles that die with the worker)
        grads = [p.grad.numpy().copy() for p in net.parameters()]
        q.put((rank, nbytes, [p.detach().numpy().copy() for p in net.parameters()], grads, unused.grad.numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("views", [False, True])
def test_two_rank_gradient_allreduce_equals_full_batch(views):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q, views)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=120) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference: full batch of 8, mean loss == mean of the two half-batch mean losses
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(12, 32), torch.nn.GELU(), torch.nn.Linear(32, 3))
    g = torch.Generator().manual_seed(1)
    x, y = torch.randn(8, 12, generator=g), torch.randn(8, 3, generator=g)
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)
    for _ in range(2):
        opt.zero_grad()
        torch.nn.functional.mse_loss(net(x), y).backward()
        opt.step()
    n_param_bytes = 4 * (sum(p.numel() for p in net.parameters()) + 5)
    for rank, nbytes, params, grads, unused_grad in got:
        assert nbytes == n_param_bytes
        for p, ref in zip(params, net.parameters()):
            assert np.allclose(p, ref.detach().numpy(), rtol=1e-5, atol=1e-6)
        for gr, ref in zip(grads, net.parameters()):
            assert np.allclose(gr, ref.grad.numpy(), rtol=1e-4, atol=1e-6)
        assert np.array_equal(unused_grad, np.zeros(5, dtype=np.float32))
    assert all(np.array_equal(a, b) for a, b in zip(got[0][2], got[1][2]))     # replicas stay bit-identical
