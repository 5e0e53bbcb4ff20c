"""Synchronised wall-clock breakdown of STF compress / decompress (developer tool).
   python tools/phase_breakdown.py [--batch 8] [--iters 3]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from stf_b200 import models, ops  # noqa: E402
from stf_b200.synth import synthetic_image  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=3)
    args = ap.parse_args()
    net = models.SymmetricalTransFormer()
    torch.nn.Module.load_state_dict(net, bench.synthetic_weights(), strict=False)
    net = net.cuda().eval()
    net.update(force=True)
    xs = [synthetic_image(args.batch, bench.H, bench.W, seed=i).cuda() for i in range(args.iters + 1)]
    enc = net.compress(xs[0])
    net.decompress(enc["strings"], enc["shape"])
    models.PHASE_TIMES = {}
    l0 = ops.launch_count()
    for x in xs[1:]:
        enc = net.compress(x)
        net.decompress(enc["strings"], enc["shape"])
    tot = 0.0
    for k, v in models.PHASE_TIMES.items():
        print(f"{k:20s} {v / args.iters:8.2f} ms")
        tot += v / args.iters
    print(f"{'sum':20s} {tot:8.2f} ms per step of {args.batch} images; stf_b200 launches/step {(ops.launch_count() - l0) / args.iters:.0f}")
    print("host threads:", os.cpu_count(), " y bytes/image:", sum(len(s) for s in enc["strings"][0]) / args.batch)


if __name__ == "__main__" and not os.environ.get("SPLIT"):
    main()


def graph_split(batch=16, iters=5):
    """GPU-only vs host-only time of one step in CUDA-graph mode (no overlap assumed)."""
    import time
    from stf_b200 import ans
    net = models.SymmetricalTransFormer()
    torch.nn.Module.load_state_dict(net, bench.synthetic_weights(), strict=False)
    net = net.cuda().eval()
    net.update(force=True)
    x = synthetic_image(batch, bench.H, bench.W, seed=3).cuda()
    enc = net.compress(x)
    net.decompress(enc["strings"], enc["shape"])
    torch.cuda.synchronize()

    def gpu_time(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    eplan = list(net._enc_plans.values())[0]
    print(f"encode graph (analysis + hyper + 12 slices), batch {eplan.static_in[0].shape[0]}: {gpu_time(lambda: eplan.graph.replay()):.2f} ms GPU")
    for key, (segs, st) in net._dec_plans.items():
        ts = [gpu_time(lambda s=s: s.graph.replay()) for s in segs]
        print(f"decode plan {key}: seg0 {ts[0]:.2f}  mid avg {sum(ts[1:-1]) / len(ts[1:-1]):.2f} x{len(ts) - 2}  last(+synthesis) {ts[-1]:.2f}  total {sum(ts):.2f} ms GPU")
    # host rANS alone
    gc = net.gaussian_conditional
    tab = gc.rans_table()
    dbg = {}
    net.cuda_graphs = False
    net.compress(x, debug=dbg)
    sym, idx = dbg["symbols"].numpy(), dbg["indexes"].numpy()
    t0 = time.perf_counter()
    for _ in range(iters):
        strings = ans.encode_batch(tab, list(sym), list(idx))
    t1 = time.perf_counter()
    n = sym.shape[1] // 12
    for _ in range(iters):
        decs = models._decoders(strings)
        for i in range(12):
            ans.decode_batch(decs, tab, [idx[b, i * n:(i + 1) * n] for b in range(batch)])
    t2 = time.perf_counter()
    print(f"host rANS: encode {1e3 * (t1 - t0) / iters:.2f} ms, decode (12 slice calls) {1e3 * (t2 - t1) / iters:.2f} ms for {batch} images on {os.cpu_count()} cores")


if __name__ == "__main__" and os.environ.get("SPLIT"):
    graph_split(int(os.environ.get("SPLIT")))
