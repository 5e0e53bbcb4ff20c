"""Wall-clock of compress() and decompress() separately in the product configuration (CUDA graphs, pipelined halves),
plus the host rANS time inside each (developer tool).   python tools/split_times.py [--batch 32]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from stf_b200 import ans, models  # noqa: E402
from stf_b200.synth import synthetic_image  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    net = models.SymmetricalTransFormer()
    torch.nn.Module.load_state_dict(net, bench.synthetic_weights(), strict=False)
    net = net.cuda().eval()
    net.update(force=True)
    xs = [synthetic_image(args.batch, bench.H, bench.W, seed=i).cuda() for i in range(args.iters + 2)]
    for x in xs[:2]:
        enc = net.compress(x)
        net.decompress(enc["strings"], enc["shape"])
    host = {"enc": 0.0, "dec": 0.0}
    orig_e, orig_d = ans.encode_batch, ans.decode_batch
    mode = ["enc"]

    def timed(fn, key):
        def w(*a, **k):
            t0 = time.perf_counter()
            r = fn(*a, **k)
            host[key] += time.perf_counter() - t0
            return r
        return w
    ans.encode_batch, ans.decode_batch = timed(orig_e, "enc"), timed(orig_d, "dec")
    models.ans = ans
    tc = td = 0.0
    for x in xs[2:]:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        enc = net.compress(x)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        net.decompress(enc["strings"], enc["shape"])
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        tc += t1 - t0
        td += t2 - t1
    n = args.iters
    print(f"batch {args.batch}: compress {tc / n * 1e3:.1f} ms (host rANS inside: {host['enc'] / n * 1e3:.1f} ms), "
          f"decompress {td / n * 1e3:.1f} ms (host rANS inside: {host['dec'] / n * 1e3:.1f} ms), threads {os.cpu_count()}")


if __name__ == "__main__":
    main()
