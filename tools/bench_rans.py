"""Host rANS micro-benchmark (developer tool, CPU only): ns per symbol of encode / decode for K images of n symbols on T threads.
   python tools/bench_rans.py [--images 32] [--threads 16] [--n 589824]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stf_b200 import ans  # noqa: E402
from stf_b200.entropy_models import GaussianConditional  # noqa: E402
from stf_b200.models import get_scale_table  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=32)
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--n", type=int, default=384 * 32 * 48)
    ap.add_argument("--iters", type=int, default=3)
    args = ap.parse_args()
    gc = GaussianConditional(None)
    gc.update_scale_table(get_scale_table())
    tab = gc.rans_table()
    table = get_scale_table().numpy()
    syms, idxs = [], []
    for i in range(args.images):
        rng = np.random.default_rng(i)
        ix = np.minimum(63, np.abs(rng.normal(20, 12, size=args.n)).astype(np.int32))     # mid-table scales dominate
        syms.append(np.rint(rng.standard_normal(args.n) * table[ix]).astype(np.int32))
        idxs.append(ix)
    best_e = best_d = 1e9
    for _ in range(args.iters):
        t0 = time.perf_counter()
        strings = ans.encode_batch(tab, syms, idxs, threads=args.threads)
        best_e = min(best_e, time.perf_counter() - t0)
        decs = []
        for s in strings:
            d = ans.RansDecoder()
            d.set_stream(s)
            decs.append(d)
        outs = [np.empty(args.n, dtype=np.int32) for _ in range(args.images)]
        t0 = time.perf_counter()
        ans.decode_batch(decs, tab, idxs, outs=outs, threads=args.threads)
        best_d = min(best_d, time.perf_counter() - t0)
        assert all(np.array_equal(o, s) for o, s in zip(outs, syms))
    per_thread = -(-args.images // args.threads)
    print(f"{args.images} images x {args.n} symbols on {args.threads} threads ({per_thread} per thread): "
          f"encode {best_e * 1e3:.1f} ms ({best_e / per_thread / args.n * 1e9:.1f} ns/symbol/thread), "
          f"decode {best_d * 1e3:.1f} ms ({best_d / per_thread / args.n * 1e9:.1f} ns/symbol/thread), "
          f"{sum(len(s) for s in strings) / args.images / 1e3:.0f} KB per image")


if __name__ == "__main__":
    main()
