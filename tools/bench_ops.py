"""Per-op microbenchmark (developer tool): CUDA-event time and achieved algorithmic GB/s / TFLOP/s of the
stf_b200 kernels at the STF 768x512 stage shapes.   python tools/bench_ops.py [--batch 8] [--iters 5] [--only linear]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from stf_b200 import _C, ops  # noqa: E402


def timeit(fn, iters, flush):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()                       # > L2 (126 MB): evict the previous iteration's data
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--stages", default="0,1,2,3")
    args = ap.parse_args()
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    B, H0, W0 = args.batch, 256, 384
    rows = []
    for st in [int(s) for s in args.stages.split(",")]:
        C, nh = 48 << st, 3 << st
        H, W = H0 >> st, W0 >> st
        T = B * H * W
        x = torch.randn(T, C, device=dev)
        g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        Wqkv = ops.PackedLinear(torch.randn(3 * C, C, device=dev) * C ** -0.5, torch.zeros(3 * C, device=dev), (g, b, 1e-5))
        Wproj = ops.PackedLinear(torch.randn(C, C, device=dev) * C ** -0.5, torch.zeros(C, device=dev))
        Wfc1 = ops.PackedLinear(torch.randn(4 * C, C, device=dev) * C ** -0.5, torch.zeros(4 * C, device=dev), (g, b, 1e-5))
        Wfc2 = ops.PackedLinear(torch.randn(C, 4 * C, device=dev) * (4 * C) ** -0.5, torch.zeros(C, device=dev))
        table = torch.randn(49, nh, device=dev)
        geom = (B, H, W, 4, 2)
        qkv = torch.empty(T, 3 * C, device=dev)
        o = torch.empty(T, C, device=dev)
        x1 = torch.empty(T, C, device=dev)
        h = torch.empty(T, 4 * C, device=dev)
        x2 = torch.empty(T, C, device=dev)
        cases = {
            "linear qkv (LN, window, shift)": (lambda: ops.linear(x, Wqkv, rows=_C.ROWS_WINDOW, epilogue=_C.EPI_QKV, q_cols=C, q_scale=0.25, geom=geom, out=qkv),
                                               4 * T * (C + 3 * C), 2 * T * C * 3 * C),
            "attention core (shift)": (lambda: ops.window_attention_core(qkv, table, T // 16, C, nh, 4, 2, H, W),
                                       4 * T * 4 * C, 4 * T * 16 * C),
            "linear proj (+window residual)": (lambda: ops.linear(o, Wproj, epilogue=_C.EPI_WINDOW_RESIDUAL, residual=x, geom=geom, out=x1, x_is_tf32=True),
                                               4 * T * 3 * C, 2 * T * C * C),
            "linear fc1 (LN, GELU)": (lambda: ops.linear(x1, Wfc1, epilogue=_C.EPI_GELU, out=h),
                                      4 * T * 5 * C, 2 * T * C * 4 * C),
            "linear fc2 (+residual)": (lambda: ops.linear(h, Wfc2, epilogue=_C.EPI_RESIDUAL, residual=x1, out=x2, x_is_tf32=True),
                                       4 * T * 6 * C, 2 * T * C * 4 * C),
        }
        for name, (fn, nbytes, flops) in cases.items():
            if args.only and args.only not in name:
                continue
            ms = timeit(fn, args.iters, flush)
            rows.append((st, C, T, name, ms, nbytes / ms / 1e6, flops / ms / 1e9))
    print(f"GEMM precision: {ops.precision()}")
    print(f"{'stage':>5} {'C':>4} {'tokens':>8}  {'op':34} {'ms':>8} {'GB/s':>8} {'TFLOP/s':>8}")
    tot = 0
    for st, C, T, name, ms, gbs, tf in rows:
        tot += ms
        print(f"{st:5d} {C:4d} {T:8d}  {name:34} {ms:8.3f} {gbs:8.0f} {tf:8.1f}")
    print(f"sum of medians: {tot:.3f} ms for one Swin block per stage at batch {B}")
    if not args.only or "entropy" in args.only:
        entropy(args, flush)


def entropy(args, flush):
    """Entropy kernels at the roofline size of SURVEY.md 8d (2^26 elements: 256 MB per fp32 tensor >> 126 MB L2)
    and at a natural slice size (batch 32 x 32 channels x 32 x 48)."""
    import json
    import math
    peak = 6453.4
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    table = torch.exp(torch.linspace(math.log(0.11), math.log(256.0), 64))    # get_scale_table() (stf.py:21-31)
    dev = "cuda"
    print(f"{'elements':>10}  {'entropy kernel':44} {'ms':>8} {'GB/s':>8} {'of HBM peak':>11}")
    for B, Cs, h, w in ((2048, 32, 32, 32), (32, 32, 32, 48)):
        n = B * Cs * h * w
        g = torch.Generator(device=dev).manual_seed(0)
        scales = torch.exp(torch.empty(B, Cs, h, w, device=dev).uniform_(math.log(0.01), math.log(400.0), generator=g))
        means = 2 * torch.randn(B, Cs, h, w, device=dev, generator=g)
        y = means + scales * torch.randn(B, Cs, h, w, device=dev, generator=g)
        sym = torch.empty(B, Cs * h * w, dtype=torch.int32, device=dev)
        idx = torch.empty_like(sym)
        C_eb = 192
        z = 3 * torch.randn(max(B * Cs // C_eb, 1), C_eb, h, w, device=dev, generator=g)
        from stf_b200.entropy_models import EntropyBottleneck
        eb = EntropyBottleneck(C_eb).to(dev)
        params = eb.packed_params()
        cases = [
            ("gaussian_likelihood (y_hat + lik, 20 B/el)", lambda: ops.gaussian_likelihood(y, 0, scales, means, ste_round=True), 20 * n),
            ("gaussian_likelihood (lik only, 16 B/el)", lambda: ops.gaussian_likelihood(y, 0, scales, means, want_y_hat=False), 16 * n),
            ("compress_step (idx + sym + y_hat, 24 B/el)", lambda: ops.gaussian_compress_step(y, 0, scales, means, table, sym, idx, 0), 24 * n),
            ("build_indexes (8 B/el)", lambda: ops.build_indexes(scales, table), 8 * n),
            ("dequantize (12 B/el)", lambda: ops.dequantize(sym, 0, means), 12 * n),
            ("quantize symbols (12 B/el)", lambda: ops.quantize_symbols(y, means), 12 * n),
        ]
        if params is not None:
            cases.append(("entropy_bottleneck (z_hat + lik, 12 B/el)", lambda: ops.entropy_bottleneck(z, params), 12 * z.numel()))
        for name, fn, nbytes in cases:
            ms = timeit(fn, args.iters, flush)
            print(f"{n:10d}  {name:44} {ms:8.3f} {nbytes / ms / 1e6:8.0f} {nbytes / ms / 1e6 / peak:11.2f}")


if __name__ == "__main__":
    main()
