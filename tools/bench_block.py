"""Per-op timing of one Swin block per STF stage on the round-2 path (batch 8 of 768x512 by default, L2 flushed between runs):
qkv GEMM (LN1 folded) -> token-order tensor-core window attention -> proj GEMM (+shortcut) -> fc1 GEMM (LN2, GELU) -> fc2 GEMM
(+shortcut).  Comparable with profiles/r1_ops_microbench_v4_fp32.txt (round-1 kernels, same shapes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stf_b200 import _C, ops

def bench(fn, n=10):
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for prec in (sys.argv[2:] or ["fp32", "tf32"]):
    ops.set_precision(prec)
    print(f"GEMM precision: {prec}   (batch {B} of 768x512)")
    print("stage    C   tokens  op                                        ms     GB/s  TFLOP/s")
    total = 0.0
    for stage, (C, heads) in enumerate(((48, 3), (96, 6), (192, 12), (384, 24))):
        H, W = 256 >> stage, 384 >> stage
        M = B * H * W
        g = torch.Generator().manual_seed(stage)
        x = (torch.randn(M, C, generator=g)).cuda()
        r = lambda *s: torch.randn(*s, generator=g).cuda()
        ln = (1 + 0.1 * r(C), 0.1 * r(C), 1e-5)
        pq = ops.PackedConv(r(3 * C, C) / C ** 0.5, r(3 * C), prec=ops.precision_code(), ln=ln, row_scale=(C, 0.25))
        pp = ops.PackedConv(r(C, C) / C ** 0.5, r(C), prec=ops.precision_code())
        p1 = ops.PackedConv(r(4 * C, C) / C ** 0.5, r(4 * C), prec=ops.precision_code(), ln=ln)
        p2 = ops.PackedConv(r(C, 4 * C) / (4 * C) ** 0.5, r(C), prec=ops.precision_code())
        table = (0.02 * r(49, heads))
        qkv = ops.gemm(x, pq)
        o = ops.window_attention_tokens(qkv, table, None, B, H, W, C, heads, 4, 2)
        h = ops.gemm(x, p1, act="gelu")
        rows = [
            ("gemm qkv (LN folded, q scaled)", lambda: ops.gemm(x, pq), 4 * M * 4 * C, 2 * M * 3 * C * C),
            ("attention (tokens, shift, mma.sync)", lambda: ops.window_attention_tokens(qkv, table, None, B, H, W, C, heads, 4, 2), 4 * M * 4 * C, 4 * M * 16 * C),
            ("gemm proj (+shortcut)", lambda: ops.gemm(o, pp, act="residual", residual=x), 4 * M * 3 * C, 2 * M * C * C),
            ("gemm fc1 (LN folded, GELU)", lambda: ops.gemm(x, p1, act="gelu"), 4 * M * 5 * C, 2 * M * 4 * C * C),
            ("gemm fc2 (+shortcut)", lambda: ops.gemm(h, p2, act="residual", residual=x), 4 * M * 6 * C, 2 * M * 4 * C * C),
        ]
        for name, fn, nbytes, flops in rows:
            t = bench(fn)
            total += t
            print(f"{stage:5d} {C:4d} {M:8d}  {name:38s} {t:7.3f} {nbytes / t / 1e6:8.0f} {flops / t / 1e9:8.1f}", flush=True)
    print(f"sum of medians: {total:.3f} ms for one Swin block per stage at batch {B}")
