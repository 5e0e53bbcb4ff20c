"""GEMM engine (stf_conv2d, ksize 1) vs stf_linear on the Mlp shapes of the four STF stages (batch 8 of 768x512), both
precision modes; checks the results against each other and against float64."""
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
for prec in ("fp32", "tf32"):
    ops.set_precision(prec)
    print(f"== precision {prec}")
    for stage, C in enumerate((48, 96, 192, 384)):
        M = B * 98304 // 4 ** stage
        g = torch.Generator().manual_seed(stage)
        x = (torch.randn(M, C, generator=g) * 2 + 0.5).cuda()
        w1 = (torch.randn(4 * C, C, generator=g) / C ** 0.5).cuda(); b1 = torch.randn(4 * C, generator=g).cuda()
        w2 = (torch.randn(C, 4 * C, generator=g) / (4 * C) ** 0.5).cuda(); b2 = torch.randn(C, generator=g).cuda()
        gam = (1 + 0.1 * torch.randn(C, generator=g)).cuda(); bet = (0.1 * torch.randn(C, generator=g)).cuda()
        ln = (gam, bet, 1e-5)
        p1 = ops.PackedConv(w1, b1, prec=ops.precision_code(), ln=ln); p2 = ops.PackedConv(w2, b2, prec=ops.precision_code())
        l1 = ops.PackedLinear(w1, b1, ln); l2 = ops.PackedLinear(w2, b2)
        h_new = ops.gemm(x, p1, act="gelu"); y_new = ops.gemm(h_new, p2, act="residual", residual=x)
        h_old = ops.linear(x, l1, epilogue=_C.EPI_GELU); y_old = ops.linear(h_old, l2, epilogue=_C.EPI_RESIDUAL, residual=x, x_is_tf32=True)
        n = min(M, 4096)
        xd = x[:n].double()
        ref_h = torch.nn.functional.gelu(torch.nn.functional.layer_norm(xd, (C,), gam.double(), bet.double(), 1e-5) @ w1.double().t() + b1.double())
        ref_y = xd + ref_h @ w2.double().t() + b2.double()
        e = lambda a, r: ((a[:n].double() - r).abs().max() / r.abs().max()).item()
        t1n, t2n = bench(lambda: ops.gemm(x, p1, act="gelu")), bench(lambda: ops.gemm(h_new, p2, act="residual", residual=x))
        t1o, t2o = bench(lambda: ops.linear(x, l1, epilogue=_C.EPI_GELU)), bench(lambda: ops.linear(h_old, l2, epilogue=_C.EPI_RESIDUAL, residual=x, x_is_tf32=True))
        by1, by2 = 4 * M * 5 * C, 4 * M * 6 * C
        fl = 2 * M * 4 * C * C
        print(f"stage {stage} C={C:3d} M={M:7d}: fc1 engine {t1n*1e3:7.1f} us ({by1/t1n/1e6:6.0f} GB/s, {fl/t1n/1e9:5.0f} TF/s) vs linear {t1o*1e3:7.1f} us | "
              f"fc2 engine {t2n*1e3:7.1f} us ({by2/t2n/1e6:6.0f} GB/s, {fl/t2n/1e9:5.0f} TF/s) vs linear {t2o*1e3:7.1f} us | "
              f"err h {e(h_new, ref_h):.1e}/{e(h_old, ref_h):.1e} y {e(y_new, ref_y):.1e}/{e(y_old, ref_y):.1e}", flush=True)
