"""Fused Swin MLP (stf_swin_mlp) against the two-launch GEMM-engine path at the STF stage shapes (developer tool)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from stf_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for prec in (os.environ.get("PRECS", "fp32,tf32").split(",")):
    ops.set_precision(prec)
    for C, tokens in ((48, 384 * 256), (96, 192 * 128), (192, 96 * 64))[: int(os.environ.get("SHAPES", "3"))]:
        M, hid = B * tokens, 4 * C
        x = torch.randn(M, C, device="cuda")
        g, be = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.1
        w1, b1 = torch.randn(hid, C, device="cuda") / C ** 0.5, torch.randn(hid, device="cuda") * 0.1
        w2, b2 = torch.randn(C, hid, device="cuda") / hid ** 0.5, torch.randn(C, device="cuda") * 0.1
        pc1 = ops.PackedConv(w1, b1, prec=ops.precision_code(), ln=(g, be, 1e-5))
        pc2 = ops.PackedConv(w2, b2, prec=ops.precision_code())
        out = torch.empty_like(x)
        t2 = timed(lambda: ops.gemm(ops.gemm(x, pc1, act="gelu"), pc2, act="residual", residual=x, out=out))
        try:
            t1 = timed(lambda: ops.swin_mlp(x, pc1, pc2, out=out))
        except Exception as e:  # shapes the fused kernel does not take
            print(f"{prec} C={C:3d} M={M}: two launches {t2:.3f} ms | fused: {type(e).__name__} {e}")
            continue
        gb = 8 * M * C / 1e9
        passes = 3 if prec == "fp32" else 1
        tf = 4 * M * C * hid * passes / 1e12
        print(f"{prec} C={C:3d} M={M}: two launches {t2:.3f} ms | fused {t1:.3f} ms = {gb / t1 * 1e3:.0f} GB/s algorithmic, "
              f"{tf / t1 * 1e3:.0f} TFLOP/s-pass")
