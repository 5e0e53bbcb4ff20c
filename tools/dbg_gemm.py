import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stf_b200 import _C, ops
ops.set_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32")
def bench(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
C, M = 48, 786432
g = torch.Generator().manual_seed(0)
x = torch.randn(M, C, generator=g).cuda()
w1 = (torch.randn(4 * C, C, generator=g) / 7).cuda(); b1 = torch.randn(4 * C, generator=g).cuda()
w2 = (torch.randn(C, 4 * C, generator=g) / 14).cuda(); b2 = torch.randn(C, generator=g).cuda()
gam = torch.ones(C).cuda(); bet = torch.zeros(C).cuda()
p1 = ops.PackedConv(w1, b1, prec=ops.precision_code(), ln=(gam, bet, 1e-5))
p1n = ops.PackedConv(w1, b1, prec=ops.precision_code())
p2 = ops.PackedConv(w2, b2, prec=ops.precision_code())
h = ops.gemm(x, p1n, act="gelu")
print("dbg", os.environ.get("STF_B200_CONV_DEBUG", "0"),
      f"fc1 LN+gelu {bench(lambda: ops.gemm(x, p1, act='gelu'))*1e3:.0f} us | fc1 noLN gelu {bench(lambda: ops.gemm(x, p1n, act='gelu'))*1e3:.0f} us | "
      f"fc1 noLN noact {bench(lambda: ops.gemm(x, p1n))*1e3:.0f} us | fc2 res {bench(lambda: ops.gemm(h, p2, act='residual', residual=x))*1e3:.0f} us | fc2 plain {bench(lambda: ops.gemm(h, p2))*1e3:.0f} us")
