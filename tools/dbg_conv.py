import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stf_b200 import ops
def bench(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
row = f"dbg {os.environ.get('STF_B200_CONV_DEBUG','0'):>4s} halo {os.environ.get('STF_B200_CONV_HALO','1')}:"
for B, cin, cout in ((21, 576, 224), (64, 128, 64), (21, 224, 176), (64, 64, 32)):
    x = torch.randn(B, 32, 48, cin, device="cuda"); w = torch.randn(cout, cin, 3, 3, device="cuda") / 50; b = torch.randn(cout, device="cuda")
    pc = ops.PackedConv(w, b, (cin,), prec=0)
    row += f"  B{B} {cin}->{cout}: {bench(lambda: ops.conv2d([x], pc, act=True)):7.1f} us"
print(row)
