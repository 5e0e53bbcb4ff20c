"""Bring-up of stf_conv2d on the GPU box: cases one by one with prints (a hang names its case), then timing vs cuDNN."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_conv import _case

cases = [
    dict(B=1, H=8, W=16, chans=(32,), N=16, k=1),
    dict(B=1, H=8, W=16, chans=(32,), N=16, k=3),
    dict(B=1, H=32, W=48, chans=(64,), N=32, k=3),
    dict(B=2, H=32, W=48, chans=(176,), N=128, k=3, act=True),
    dict(B=3, H=32, W=48, chans=(384, 96, 32), N=224, k=3, act=True),
    dict(B=2, H=32, W=48, chans=(336,), N=288, k=3, stride=2, act=True),
    dict(B=2, H=8, W=12, chans=(240,), N=1152, k=3, shuffle=2, act=True),
    dict(B=1, H=64, W=96, chans=(48,), N=192, k=5, shuffle=2),
    dict(B=2, H=30, W=44, chans=(64, 32), N=64, k=3, act=True),
]
for prec in ("tf32", "fp32"):
    for c in cases:
        print(prec, c, flush=True)
        for exact in (True, False):
            try:
                _, err = _case(prec=prec, exact_inputs=exact, **c)
                torch.cuda.synchronize()
                print(f"   exact_inputs={exact}: rel err {err:.3e}", flush=True)
            except Exception as e:
                print("   FAILED:", type(e).__name__, e, flush=True)
                if "CUDA" in str(e) or "launch" in str(e):
                    sys.exit(1)

# timing vs cuDNN (TF32) on the slice-loop shapes at batch 21 (one of three sub-batches of 64)
from stf_b200 import ops
torch.backends.cudnn.allow_tf32 = True
def bench(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for B in (1, 21, 64):
    for (cin, cout) in ((576, 224), (224, 176), (176, 128), (128, 64), (64, 32), (384, 384)):
        x = torch.randn(B, 32, 48, cin, device="cuda")
        w = torch.randn(cout, cin, 3, 3, device="cuda") / 50
        b = torch.randn(cout, device="cuda")
        xc = x.permute(0, 3, 1, 2)           # channels_last view
        wc = w.contiguous(memory_format=torch.channels_last)
        flops = 2 * B * 32 * 48 * cin * cout * 9
        t_cudnn = bench(lambda: F.gelu(F.conv2d(xc, wc, b, padding=1)))
        t_cudnn_conv = bench(lambda: F.conv2d(xc, wc, None, padding=1))
        row = f"B={B:3d} {cin:4d}->{cout:4d}: cuDNN conv+bias+gelu {t_cudnn*1e3:7.1f} us (conv only {t_cudnn_conv*1e3:7.1f} us = {flops/t_cudnn_conv/1e9:6.1f} TFLOP/s)"
        for prec in ("tf32", "fp32"):
            pc = ops.PackedConv(w, b, (cin,), prec=ops._PRECISIONS[prec])
            t = bench(lambda: ops.conv2d([x], pc, act=True))
            row += f" | ours {prec} {t*1e3:7.1f} us = {flops/t/1e9:6.1f} TFLOP/s"
        print(row, flush=True)
