"""One stf_swin_mlp launch at the stage-0 shape of a batch (developer tool for ncu captures).  python tools/one_mlp.py [images]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from stf_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ops.set_precision(os.environ.get("PREC", "fp32"))
C, hid = 48, 192
M = B * 384 * 256
x = torch.randn(M, C, device="cuda")
g, be = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda") * 0.1
pc1 = ops.PackedConv(torch.randn(hid, C, device="cuda") / 7, torch.randn(hid, device="cuda") * 0.1, prec=ops.precision_code(), ln=(g, be, 1e-5))
pc2 = ops.PackedConv(torch.randn(C, hid, device="cuda") / 14, torch.randn(C, device="cuda") * 0.1, prec=ops.precision_code())
out = torch.empty_like(x)
for _ in range(2):
    ops.swin_mlp(x, pc1, pc2, out=out)
torch.cuda.synchronize()
print("ok", M)
