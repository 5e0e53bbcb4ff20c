"""One eager (no CUDA graphs) compress+decompress step for ncu captures.  python tools/one_step.py [batch]"""
import os
import sys

os.environ["STF_B200_CUDA_GRAPHS"] = "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from stf_b200 import models  # noqa: E402
from stf_b200.synth import synthetic_image  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
net = models.SymmetricalTransFormer()
torch.nn.Module.load_state_dict(net, bench.synthetic_weights(), strict=False)
net = net.cuda().eval()
net.update(force=True)
x = synthetic_image(B, bench.H, bench.W, seed=1).cuda()
enc = net.compress(x)
dec = net.decompress(enc["strings"], enc["shape"])
torch.cuda.synchronize()
print("ok", B, sum(len(s) for s in enc["strings"][0]))
