"""Wall time of compress() / decompress() for one sub-batch split (developer tool; the split comes from the environment:
STF_B200_ENC_SPLIT / STF_B200_DEC_SPLIT / STF_B200_DEC_LEAD).  Prints one line."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from stf_b200 import models  # noqa: E402
from stf_b200.synth import synthetic_image  # noqa: E402

B, steps = 64, 4
net = models.SymmetricalTransFormer()
torch.nn.Module.load_state_dict(net, bench.synthetic_weights(), strict=False)
net = net.cuda().eval()
net.update(force=True)
xs = [synthetic_image(B, bench.H, bench.W, seed=i).cuda() for i in range(steps + 2)]
tc, td = [], []
for i, x in enumerate(xs):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    enc = net.compress(x)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    net.decompress(enc["strings"], enc["shape"])
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    if i >= 2:
        tc.append((t1 - t0) * 1e3), td.append((t2 - t1) * 1e3)
tc.sort(), td.sort()
print(f"cores {len(os.sched_getaffinity(0)):2d} enc {os.environ.get('STF_B200_ENC_SPLIT', 'default'):12s} dec {os.environ.get('STF_B200_DEC_SPLIT', 'default'):12s} "
      f"lead {os.environ.get('STF_B200_DEC_LEAD', '3')}: compress {tc[len(tc) // 2]:6.1f} ms  decompress {td[len(td) // 2]:6.1f} ms")
