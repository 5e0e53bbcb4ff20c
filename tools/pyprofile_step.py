"""cProfile of one compress+decompress step in CUDA-graph mode (developer tool): where the host time goes."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from stf_b200 import models  # noqa: E402
from stf_b200.synth import synthetic_image  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
net = models.SymmetricalTransFormer()
torch.nn.Module.load_state_dict(net, bench.synthetic_weights(), strict=False)
net = net.cuda().eval()
net.update(force=True)
xs = [synthetic_image(B, bench.H, bench.W, seed=i).cuda() for i in range(4)]
for x in xs[:2]:
    enc = net.compress(x)
    net.decompress(enc["strings"], enc["shape"])
torch.cuda.synchronize()
t0 = time.perf_counter()
enc = net.compress(xs[2])
t1 = time.perf_counter()
net.decompress(enc["strings"], enc["shape"])
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"compress {1e3 * (t1 - t0):.1f} ms, decompress {1e3 * (t2 - t1):.1f} ms")
pr = cProfile.Profile()
pr.enable()
enc = net.compress(xs[3])
net.decompress(enc["strings"], enc["shape"])
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
