"""Dump the y symbols / indexes of two images of the bench workload (developer tool, for host-coder tuning on CPU)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch  # noqa: E402

import bench  # noqa: E402
from stf_b200 import models  # noqa: E402
from stf_b200.synth import synthetic_image  # noqa: E402

net = models.SymmetricalTransFormer()
torch.nn.Module.load_state_dict(net, bench.synthetic_weights(), strict=False)
net = net.cuda().eval()
net.update(force=True)
x = synthetic_image(4, bench.H, bench.W, seed=1).cuda()
enc = net.compress(x)
sym_h, idx_h = net._pinned[[k for k in net._pinned if k[0] == ("y", 0)][0]]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "bench_symbols.npz"), sym=sym_h[:2].numpy(), idx=idx_h[:2].numpy(),
                    lens=np.array([len(s) for s in enc["strings"][0]]))
print("ok", sym_h.shape, [len(s) for s in enc["strings"][0]])
