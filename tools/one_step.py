"""One eager (no CUDA graphs) compress+decompress step for ncu captures.  python tools/one_step.py [batch]
Writes gpurun_out/one_step_families.json: the roofline family ("kernel|bound", as bench.py classes it) of every stf_b200
launch in launch order, so that tools/ncu_summary.py can attribute the ncu launch list of the same command to families."""
import json
import os
import sys

os.environ["STF_B200_CUDA_GRAPHS"] = "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from stf_b200 import models, ops, profiler  # noqa: E402
from stf_b200.synth import synthetic_image  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
net = models.SymmetricalTransFormer()
torch.nn.Module.load_state_dict(net, bench.synthetic_weights(), strict=False)
net = net.cuda().eval()
net.update(force=True)
x = synthetic_image(B, bench.H, bench.W, seed=1).cuda()
peaks = bench.load_peaks()
passes = 3 if ops.precision() == "fp32" else 1
ridge = peaks["tf32_tflops"] * 1e12 / passes / (peaks["hbm_gbs"] * 1e9)
enc = net.compress(x)                       # warm step: weight packing, table set-up, cuDNN plans
net.decompress(enc["strings"], enc["shape"])
torch.cuda.synchronize()
torch.cuda.profiler.start()                 # ncu --profile-from-start off: only the second step is profiled
with profiler.capture() as prof:
    enc = net.compress(x)
    dec = net.decompress(enc["strings"], enc["shape"])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
order = []
for name, nbytes, e0, e1, flops in prof.records:
    bound = "tensor" if flops and flops / max(nbytes, 1) > ridge else "hbm"
    order.append({"family": f"{name}|{bound}", "bytes": int(nbytes), "flops": int(flops)})
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"batch": B, "precision": ops.precision(), "launches": order},
          open(os.path.join(ROOT, "gpurun_out", "one_step_families.json"), "w"))
print("ok", B, sum(len(s) for s in enc["strings"][0]), len(order), "stf_b200 launches")
