"""torch.profiler table of one training step (developer tool).  python tools/profile_train.py [--batch 16]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from stf_b200.models import SymmetricalTransFormer  # noqa: E402
from stf_b200.synth import synthetic_image  # noqa: E402
from stf_b200.training import RateDistortionLoss, configure_optimizers, train_step  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    args = ap.parse_args()
    net = SymmetricalTransFormer()
    torch.nn.Module.load_state_dict(net, bench.synthetic_weights(), strict=False)
    net = net.cuda().train()
    opt, aux = configure_optimizers(net)
    crit = RateDistortionLoss(0.0035)
    xs = [synthetic_image(args.batch, 256, 256, seed=i).cuda() for i in range(4)]
    for x in xs[:3]:
        train_step(net, x, crit, opt, aux)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        train_step(net, xs[3], crit, opt, aux)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))


if __name__ == "__main__":
    main()
