"""GPU idle-gap analysis of one compress + decompress in the product configuration (developer tool): torch.profiler
kernel intervals -> busy time, idle gaps, and which host-side op each long gap follows.   python tools/gap_trace.py [--batch 32]"""
import argparse
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from stf_b200 import models  # noqa: E402
from stf_b200.synth import synthetic_image  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    args = ap.parse_args()
    net = models.SymmetricalTransFormer()
    torch.nn.Module.load_state_dict(net, bench.synthetic_weights(), strict=False)
    net = net.cuda().eval()
    net.update(force=True)
    xs = [synthetic_image(args.batch, bench.H, bench.W, seed=i).cuda() for i in range(3)]
    for x in xs[:2]:
        enc = net.compress(x)
        net.decompress(enc["strings"], enc["shape"])
    torch.cuda.synchronize()
    marks = {}
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        with torch.profiler.record_function("COMPRESS"):
            enc = net.compress(xs[2])
            torch.cuda.synchronize()
        with torch.profiler.record_function("DECOMPRESS"):
            net.decompress(enc["strings"], enc["shape"])
            torch.cuda.synchronize()
    path = os.path.join(tempfile.gettempdir(), "stf_trace.json")
    prof.export_chrome_trace(path)
    ev = json.load(open(path))["traceEvents"]
    spans = {e["name"]: (e["ts"], e["ts"] + e["dur"]) for e in ev if e.get("name") in ("COMPRESS", "DECOMPRESS") and e.get("cat") == "user_annotation"}
    kern = sorted(((e["ts"], e["ts"] + e["dur"], e["name"]) for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")), key=lambda k: k[0])
    for phase, (t0, t1) in spans.items():
        ks = [k for k in kern if k[0] >= t0 and k[0] <= t1]
        if not ks:
            continue
        # union of intervals (several streams)
        busy, cur_s, cur_e, gaps = 0.0, ks[0][0], ks[0][1], []
        for s, e, n in ks[1:]:
            if s > cur_e:
                busy += cur_e - cur_s
                gaps.append((s - cur_e, cur_e - t0, n))
                cur_s, cur_e = s, e
            else:
                cur_e = max(cur_e, e)
        busy += cur_e - cur_s
        wall = t1 - t0
        big = sorted(gaps, reverse=True)[:12]
        print(f"{phase}: wall {wall / 1e3:.1f} ms, GPU busy {busy / 1e3:.1f} ms, idle {(wall - busy) / 1e3:.1f} ms "
              f"(before first kernel {(ks[0][0] - t0) / 1e3:.2f} ms, after last {(t1 - cur_e) / 1e3:.2f} ms); "
              f"gaps > 100 us: {sum(1 for g in gaps if g[0] > 100)} totalling {sum(g[0] for g in gaps if g[0] > 100) / 1e3:.1f} ms; "
              f"gaps <= 100 us: {sum(g[0] for g in gaps if g[0] <= 100) / 1e3:.1f} ms in {sum(1 for g in gaps if g[0] <= 100)}")
        for g, at, n in big:
            print(f"    gap {g / 1e3:6.2f} ms at +{at / 1e3:6.1f} ms before {n[:70]}")


if __name__ == "__main__":
    main()
