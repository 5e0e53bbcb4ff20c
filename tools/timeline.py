"""Host / device timeline of pipelined compress()+decompress() steps (developer tool): how long is the GPU idle while the
host codes, how long does the host wait for the GPU?   [taskset -c 0-3] python tools/timeline.py [batch] [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from stf_b200 import models  # noqa: E402
from stf_b200.synth import synthetic_image  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
net = models.SymmetricalTransFormer()
torch.nn.Module.load_state_dict(net, bench.synthetic_weights(), strict=False)
net = net.cuda().eval()
net.update(force=True)
xs = [synthetic_image(B, bench.H, bench.W, seed=i).cuda() for i in range(steps + 2)]
for x in xs[:2]:
    enc = net.compress(x)
    net.decompress(enc["strings"], enc["shape"])
torch.cuda.synchronize()


def union(spans):
    spans = sorted(spans)
    tot, cur0, cur1 = 0.0, None, None
    for a, b in spans:
        if cur1 is None or a > cur1:
            if cur1 is not None:
                tot += cur1 - cur0
            cur0, cur1 = a, b
        else:
            cur1 = max(cur1, b)
    return tot + (cur1 - cur0 if cur1 is not None else 0.0)


for x in xs[2:]:
    for phase in ("compress", "decompress"):
        models.TRACE = []
        base = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        base.record()
        torch.cuda.synchronize()
        t_base = time.perf_counter()
        if phase == "compress":
            enc = net.compress(x)
        else:
            net.decompress(enc["strings"], enc["shape"])
        torch.cuda.synchronize()
        t_end = time.perf_counter()
        tr, models.TRACE = models.TRACE, None
        host = [(k[1:4], (k[4] - t_base) * 1e3, (k[5] - t_base) * 1e3) for k in tr if k[0] == "host"]
        dev = [(k[1:4], base.elapsed_time(k[4]), base.elapsed_time(k[5])) for k in tr if k[0] == "dev"]
        total = (t_end - t_base) * 1e3
        print(f"--- {phase}: {total:.1f} ms wall; device busy (union of traced spans) {union([(a, b) for _, a, b in dev]):.1f} ms; "
              f"host rANS {sum(b - a for k, a, b in host if 'rans' in k[0] or k[0] == 'dec.first'):.1f} ms; "
              f"host waits {sum(b - a for k, a, b in host if 'wait' in k[0]):.1f} ms")
        by = {}
        for k, a, b in dev:
            by.setdefault(k[0], []).append(b - a)
        print("    device spans: " + ", ".join(f"{n} x{len(v)} avg {sum(v) / len(v):.2f} ms" for n, v in by.items()))
        by = {}
        for k, a, b in host:
            by.setdefault(k[0], []).append(b - a)
        print("    host spans:   " + ", ".join(f"{n} x{len(v)} avg {sum(v) / len(v):.2f} ms" for n, v in by.items()))
        if os.environ.get("TIMELINE_DUMP"):
            rows = [("H", k, a, b) for k, a, b in host] + [("D", k, a, b) for k, a, b in dev]
            for kind, k, a, b in sorted(rows, key=lambda r: r[2]):
                print(f"      {kind} {k[0]:10s} part {k[1]} step {k[2]:2d}  {a:8.2f} -> {b:8.2f}  ({b - a:6.2f} ms)")
