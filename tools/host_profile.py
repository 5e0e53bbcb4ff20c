"""cProfile of the host side of pipelined compress()+decompress() steps (developer tool): where does the calling thread wait?
   [taskset -c 0-3] python tools/host_profile.py [batch] [steps]"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from stf_b200 import ans, models  # noqa: E402
from stf_b200.synth import synthetic_image  # noqa: E402

rank, world = int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(rank)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
net = models.SymmetricalTransFormer()
torch.nn.Module.load_state_dict(net, bench.synthetic_weights(), strict=False)
net = net.cuda().eval()
net.update(force=True)
xs = [synthetic_image(B, bench.H, bench.W, seed=i).cuda() for i in range(steps + 2)]
for x in xs[:2]:
    enc = net.compress(x)
    net.decompress(enc["strings"], enc["shape"])
torch.cuda.synchronize()
# pinned-copy bandwidth with every rank copying at once (the slice loop moves ~5 MB each way per sub-batch slice)
hbuf = torch.empty(64 << 20, dtype=torch.uint8, pin_memory=True)
dbuf = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
if world > 1:
    dist.barrier()
for name, src, dst in (("d2h", dbuf, hbuf), ("h2d", hbuf, dbuf)):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(8):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if rank == 0:
        print(f"{name}: {8 * 64 / 1024 / (time.perf_counter() - t0):.1f} GiB/s per rank with {world} ranks copying")
small_h = torch.empty(5 << 20, dtype=torch.uint8, pin_memory=True)
small_d = torch.empty(5 << 20, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    small_h.copy_(small_d, non_blocking=True)
    torch.cuda.synchronize()
if rank == 0:
    print(f"5 MiB d2h + sync: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms")
if world > 1:
    dist.barrier()
if rank != 0:
    sys.stdout = open(os.devnull, "w")
print("affinity", len(os.sched_getaffinity(0)), "rans threads", ans.default_threads(), "torch threads", torch.get_num_threads())
pr = cProfile.Profile()
t_c = t_d = 0.0
pr.enable()
for x in xs[2:]:
    t0 = time.perf_counter()
    enc = net.compress(x)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    net.decompress(enc["strings"], enc["shape"])
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    t_c += t1 - t0
    t_d += t2 - t1
pr.disable()
if world > 1:
    dist.barrier()
print(f"compress {t_c / steps * 1e3:.1f} ms  decompress {t_d / steps * 1e3:.1f} ms per step of {B} images")
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(14)
