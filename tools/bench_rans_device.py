"""Device rANS decoder micro-benchmark: B streams x n symbols per call (one slice of a sub-batch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from stf_b200 import ans
from stf_b200.entropy_models import GaussianConditional
from stf_b200.models import get_scale_table
gc = GaussianConditional(None); gc.update_scale_table(get_scale_table()); tab = gc.rans_table()
table = get_scale_table().numpy()
for B, spread in ((24, 1.0), (24, 0.2), (8, 1.0), (32, 1.0)):
    n, slices = 49152, 4
    rng = np.random.default_rng(B)
    ix = np.minimum(63, np.abs(rng.normal(30, 15, size=(B, n * slices))).astype(np.int32))
    sy = np.rint(rng.standard_normal((B, n * slices)) * table[ix] * spread).astype(np.int32)
    strings = ans.encode_rows(tab, sy, ix)
    ds = ans.DeviceStreams(B, sum(len(s) for s in strings) // 4 + 4 * B, "cuda")
    ds.load(strings); ds.upload()
    idx_d = torch.from_numpy(ix).cuda(); out = torch.empty_like(idx_d)
    def run():
        for k in range(slices):
            ans.decode_device(tab, ds, idx_d[:, k * n:(k + 1) * n], out[:, k * n:(k + 1) * n], first=(k == 0))
    run(); torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), sy)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / slices
    print(f"B={B} n={n} bits/symbol {8 * sum(len(s) for s in strings) / (B * n * slices):.2f}: {ms:.2f} ms per slice call = {ms * 1e6 / n:.0f} ns per symbol step "
          f"({ms * 1e6 / n * 1.9:.0f} cycles)")
