"""Per-kernel CUDA-event timing of the stf_b200 launches (used by bench.py's roofline leg).

    with profiler.capture() as prof:
        net.compress(x)
    torch.cuda.synchronize(); prof.summary()

While a capture is active every C-ABI launch made through stf_b200.ops is bracketed by a pair of
CUDA events recorded on the launching stream; summary() groups them by kernel with the algorithmic
bytes the wrapper declared (DESIGN.md "algorithmic bytes per unit").  Event pairs add a few
microseconds per launch, so captured steps are never used for throughput numbers.
"""
import contextlib

ACTIVE = None


class Capture:
    def __init__(self):
        self.records = []      # (name, bytes, ev0, ev1, flops)
        self.total_ms = None
        self._t0 = self._t1 = None

    def add(self, name, nbytes, e0, e1, flops=0):
        self.records.append((name, nbytes, e0, e1, flops))

    def summary(self, ridge_flop_per_byte=None):
        """Group by kernel family.  With a ridge point (flop per byte at which the tensor peak equals the HBM peak) every
        launch is also classed "hbm" or "tensor" by its own arithmetic intensity, and families are split per class:
        one kernel (the GEMM engine) serves HBM-bound C <= 96 layers and tensor-bound C >= 192 layers / convolutions."""
        fam = {}
        for name, nbytes, e0, e1, flops in self.records:
            key = name
            if ridge_flop_per_byte is not None:
                bound = "tensor" if flops and flops / max(nbytes, 1) > ridge_flop_per_byte else "hbm"
                key = f"{name}|{bound}"
            f = fam.setdefault(key, {"name": name, "ms": 0.0, "bytes": 0, "flops": 0, "launches": 0})
            if ridge_flop_per_byte is not None:
                f["bound"] = bound
            f["ms"] += e0.elapsed_time(e1)
            f["bytes"] += int(nbytes)
            f["flops"] += int(flops)
            f["launches"] += 1
        if self._t0 is not None:
            self.total_ms = self._t0.elapsed_time(self._t1)
        return fam


@contextlib.contextmanager
def capture():
    global ACTIVE
    import torch
    cap = Capture()
    cap._t0, cap._t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prev, ACTIVE = ACTIVE, cap
    cap._t0.record()
    try:
        yield cap
    finally:
        cap._t1.record()
        ACTIVE = prev
