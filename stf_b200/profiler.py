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
        self.records = []      # (name, bytes, ev0, ev1)
        self.total_ms = None
        self._t0 = self._t1 = None

    def add(self, name, nbytes, e0, e1):
        self.records.append((name, nbytes, e0, e1))

    def summary(self):
        fam = {}
        for name, nbytes, e0, e1 in self.records:
            f = fam.setdefault(name, {"name": name, "ms": 0.0, "bytes": 0, "launches": 0})
            f["ms"] += e0.elapsed_time(e1)
            f["bytes"] += int(nbytes)
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
