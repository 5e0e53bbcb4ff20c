"""CUDA-graph plumbing: capture a fixed-shape GPU segment once, replay it per call.

The codec's GPU work between two host hand-offs (entropy coding runs on the host) is a static chain of
~1000 small launches (our kernels + cuDNN convolutions); launched eagerly it is CPU launch-bound.  A
`Segment` captures such a chain into one CUDA graph on a private memory pool shared by all segments of
a plan, so tensors produced by one segment stay valid for the next."""
import torch

from . import ops


class Segment:
    """fn(*static_inputs) -> tensor or tuple of tensors, captured once.  __call__ copies the caller's inputs
    into the static input buffers, replays, and returns the (static) outputs: consume or clone them before the
    next replay."""

    def __init__(self, fn, example_inputs, pool=None, warmup=2):
        self.static_in = [t.detach().clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        l0 = ops.launch_count()
        with torch.cuda.graph(self.graph, pool=pool):
            self.static_out = fn(*self.static_in)
        self.launches = ops.launch_count() - l0    # stf_b200 kernels inside the graph (per replay)
        self.graph.replay()                        # leave valid data in the static buffers for later captures
        torch.cuda.synchronize()

    def pool(self):
        return self.graph.pool()

    def __call__(self, *inputs):
        for s, t in zip(self.static_in, inputs):
            if s.data_ptr() != t.data_ptr():
                s.copy_(t, non_blocking=True)
        self.graph.replay()
        ops.count_replayed_launches(self.launches)
        return self.static_out
