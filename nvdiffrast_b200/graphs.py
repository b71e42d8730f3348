"""CUDA-graph capture of rendering steps.

Every op of `nvdiffrast_b200.torch` only enqueues work on the current stream: the library never synchronises, never reads a
device value back and never allocates device memory itself (outputs and scratch come from torch's caching allocator).  A whole
forward + backward step can therefore be recorded once into a CUDA graph and replayed with one launch -- which is what a
launch-bound step wants (small resolutions, pose / texture fitting loops: ~25 kernel launches plus the Python front end
against tens of microseconds of GPU work).  The reference cannot be captured: its rasterizer blocks the host and reads a
counter back on every call (csrc/common/cudaraster/impl/RasterImpl.cpp:362-367).

    step = capture(lambda: render_and_grads(pos, tex))     # pos, tex: static tensors, updated in place between replays
    step.replay(); loss = step.outputs[0]

Limits (same as any torch CUDA graph): static shapes and static input/output tensors; range-mode `rasterize` copies its CPU
`ranges` to the device (not capturable); a partial `diff_attrs` list is passed from host memory at launch time (captured by
value, so it is fine as long as it does not change).
"""
import torch


class CapturedStep:
    def __init__(self, graph, outputs):
        self.graph, self.outputs = graph, outputs

    def replay(self):
        self.graph.replay()
        return self.outputs


def capture(fn, warmup=3, pool=None):
    """Run `fn()` `warmup` times on a side stream (sizes every scratch buffer, builds lazily created state), then record one
    more call into a `torch.cuda.CUDAGraph`.  `fn` must read its inputs from, and may return, tensors that stay alive; returns a
    CapturedStep whose `outputs` are the tensors returned by the captured call (rewritten by every `replay()`)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(max(int(warmup), 1)):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, pool=pool):
        outputs = fn()
    return CapturedStep(graph, outputs)
