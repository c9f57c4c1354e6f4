"""CUDA-graph replay of a converted network's forward.

The small-tensor configurations (CIFAR ResNet-20, MobileNetV2 at 32x32) are launch-bound: a forward is
~100 kernels of a few microseconds each, and eager execution spends more time in Python than on the GPU.
Every libfq_b200 entry point is capture-safe (no allocation, no host synchronisation, explicit stream,
plain launches only), so the whole forward -- framework convolutions and fake-quant kernels
alike -- can be recorded once and replayed with no host work in between.  The per-block state the
reference exposes (``current_input_max``, ``input_max``, fake-BN ``current_mean/var``) lives in arenas whose
pointers are fixed at conversion time (quantize/convert/_state.py), so the replayed kernels and
``net.update_ema()`` -- called between replays, as in ``for x in loader: g(x); net.update_ema()`` -- keep
reading and writing the same memory.
"""
import torch

__all__ = ["GraphedForward"]


class GraphedForward:
    """``g = GraphedForward(net, example_input); y = g(x)``.

    ``x`` is copied into a static buffer, the captured graph is replayed and the static output is returned
    (clone it if it must outlive the next call).  Capture happens after ``warmup`` eager iterations on a
    side stream, so one-shot state changes (``fix_params()`` caching the quantised weights) have settled.
    """

    def __init__(self, net, example_input, warmup=3):
        self.net = net
        if hasattr(net, "collect_quantized_blocks"):
            # the graph records raw pointers: the packed range / statistics arenas must be final before capture
            # (convert_model builds them; this re-validates after any later move) and must not move afterwards
            from .quantize.convert import _state
            _state.pack(net)
        self.static_in = example_input.detach().clone()
        self.stream = torch.cuda.Stream(device=example_input.device)
        self.stream.wait_stream(torch.cuda.current_stream(example_input.device))
        with torch.cuda.stream(self.stream), torch.no_grad():
            for _ in range(max(1, warmup)):
                net(self.static_in)
        torch.cuda.current_stream(example_input.device).wait_stream(self.stream)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream), torch.no_grad():
            self.static_out = net(self.static_in)
        if hasattr(net, "collect_quantized_blocks"):
            net.__dict__["_fq_graph_live"] = True        # _state.pack raises instead of re-pointing tensors

    def release(self):
        """Drop the graph; the net may be moved / repacked again."""
        self.graph = None
        self.net.__dict__.pop("_fq_graph_live", None)

    def __call__(self, x):
        self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out

    def replay(self):
        self.graph.replay()
        return self.static_out
