"""Decode-time runtime helpers around QuantLinearLUT (public API).

`GraphedDecodeStep` turns a token step made of QuantLinearLUT calls (224 of them per LLaMA-7B token, each a
~2-10 us kernel) into ONE CUDA-graph replay with static device buffers and pinned host staging:

    runner = GraphedDecodeStep(step_fn, example_x)      # captures step_fn(x_dev) -> y_dev
    y_host = runner(x_host)                             # H2D copy, replay, D2H copy, all on one stream

It exists because the reference drives the path from Python one launch at a time on the legacy stream
(llama.py:226-234, quant_cuda_kernel.cu:147), which costs more host time per call than the kernels take on a
B200; our kernels launch on the current stream precisely so that they can be captured.
"""
import torch


class GraphedDecodeStep:
    def __init__(self, step_fn, example_x, warmup=3):
        assert example_x.is_cuda, "example_x must be a CUDA tensor (static input buffer is cloned from it)"
        self.step_fn = step_fn
        self.x_dev = example_x.clone()
        self.x_host = torch.empty(example_x.shape, dtype=example_x.dtype, pin_memory=True)
        self.stream = torch.cuda.Stream()
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):  # also sizes the per-stream fused-path workspace before capture
                y = step_fn(self.x_dev)
        torch.cuda.current_stream().wait_stream(self.stream)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self.y_dev = step_fn(self.x_dev)
        self.y_host = torch.empty(self.y_dev.shape, dtype=self.y_dev.dtype, pin_memory=True)
        self.h2d_bytes = self.x_host.numel() * self.x_host.element_size()
        self.d2h_bytes = self.y_host.numel() * self.y_host.element_size()

    def replay(self):
        """Device-only: replay on the current stream with whatever is in the static input buffer."""
        self.graph.replay()
        return self.y_dev

    def __call__(self, x_host):
        """End to end from host memory: pinned H2D, replay, pinned D2H, synchronise; returns the pinned host result."""
        self.x_host.copy_(x_host)
        self.x_dev.copy_(self.x_host, non_blocking=True)
        self.graph.replay()
        self.y_host.copy_(self.y_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.y_host
