"""CUDA-graph capture of a scattering forward (no-grad).

Small configurations are launch-bound: BASELINE configs[0] (J=2, 32 x 32, batch 128) moves 1.2 MB per image through
a handful of kernels that each finish in microseconds, so the forward costs what its launches cost.  The whole launch
schedule of one forward (a fixed sequence of kernels on one stream, no host synchronisation once the filters are
bound) is captured once into a CUDA graph and replayed with a single launch.

    S = Scattering2D(2, (32, 32)).cuda()          # or kymatio.torch.Scattering2D(..., backend='torch_b200')
    G = GraphedScattering(S, torch.empty(128, 32, 32, device='cuda'))
    y = G(x)                                     # x: same shape/dtype/device; returns G's static output tensor
"""
import torch

__all__ = ["GraphedScattering"]


class GraphedScattering:
    def __init__(self, module, example, warmup=2):
        if not example.is_cuda:
            raise TypeError("GraphedScattering needs a CUDA example input")
        self.module = module
        self._x = example.clone()
        with torch.no_grad():
            side = torch.cuda.Stream(device=example.device)
            side.wait_stream(torch.cuda.current_stream(example.device))
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):        # binds the filters (the only host synchronisation) before capture
                    module(self._x)
            torch.cuda.current_stream(example.device).wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._y = module(self._x)

    def __call__(self, x):
        if x.shape != self._x.shape or x.dtype != self._x.dtype or x.device != self._x.device:
            raise ValueError("GraphedScattering was captured for input %s %s on %s" %
                             (tuple(self._x.shape), self._x.dtype, self._x.device))
        self._x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self._y
