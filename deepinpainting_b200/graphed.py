"""CUDA-graph replay of one forward + backward of the shift layer through the module API.

The layer enqueues its kernels without host synchronisation or allocation-dependent control flow, so a whole
``y = layer(x); y.backward(g)`` is capturable.  ``GraphedShiftStep`` captures it once per shape around STATIC buffers
(``x``, ``ref``, ``g`` in; ``out``, ``gin`` out): a step is then "fill the static inputs, replay" -- one driver call
instead of a dozen launches and the Python between them (models/IPSR_model.py:42-63 + models/IPSRFunction.py:13-178 of the
reference, per training iteration).
"""
from __future__ import annotations

import collections

import torch

Ref = collections.namedtuple("Ref", ["relu4_3"])


class GraphedShiftStep:
    def __init__(self, layer, B: int, C: int, H: int, W: int, device, warmup: int = 3):
        """``layer``: an ``IPSR_model`` whose mask has been set (``set_mask``).  The mask must not change afterwards
        (the flag vectors are baked into the graph); call ``recapture()`` after a ``set_mask``."""
        self.layer = layer
        dev = torch.device(device)
        self.x = torch.zeros(B, C, H, W, device=dev)
        self.ref = torch.zeros(B, C, H, W, device=dev)
        self.g = torch.zeros(B, C, H, W, device=dev)
        self.out = None
        self.gin = None
        self.graph = None
        self._warmup = warmup
        self.recapture()

    def _step(self):
        xin = self.x.detach().requires_grad_(True)
        self.layer.set_ref(Ref(self.ref))
        y = self.layer(xin)
        y.backward(self.g)
        return y.detach(), xin.grad

    def recapture(self):
        dev = self.x.device
        # non-degenerate contents for the warm-up (all-zero patches are all ties; harmless, but pointless work)
        self.x.normal_()
        self.ref.normal_().relu_()
        self.g.normal_()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(self._warmup):                   # builds the flag vectors / plan outside the capture
                self._step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out, self.gin = self._step()

    def replay(self):
        """Runs the captured forward + backward on the current stream; results land in ``self.out`` / ``self.gin``."""
        self.graph.replay()
        return self.out, self.gin

    def __call__(self, x=None, ref=None, g=None):
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if ref is not None:
            self.ref.copy_(ref, non_blocking=True)
        if g is not None:
            self.g.copy_(g, non_blocking=True)
        return self.replay()
