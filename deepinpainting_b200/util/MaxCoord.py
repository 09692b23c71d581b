"""``MaxCoord`` with the reference's interface (util/MaxCoord.py:12-28).

The fused layer never materialises the score tensor and therefore never calls this class; it is
kept for callers that hold a score tensor and is backed by the ``ipsr_maxcoord`` kernel.
"""
import torch

from .. import shift_ops


class MaxCoord():
    def __init__(self):
        pass

    def update_output(self, input, sp_x, sp_y):
        assert input.dim() == 4, "Input must be 3D or 4D(batch)."
        assert input.size(0) == 1, "The first dimension of input has to be 1!"
        _, P, H, W = input.size()
        ind, v_max = shift_ops.maxcoord(input.reshape(P, H * W))
        # the reference returns zeros_like(input) first; its caller discards it (IPSRFunction.py:65,78)
        output = torch.zeros_like(input)
        return output, ind, v_max
