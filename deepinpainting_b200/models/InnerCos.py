"""``InnerCos`` -- identity layer that records a masked consistency loss, with the reference's
interface (models/InnerCos.py:5-53).  loss = crit(in_data * mask * strength, target) is one fused
kernel (``innercos_loss_fwd``) and stays differentiable w.r.t. in_data (``innercos_loss_bwd``).
"""
import torch
import torch.nn as nn

from .. import shift_ops
from ..util import util


class _InnerCosLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, in_data, mask, target, strength, crit, c_limit):
        # fp16 / bf16 activations (autocast host network): the loss is computed in fp32
        ctx.in_dtype = in_data.dtype
        in_data, target = in_data.detach().float(), target.detach().float()
        ctx.save_for_backward(in_data, mask, target)
        ctx.strength, ctx.crit, ctx.c_limit = strength, crit, c_limit
        return shift_ops.innercos_loss(in_data, mask, target, strength, crit, c_limit)

    @staticmethod
    def backward(ctx, grad_loss):
        in_data, mask, target = ctx.saved_tensors
        gx = shift_ops.innercos_loss_grad(in_data, mask, target, grad_loss.float().contiguous(), ctx.strength, ctx.crit, ctx.c_limit)
        if ctx.in_dtype != torch.float32:
            gx = gx.to(ctx.in_dtype)
        return gx, None, None, None, None, None


class _FusedInnerCosLoss(torch.autograd.Function):
    """The loss value was already computed by the shift layer's paste kernel (IPSR_model.link_innercos); this node only
    keeps it differentiable w.r.t. in_data, with the same backward kernel as _InnerCosLoss."""

    @staticmethod
    def forward(ctx, in_data, value, mask, target, strength, crit):
        ctx.save_for_backward(in_data.detach(), mask, target.detach())
        ctx.strength, ctx.crit = strength, crit
        return value.clone()

    @staticmethod
    def backward(ctx, grad_loss):
        in_data, mask, target = ctx.saved_tensors
        gx = shift_ops.innercos_loss_grad(in_data, mask, target, grad_loss.float().contiguous(), ctx.strength, ctx.crit, None)
        return gx, None, None, None, None, None


class InnerCos(nn.Module):
    def __init__(self, crit='MSE', strength=1, skip=0):
        super(InnerCos, self).__init__()
        self.crit = crit
        self.strength = strength
        self.target = None
        self.skip = skip
        self._c_limit = None
        if type(self) is InnerCos:
            # the generator constructs the InnerCos that FOLLOWS a shift layer right after that layer
            # (models/networks.py:307-314): let the layer compute this module's loss in its paste kernel
            import importlib
            shift_module = importlib.import_module(__package__ + ".IPSR_model")     # (the package re-exports the class under this name)
            ref = getattr(shift_module, "_last_unlinked", None)
            layer = ref() if ref is not None else None
            if layer is not None and layer._cos_ref is None:
                layer.link_innercos(self)
            shift_module._last_unlinked = None

    def fuse_key(self):
        """Identity of everything the loss depends on besides in_data: a fused value is only accepted when it was computed
        for exactly this target / mask / strength / criterion."""
        t, m = self.target, getattr(self, "mask", None)
        return (id(t), getattr(t, "_version", None), id(m), getattr(m, "_version", None), float(self.strength), self.crit)

    def set_mask(self, mask_global, opt):
        mask = util.cal_feat_mask(mask_global, 3, opt.threshold)
        self.mask = mask.squeeze().float()

    def set_target(self, targetIn):
        self.target = targetIn

    def get_target(self):
        return self.target

    def forward(self, in_data):
        if not self.skip:
            self.bs = in_data.size(0)
            self.c = in_data.size(1) if self._c_limit is None else min(self._c_limit, in_data.size(1))
            self.former = in_data if self._c_limit is None else in_data.narrow(1, 0, self.c)   # InnerCos2.py:38
            mask = self.mask if self.mask.device == in_data.device else self.mask.to(in_data.device)
            fused = getattr(in_data, "_ipsr_fused_cos", None)
            if fused is not None and self._c_limit is None and fused[1] == self.fuse_key():
                # already computed by the shift layer's paste kernel on this very tensor
                if in_data.requires_grad and torch.is_grad_enabled():
                    self.loss = _FusedInnerCosLoss.apply(in_data, fused[0], mask.float(), self.target, float(self.strength), self.crit)
                else:
                    self.loss = fused[0]
            else:
                self.loss = _InnerCosLoss.apply(in_data, mask, self.target, float(self.strength), self.crit, self._c_limit)
            self.output = in_data
        else:
            self.loss = 0
            self.output = in_data
        return self.output

    @property
    def former_in_mask(self):
        """``torch.mul(self.former, self.mask)`` of the reference (InnerCos.py:33): kept as an attribute for the contract,
        materialised only when somebody reads it (the loss kernel never writes the product to memory)."""
        return torch.mul(self.former, self.mask.to(self.former.dtype))

    def backward(self, retain_graph=True):
        if not self.skip:
            self.loss.backward(retain_graph=retain_graph)
        return self.loss

    def __repr__(self):
        # the reference prints 'True' when the layer is NOT skipped (InnerCos.py:50); kept as is
        skip_str = 'True' if not self.skip else 'False'
        return self.__class__.__name__ + '(' \
            + 'skip: ' + skip_str \
            + ' ,strength: ' + str(self.strength) + ')'
