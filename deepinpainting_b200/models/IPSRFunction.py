"""``IPSRFunction`` -- the shift operator as a torch.autograd.Function with the reference's exact
call signature (models/IPSRFunction.py:10-178), executed by libipsr_sm100.so.

    IPSRFunction.apply(input, mask, ref, shift_sz, stride, triple_w, flag, nonmask_point_idx,
                       mask_point_idx, flatten_offsets, sp_x, sp_y) -> output

Semantics: per image, every position q is matched against the bank of ALL
L2-normalised 1x1 patches of ``input`` by correlating ``ref.relu4_3`` with it; unmasked positions
receive their best match, masked positions (``flag``) the sequentially blended patch; the backward
routes gradients through the truncated attention exactly as the reference's LongTensor store does.
"""
import torch

from .. import shift_ops


class IPSRFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, input, mask, ref, shift_sz, stride, triple_w, flag, nonmask_point_idx, mask_point_idx,
                flatten_offsets, sp_x, sp_y):
        assert input.dim() == 4, "Input Dim has to be 4"
        assert mask.dim() == 2, "Mask dimension must be 2"
        ctx.triple_w = triple_w
        ctx.flag = flag
        ctx.flatten_offsets = flatten_offsets
        ctx.bz, c_real, ctx.h, ctx.w = input.size()
        ctx.saved_shift = None
        fused_cos = shift_ops.take_fused_request()          # set by IPSR_model.forward for this very call, or None
        # fp16 / bf16 activations (autocast around the host network's convolutions): the layer computes in fp32, as the
        # reference does, and hands back the input's dtype
        ctx.in_dtype = input.dtype
        ref_feat = ref.relu4_3
        if input.dtype != torch.float32:
            input = input.float()
        if ref_feat.dtype != torch.float32:
            ref_feat = ref_feat.float()
        if shift_sz != 1 or stride != 1:
            # The reference computes the output for these settings (:46-133) and then fails storing the attention
            # for backward (:134); its backward indexes with the 1 x 1 geometry (:158-163).  Forward only.
            mi = shift_ops.lookup_mask_index(flag, input.device)
            output, ctx.ind_lst = shift_ops.shift_forward_patches(input.detach(), ref_feat.detach(), mi, shift_sz, stride)
            return output if ctx.in_dtype == torch.float32 else output.to(ctx.in_dtype)
        # sp_x, sp_y, nonmask_point_idx, flatten_offsets and the 2-D mask are accepted and unused,
        # as in the reference; mask_point_idx is implied by flag.
        mi = shift_ops.lookup_mask_index(flag, input.device)
        need_grad = bool(ctx.needs_input_grad[0])
        output, saved = shift_ops.shift_forward(input.detach(), ref_feat.detach(), mi, need_grad=need_grad,
                                                fused_cos=fused_cos if ctx.in_dtype == torch.float32 else None)
        shift_ops.publish_fused_loss(saved.cos_loss)
        ctx.saved_shift = saved
        ctx.ind_lst = saved.ind          # [B, N] int32 arg-max indices (the reference keeps A as int64 [B,N,H,W])
        return output if ctx.in_dtype == torch.float32 else output.to(ctx.in_dtype)

    @staticmethod
    def backward(ctx, grad_output):
        if ctx.saved_shift is None:
            raise NotImplementedError("IPSRFunction.backward is undefined for shift_sz != 1 / stride != 1: the reference "
                                      "fails in forward before it can save the attention (IPSRFunction.py:134)")
        g = grad_output if grad_output.dtype == torch.float32 else grad_output.float()
        grad_input = shift_ops.shift_backward(g, ctx.saved_shift, ctx.triple_w)
        if ctx.in_dtype != torch.float32:
            grad_input = grad_input.to(ctx.in_dtype)
        return grad_input, None, None, None, None, None, None, None, None, None, None, None
