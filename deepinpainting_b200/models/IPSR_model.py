"""``IPSR_model`` -- the shift module with the reference's constructor and methods
(models/IPSR_model.py:9-68): holds the feature mask, the reference features and the flag vectors
and calls IPSRFunction.  It owns no parameters or buffers, so state_dicts are unaffected.
"""
import weakref

import torch
import torch.nn as nn

from .. import shift_ops
from ..util import util
from .IPSRFunction import IPSRFunction


# the most recently constructed shift module that no InnerCos has claimed yet (the generator builds
# `ipsr = IPSR_model(...)` and then `innerCos = InnerCos(...)` for the module right behind it, models/networks.py:307-314)
_last_unlinked = None


class IPSR_model(nn.Module):
    def __init__(self, threshold, fixed_mask, shift_sz=1, stride=1, mask_thred=1, triple_weight=1):
        super(IPSR_model, self).__init__()
        global _last_unlinked
        self._cos_ref = None
        _last_unlinked = weakref.ref(self)
        self.threshold = threshold
        self.fixed_mask = fixed_mask
        self.shift_sz = shift_sz
        self.stride = stride
        self.mask_thred = mask_thred
        self.triple_weight = triple_weight
        self.cal_fixed_flag = True
        self.sp_x = None
        self.sp_y = None
        self._flag_key = None

    def set_mask(self, mask_global, layer_to_last, threshold):
        """Reference contract: ``mask_global`` is [1,1,S,S] and is shared by the whole batch
        (models/IPSR_model.py:30-34, models/IPSRFunction.py:32).  Extension (BASELINE.json configs[4], free-form
        masks per sample): a [B,1,S,S] mask gives every sample its own feature mask; the forward is still ONE batched
        operator call -- every kernel indexes its image's own flag / index rows (images are independent,
        models/IPSRFunction.py:46)."""
        if mask_global.dim() == 4 and mask_global.size(0) > 1:
            dev = mask_global.device if mask_global.is_cuda else torch.device("cuda", torch.cuda.current_device())
            util._require_binary(mask_global, "mask_global")
            feats = shift_ops.feat_mask_batched(mask_global.to(dev)[:, 0], layer_to_last, threshold)      # one launch per layer
            self.masks = list(feats.unbind(0))
            self._feats = feats
            self.mask = self.masks[0]
            self._flag_key = None
            return feats
        self.masks = None
        mask = util.cal_feat_mask(mask_global, layer_to_last, threshold)
        self.mask = mask.squeeze()
        self._flag_key = None                      # a new mask invalidates the cached flag vectors
        return self.mask

    def set_ref(self, latent_ref):
        self.ref = latent_ref

    def link_innercos(self, innercos):
        """Tell the layer which InnerCos module consumes its output (`ipsr, innerCos, downnorm_3`, models/networks.py:347):
        the forward then computes that module's loss in its paste kernel, while the output tiles are still in shared
        memory, and InnerCos.forward picks the value up instead of reading the output again.  Done automatically for an
        InnerCos constructed right after the layer; results are identical either way."""
        self._cos_ref = None if innercos is None else weakref.ref(innercos)

    def _fused_cos_request(self, input):
        cos = self._cos_ref() if self._cos_ref is not None else None
        if cos is None or not shift_ops.config["fuse_innercos"] or cos.skip or self.shift_sz != 1 or self.stride != 1:
            return None, None
        t, m = cos.target, getattr(cos, "mask", None)
        if not (torch.is_tensor(t) and torch.is_tensor(m) and input.dtype == torch.float32 and t.dtype == torch.float32
                and t.is_cuda and t.shape == input.shape and m.numel() == input.size(2) * input.size(3)):
            return None, None
        m = m if (m.device == input.device and m.dtype == torch.float32) else m.to(input.device, torch.float32)
        return shift_ops.FusedCos(target=t.detach(), mask=m, strength=float(cos.strength), crit=cos.crit), cos.fuse_key()

    def _call_function(self, input, mask2d):
        req, key = self._fused_cos_request(input)
        shift_ops.request_fused_cos(req)
        try:
            out = IPSRFunction.apply(input, mask2d, self.ref, self.shift_sz, self.stride, self.triple_weight, self.flag,
                                     self.nonmask_point_idx, self.mask_point_idx, self.flatten_offsets, self.sp_x, self.sp_y)
        finally:
            shift_ops.request_fused_cos(None)
        loss = shift_ops.take_fused_loss()
        if req is not None and loss is not None:
            out._ipsr_fused_cos = (loss, key)               # read by the InnerCos that receives this very tensor
        return out

    def _forward_per_sample(self, input):
        """One batched operator call with a flag row per sample (flag vectors cached per mask)."""
        if len(self.masks) != input.size(0):
            raise ValueError("%d per-sample masks for a batch of %d" % (len(self.masks), input.size(0)))
        _, self.c, self.h, self.w = input.size()
        if not (torch.is_tensor(self.sp_x) or torch.is_tensor(self.sp_y)):
            self.sp_x, self.sp_y = util.cal_sps_for_Advanced_Indexing(self.h, self.w)
        key = (id(self._feats), self._feats._version, self.h, self.w, self.shift_sz, self.stride, self.mask_thred, input.device)
        if key != self._flag_key:
            feats = self._feats if self._feats.device == input.device else self._feats.to(input.device)
            if tuple(feats.shape[1:]) != (self.h, self.w):
                raise ValueError("mask %s does not match the feature map %dx%d" % (tuple(feats.shape[1:]), self.h, self.w))
            # flag / index vectors of every sample in ONE launch and one host synchronisation
            import math
            mi = shift_ops.build_flags_batched(feats, self.shift_sz, self.stride, int(math.ceil(self.mask_thred)))
            counts = mi.m_count.tolist()
            # the reference's vectors, one row per sample; mask_point_idx rows are ragged, so it stays a list
            self.flag = mi.flag.long()
            P = self.flag.size(1)
            self.nonmask_point_idx = torch.arange(P, dtype=torch.int64, device=input.device)
            self.mask_point_idx = [mi.mask_idx[b, :c].long() for b, c in enumerate(counts)]
            self.flatten_offsets = None                       # unused by the operator (models/IPSRFunction.py:88-89); per-sample: not built
            shift_ops.register_mask_index(self.flag, mi)
            self._flag_key = key
        return self._call_function(input, self.masks[0])

    def forward(self, input):
        if getattr(self, "masks", None) is not None:
            return self._forward_per_sample(input)
        _, self.c, self.h, self.w = input.size()
        # The reference recomputes the flag vectors on every forward (cal_fixed_flag never turns
        # False, IPSR_model.py:23,45,53).  They depend only on (mask, h, w, shift_sz, stride,
        # mask_thred), so they are rebuilt here exactly when one of those changed.
        key = (id(self.mask), self.mask._version, self.h, self.w, self.shift_sz, self.stride, self.mask_thred,
               input.device)
        if key != self._flag_key:
            latter = input.narrow(0, 0, 1).data
            mask_dev = self.mask.to(input.device) if self.mask.device != input.device else self.mask
            self.flag, self.nonmask_point_idx, self.flatten_offsets, self.mask_point_idx = \
                util.cal_mask_given_mask_thred(latter.squeeze(0), mask_dev, self.shift_sz, self.stride, self.mask_thred)
            self.cal_fixed_flag = True
            self._flag_key = key
        if not (torch.is_tensor(self.sp_x) or torch.is_tensor(self.sp_y)):
            self.sp_x, self.sp_y = util.cal_sps_for_Advanced_Indexing(self.h, self.w)
        return self._call_function(input, self.mask)

    def __repr__(self):
        return self.__class__.__name__ + '(' \
            + 'threshold: ' + str(self.threshold) \
            + ' ,triple_weight ' + str(self.triple_weight) + ')'
