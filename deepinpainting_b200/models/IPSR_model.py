"""``IPSR_model`` -- the shift module with the reference's constructor and methods
(models/IPSR_model.py:9-68): holds the feature mask, the reference features and the flag vectors
and calls IPSRFunction.  It owns no parameters or buffers, so state_dicts are unaffected.
"""
import torch
import torch.nn as nn

from ..util import util
from .IPSRFunction import IPSRFunction


class IPSR_model(nn.Module):
    def __init__(self, threshold, fixed_mask, shift_sz=1, stride=1, mask_thred=1, triple_weight=1):
        super(IPSR_model, self).__init__()
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
        mask = util.cal_feat_mask(mask_global, layer_to_last, threshold)
        self.mask = mask.squeeze()
        self._flag_key = None                      # a new mask invalidates the cached flag vectors
        return self.mask

    def set_ref(self, latent_ref):
        self.ref = latent_ref

    def forward(self, input):
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
        return IPSRFunction.apply(input, self.mask, self.ref, self.shift_sz, self.stride, self.triple_weight,
                                  self.flag, self.nonmask_point_idx, self.mask_point_idx, self.flatten_offsets,
                                  self.sp_x, self.sp_y)

    def __repr__(self):
        return self.__class__.__name__ + '(' \
            + 'threshold: ' + str(self.threshold) \
            + ' ,triple_weight ' + str(self.triple_weight) + ')'
