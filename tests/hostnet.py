"""A small host network for the drop-in tests, written for this repository (the reference's networks.py cannot travel
to the GPU box).  It uses the shift-layer modules exactly the way the reference's generator does:

* construction (models/networks.py:307-319): ``IPSR_model(opt.threshold, opt.fixed_mask, opt.shift_sz, opt.stride,
  opt.mask_thred, opt.triple_weight)``, ``InnerCos(strength=, skip=)``, ``InnerCos2(strength=, skip=)``, each given the
  global mask and appended to a caller-owned list -- the only handle models/IPSR.py keeps on them (:51);
* placement (models/networks.py:347-348): ``[..., conv3x3, ipsr, innerCos, norm, <inner levels>, innerCos2, ...]`` with
  the level's input concatenated to its output (:362-366);
* per iteration (models/IPSR.py:155-164,186-189,253-263): set_mask on the three lists, set_ref with an object that has
  ``.relu4_3``, set_target, forward, read ``.loss.data`` of every InnerCos, backward.
"""
import collections

import torch
import torch.nn as nn

from deepinpainting_b200.models import IPSR_model, InnerCos, InnerCos2

RefFeatures = collections.namedtuple("RefFeatures", ["relu1_2", "relu2_2", "relu3_3", "relu4_3"])


class Opt:
    threshold = 5 / 16.0
    fixed_mask = 1
    shift_sz = 1
    stride = 1
    mask_thred = 1
    triple_weight = 1
    strength = 1
    skip = 0


class _Innermost(nn.Module):
    def __init__(self, nc):
        super().__init__()
        self.body = nn.Sequential(nn.LeakyReLU(0.2), nn.Conv2d(nc, nc, 3, 1, 1), nn.ReLU(), nn.Conv2d(nc, nc, 3, 1, 1),
                                  nn.InstanceNorm2d(nc, affine=True))

    def forward(self, x):
        return torch.cat([self.body(x), x], 1)


class ShiftLevel(nn.Module):
    """One encoder / decoder level that hosts the shift layer at half the input resolution."""

    def __init__(self, outer_nc, inner_nc, opt, shift_list, cos_list, cos2_list, mask_global):
        super().__init__()
        shift = IPSR_model(opt.threshold, opt.fixed_mask, opt.shift_sz, opt.stride, opt.mask_thred, opt.triple_weight)
        shift.set_mask(mask_global, 3, opt.threshold)
        shift_list.append(shift)
        cos = InnerCos(strength=opt.strength, skip=opt.skip)
        cos.set_mask(mask_global, opt)
        cos_list.append(cos)
        cos2 = InnerCos2(strength=opt.strength, skip=opt.skip)
        cos2.set_mask(mask_global, opt)
        cos2_list.append(cos2)
        down = [nn.LeakyReLU(0.2), nn.Conv2d(outer_nc, outer_nc, 4, 2, 3, dilation=2), nn.InstanceNorm2d(outer_nc, affine=True),
                nn.LeakyReLU(0.2), nn.Conv2d(outer_nc, inner_nc, 3, 1, 1), shift, cos, nn.InstanceNorm2d(inner_nc, affine=True)]
        up = [cos2, nn.ReLU(), nn.ConvTranspose2d(inner_nc * 2, outer_nc, 3, 1, 1), nn.InstanceNorm2d(outer_nc, affine=True),
              nn.ReLU(), nn.ConvTranspose2d(outer_nc, outer_nc, 4, 2, 1), nn.InstanceNorm2d(outer_nc, affine=True)]
        self.model = nn.Sequential(*(down + [_Innermost(inner_nc)] + up))

    def forward(self, x):
        return torch.cat([self.model(x), x], 1)


class HostModel:
    """The slice of models/IPSR.py that talks to the shift-layer modules."""

    def __init__(self, outer_nc, inner_nc, opt, mask_size, device):
        self.opt = opt
        self.shift_list, self.cos_list, self.cos2_list = [], [], []
        self.mask_global = torch.zeros(1, 1, mask_size, mask_size, dtype=torch.bool, device=device)
        self.net = ShiftLevel(outer_nc, inner_nc, opt, self.shift_list, self.cos_list, self.cos2_list, self.mask_global).to(device)

    def set_latent_mask(self, mask_global):                       # models/IPSR.py:155-158
        self.mask_global = mask_global
        self.shift_list[0].set_mask(mask_global, 3, self.opt.threshold)
        self.cos_list[0].set_mask(mask_global, self.opt)
        self.cos2_list[0].set_mask(mask_global, self.opt)

    def set_ref_latent(self, ref_relu4_3):                        # models/IPSR.py:162-164
        self.shift_list[0].set_ref(RefFeatures(None, None, None, ref_relu4_3))

    def set_gt_latent(self, gt_relu4_3):                          # models/IPSR.py:186-189
        self.cos_list[0].set_target(gt_relu4_3)
        self.cos2_list[0].set_target(gt_relu4_3)

    def side_losses(self):                                        # models/IPSR.py:253-263 (values only: .loss.data)
        total = 0
        for layer in self.cos_list + self.cos2_list:
            total = total + layer.loss.data
        return total
