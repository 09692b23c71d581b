"""``NonparametricShift`` with the reference's interface (util/NonparametricShift.py:8-86).

The fused layer does not build these modules (the reference builds four per image and uses two);
this class serves callers that want the encoder / decoder convolutions themselves.  Extraction and
L2-normalisation run in the ``ipsr_extract_normalize`` kernel (1 x 1 patches) or the ``ipsr_patch_rows`` kernel
(patch_size / stride > 1).
"""
import torch
import torch.nn as nn

from .. import shift_ops


class NonparametricShift(object):
    def buildAutoencoder(self, target_img, normalize, interpolate, nonmask_point_idx, mask_point_idx, patch_size=1, stride=1):
        nDim = 3
        assert target_img.dim() == nDim, 'target image must be of dimension 3.'
        C = target_img.size(0)
        patches_all, patches_part, patches_mask = self._extract_patches(target_img, patch_size, stride,
                                                                        nonmask_point_idx, mask_point_idx)
        npatches_part = patches_part.size(0)
        npatches_all = patches_all.size(0)
        conv_enc_non_mask, conv_dec_non_mask = self._build(patch_size, stride, C, patches_part, npatches_part, normalize, interpolate)
        conv_enc_all, conv_dec_all = self._build(patch_size, stride, C, patches_all, npatches_all, normalize, interpolate)
        return conv_enc_all, conv_enc_non_mask, conv_dec_all, conv_dec_non_mask, patches_part, patches_mask

    def _build(self, patch_size, stride, C, target_patches, npatches, normalize, interpolate):
        if normalize:
            raise NotImplementedError
        if interpolate:
            raise NotImplementedError
        # p * (1 / (||p|| + 1e-8)) per patch (:36-40): the kernel computes 1/(||p||+1e-8)
        if patch_size == 1:
            flat = target_patches.reshape(1, npatches, -1).permute(0, 2, 1).contiguous()      # [1, C, P]
            _, inv = shift_ops.extract_normalize(flat.view(1, flat.size(1), 1, npatches))
        else:
            # every [C,k,k] patch is an image holding exactly one k x k patch
            inv = torch.cat([shift_ops.patch_rows(target_patches[i:i + 32768].contiguous(), patch_size, 1)[1]
                             for i in range(0, npatches, 32768)])
        enc_patches = target_patches * inv.view(npatches, 1, 1, 1)
        conv_enc = nn.Conv2d(C, npatches, kernel_size=patch_size, stride=stride, bias=False).to(target_patches.device)
        conv_enc.weight.data = enc_patches
        conv_dec = nn.ConvTranspose2d(npatches, C, kernel_size=patch_size, stride=stride, bias=False).to(target_patches.device)
        conv_dec.weight.data = target_patches
        return conv_enc, conv_dec

    def _extract_patches(self, img, patch_size, stride, nonmask_point_idx, mask_point_idx):
        n_dim = 3
        assert img.dim() == n_dim, 'image must be of dimension 3.'
        C, H, W = img.size()
        if patch_size != 1 or stride != 1:
            rows, _ = shift_ops.patch_rows(img.unsqueeze(0), patch_size, stride)          # unfold order (c, dy, dx), :65-68
            patches_all = rows.view(-1, C, patch_size, patch_size)
        else:
            xt, _ = shift_ops.extract_normalize(img.unsqueeze(0))
            patches_all = xt.view(H * W, C, 1, 1)
        dev = patches_all.device
        patches = patches_all.index_select(0, nonmask_point_idx.to(dev))
        maskpatches = patches_all.index_select(0, mask_point_idx.to(dev))
        return patches_all, patches, maskpatches

    def _extract_patches_mask(self, img, patch_size, stride, nonmask_point_idx, mask_point_idx):
        return self._extract_patches(img, patch_size, stride, nonmask_point_idx, mask_point_idx)[2]
