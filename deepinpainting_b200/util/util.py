"""Mask helpers of the shift layer -- same names, arguments and results as the reference's
``util/util.py`` (:68-174), computed by libipsr_sm100.so on the GPU instead of host loops.

Only the functions on the hot path are mirrored (cal_feat_mask, cal_mask_given_mask_thred,
cal_sps_for_Advanced_Indexing); the image / diagnostic helpers of the reference's util.py are out
of scope.
"""
from __future__ import annotations

import math

import torch

from .. import shift_ops


def _device_of(t: torch.Tensor):
    if t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("deepinpainting_b200 needs a CUDA device: there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _require_binary(m: torch.Tensor, name: str) -> None:
    """The reference box-filters / sums ``mask.float()`` as given; the device kernels work on 0/1 masks in exact integer
    arithmetic.  Soft masks would silently give other flags, so they are refused (one host sync per NEW mask)."""
    if m.dtype == torch.bool:
        return
    if bool(((m != 0) & (m != 1)).any()):
        raise ValueError("%s must hold only 0 / 1 values (soft masks are not supported by the device mask kernels)" % name)


def cal_feat_mask(inMask, conv_layers, threshold):
    """util/util.py:68-84.  inMask [1,1,S,S] (bool / byte / float) -> ByteTensor [1,1,S>>L,S>>L]
    holding 1 where the L-times box-filtered mask exceeds ``threshold``."""
    assert inMask.dim() == 4, "mask must be 4 dimensions"
    assert inMask.size(0) == 1, "the first dimension must be 1 for mask"
    dev = _device_of(inMask)
    m2 = inMask.to(dev)[0, 0]
    _require_binary(m2, "inMask")
    out = shift_ops.feat_mask(m2, conv_layers, threshold)
    return out.view(1, 1, out.size(0), out.size(1))


def flatten_offsets_from_flag(flag):
    """util/util.py:150-157 in closed form (pure index arithmetic, any device).  The reference
    writes ``flatten_offsets_all[i + ov_i] = -ov_i`` for i = 0..N-1 with ov_i = -(masked positions
    before i); later writes win, so entry j < U (U unmasked positions) ends as (j-th unmasked
    position) - j, a trailing masked run leaves M-1 at entry U, and the rest stays 0."""
    flag = flag.long()
    N = flag.numel()
    out = torch.zeros(N, dtype=torch.int64, device=flag.device)
    unmasked = torch.nonzero(flag == 0, as_tuple=False).flatten()
    U = int(unmasked.numel())
    M = N - U
    if U:
        out[:U] = unmasked - torch.arange(U, dtype=torch.int64, device=flag.device)
    if M and U < N and int(flag[N - 1]) == 1:
        out[U] = M - 1
    return out


def cal_mask_given_mask_thred(img, mask, patch_size, stride, mask_thred):
    """util/util.py:88-161.  Returns (flag, nonmask_point_idx, flatten_offsets, mask_point_idx), all
    int64 like the reference's, on the mask's device.  ``nonmask_point_idx`` is every position
    (:137-139); ``flatten_offsets`` reproduces :150-157 (it is unused by the operator)."""
    assert img.dim() == 3, 'img has to be 3 dimenison!'
    assert mask.dim() == 2, 'mask has to be 2 dimenison!'
    dev = _device_of(mask)
    _require_binary(mask, "mask")
    m8 = (mask.to(dev) != 0).to(torch.uint8).contiguous()
    H, W = img.size(1), img.size(2)
    if (m8.size(0), m8.size(1)) != (H, W):
        raise ValueError("mask %s does not match the feature map %dx%d" % (tuple(m8.shape), H, W))
    # the window sums are integers: sum >= mask_thred  <=>  sum >= ceil(mask_thred)   (util/util.py:113-118)
    mi = shift_ops.build_flags(m8, patch_size, stride, int(math.ceil(mask_thred)))
    N = mi.flag.numel()
    flag = mi.flag.long()
    shift_ops.register_mask_index(flag, mi)       # lets IPSRFunction.apply reuse the device vectors
    mask_point_idx = mi.mask_idx.long()
    nonmask_point_idx = torch.arange(N, dtype=torch.int64, device=dev)
    flatten_offsets = flatten_offsets_from_flag(flag)
    return flag, nonmask_point_idx, flatten_offsets, mask_point_idx


def cal_sps_for_Advanced_Indexing(h, w):
    """util/util.py:166-174."""
    sp_y = torch.arange(0, w).long().repeat(h)
    sp_x = torch.arange(0, h).long().repeat_interleave(w)
    return sp_x, sp_y
