"""Host networks for the full generator step (BASELINE.json configs[4]).

The reference's generator is stock torch.nn around the shift layer: a rough U-Net ``netP`` (models/networks.py:371-452,
``unet_256``), the refinement U-Net ``netG`` that hosts the shift layer, InnerCos and InnerCos2 at its 32 x 32 level
(models/networks.py:187-366, ``unet_ipsr``), and a VGG-16 whose relu4_3 features of the reference image guide the shift
(models/vgg16.py:6-36, models/IPSR.py:162-164).  None of that is a custom kernel in the reference and none is here: the
convolutions are cuDNN's, run in bf16 (autocast, channels_last) -- what matters is that the B200 shift layer drops into
them.  The module trees below have the reference's layer ORDER, so their ``state_dict`` keys are the reference's and its
checkpoints (``<epoch>_net_G.pt`` / ``_net_P.pt``, models/base_model.py:43-64) load unchanged
(tests/test_gpu_generator.py loads the reference generator's own weights and compares outputs).

``GeneratorStep`` is the training iteration of models/IPSR.py:120-267 restricted to the generator:
set_input (mean-colour fill, :148-150) -> set_latent_mask -> set_ref_latent -> set_gt_latent -> netP -> composite -> netG ->
L1 losses + the (detached) InnerCos losses -> backward -> Adam.  The discriminators and the GAN loss are out of scope.
"""
from __future__ import annotations

import collections
import functools

import torch
import torch.nn as nn
import torch.nn.functional as F

from .models import IPSR_model, InnerCos, InnerCos2

VggOutputs = collections.namedtuple("VggOutputs", ["relu1_2", "relu2_2", "relu3_3", "relu4_3"])


class ShiftOptions:
    """The hot-path fields of the reference's option object (app.py:1-60, train.ipynb cell 0)."""
    threshold = 5 / 16.0
    fixed_mask = 1
    shift_sz = 1
    stride = 1
    mask_thred = 1
    triple_weight = 1
    strength = 1
    skip = 0


def _norm(kind: str):
    if kind == "instance":
        return functools.partial(nn.InstanceNorm2d, affine=True)
    if kind == "batch":
        return functools.partial(nn.BatchNorm2d, affine=True)
    raise NotImplementedError("normalization layer [%s] is not found" % kind)


class _SkipLevel(nn.Module):
    """One level of a U-Net built inside out: ``model`` is down-path + inner level + up-path, and every level but the
    outermost returns its output concatenated to its input.  ``layers`` fixes the order (and with it the state_dict keys)."""

    def __init__(self, layers, outermost):
        super().__init__()
        self.outermost = outermost
        self.model = nn.Sequential(*layers)

    def forward(self, x):
        y = self.model(x)
        if self.outermost:
            return y
        if y.shape[-2:] != x.shape[-2:]:
            y = F.interpolate(y, size=x.shape[-2:], mode="bilinear", align_corners=False)
        return torch.cat([y, x], 1)


def _plain_level(outer, inner, in_ch=None, sub=None, outermost=False, innermost=False, norm=None, dropout=False):
    """Level of the rough U-Net (4x4 stride-2 convolutions)."""
    in_ch = outer if in_ch is None else in_ch
    down = nn.Conv2d(in_ch, inner, 4, 2, 1)
    if outermost:
        layers = [down, sub, nn.ReLU(True), nn.ConvTranspose2d(inner * 2, outer, 4, 2, 1), nn.Tanh()]
    elif innermost:
        layers = [nn.LeakyReLU(0.2, True), down, nn.ReLU(True), nn.ConvTranspose2d(inner, outer, 4, 2, 1), norm(outer)]
    else:
        layers = [nn.LeakyReLU(0.2, True), down, norm(inner), sub, nn.ReLU(True), nn.ConvTranspose2d(inner * 2, outer, 4, 2, 1),
                  norm(outer)]
        if dropout:
            layers.append(nn.Dropout(0.5))
    return _SkipLevel(layers, outermost)


def _refine_level(outer, inner, in_ch=None, sub=None, outermost=False, innermost=False, norm=None, dropout=False, shift=None):
    """Level of the refinement U-Net: a dilated 4x4 stride-2 convolution, then a 3x3 one; ``shift`` = (layer, cos, cos2)
    places the shift layer and its two side losses where the reference does."""
    in_ch = outer if in_ch is None else in_ch
    if outermost:
        layers = [nn.Conv2d(in_ch, inner, 3, 1, 1), sub, nn.ReLU(True), nn.ConvTranspose2d(inner * 2, outer, 3, 1, 1)]
        return _SkipLevel(layers, True)
    strided = nn.Conv2d(in_ch, in_ch, 4, 2, 3, dilation=2)
    if innermost:
        layers = [nn.LeakyReLU(0.2, True), strided, nn.ReLU(True), nn.ConvTranspose2d(inner, outer, 4, 2, 1), norm(outer)]
        return _SkipLevel(layers, False)
    head = [nn.LeakyReLU(0.2, True), strided, norm(in_ch), nn.LeakyReLU(0.2, True), nn.Conv2d(in_ch, inner, 3, 1, 1)]
    tail = [nn.ReLU(True), nn.ConvTranspose2d(inner * 2, outer, 3, 1, 1), norm(outer), nn.ReLU(True),
            nn.ConvTranspose2d(outer, outer, 4, 2, 1), norm(outer)]
    if shift is None:
        layers = head + [norm(inner), sub] + tail
    else:
        layer, cos, cos2 = shift
        layers = head + [layer, cos, norm(inner), sub, cos2] + tail
    if dropout:
        layers.append(nn.Dropout(0.5))
    return _SkipLevel(layers, False)


class UnetGenerator(nn.Module):
    """The rough network netP (``unet_256``): eight 4x4 stride-2 levels."""

    def __init__(self, input_nc=3, output_nc=3, num_downs=8, ngf=64, norm="instance", use_dropout=False):
        super().__init__()
        n = _norm(norm)
        level = _plain_level(ngf * 8, ngf * 8, innermost=True, norm=n)
        for _ in range(num_downs - 5):
            level = _plain_level(ngf * 8, ngf * 8, sub=level, norm=n, dropout=use_dropout)
        level = _plain_level(ngf * 4, ngf * 8, sub=level, norm=n)
        level = _plain_level(ngf * 2, ngf * 4, sub=level, norm=n)
        level = _plain_level(ngf, ngf * 2, sub=level, norm=n)
        self.model = _plain_level(output_nc, ngf, in_ch=input_nc, sub=level, outermost=True, norm=n)

    def forward(self, x):
        return self.model(x)


class UnetGeneratorIPSR(nn.Module):
    """The refinement network netG (``unet_ipsr``) with the shift layer, InnerCos and InnerCos2 at the 32 x 32 level.
    ``shift_layers`` / ``cos_layers`` / ``cos2_layers`` are the caller-owned lists through which the training model reaches
    those modules (models/IPSR.py:51)."""

    def __init__(self, input_nc=6, output_nc=3, num_downs=8, opt=ShiftOptions, mask_global=None, ngf=64, norm="instance",
                 use_dropout=False):
        super().__init__()
        n = _norm(norm)
        self.shift_layers, self.cos_layers, self.cos2_layers = [], [], []
        level = _refine_level(ngf * 8, ngf * 8, innermost=True, norm=n)
        for _ in range(num_downs - 5):
            level = _refine_level(ngf * 8, ngf * 8, sub=level, norm=n, dropout=use_dropout)
        level = _refine_level(ngf * 8, ngf * 8, sub=level, norm=n, dropout=use_dropout)
        layer = IPSR_model(opt.threshold, opt.fixed_mask, opt.shift_sz, opt.stride, opt.mask_thred, opt.triple_weight)
        cos = InnerCos(strength=opt.strength, skip=opt.skip)
        cos2 = InnerCos2(strength=opt.strength, skip=opt.skip)
        if mask_global is not None:
            layer.set_mask(mask_global, 3, opt.threshold)
            cos.set_mask(mask_global, opt)
            cos2.set_mask(mask_global, opt)
        self.shift_layers.append(layer)
        self.cos_layers.append(cos)
        self.cos2_layers.append(cos2)
        level = _refine_level(ngf * 4, ngf * 8, sub=level, norm=n, shift=(layer, cos, cos2))
        level = _refine_level(ngf * 2, ngf * 4, sub=level, norm=n)
        level = _refine_level(ngf, ngf * 2, sub=level, norm=n)
        self.model = _refine_level(output_nc, ngf, in_ch=input_nc, sub=level, outermost=True, norm=n)

    def forward(self, x):
        return self.model(x)


class Vgg16Features(nn.Module):
    """VGG-16 up to relu4_3 in the reference's four slices (models/vgg16.py:11-22).  The reference downloads torchvision's
    pretrained weights; there is no network here, so the weights are whatever the caller loads (random for benchmarks).
    ``forward`` caches its last result per input tensor: the reference runs VGG on the same ground-truth batch twice per
    iteration (models/IPSR.py:187,213)."""

    CFG = ((64, 64), (128, 128), (256, 256, 256), (512, 512, 512))

    def __init__(self):
        super().__init__()
        slices, idx, cin = [], 0, 3
        for si, widths in enumerate(self.CFG):
            seq = nn.Sequential()
            for w in widths:
                seq.add_module(str(idx), nn.Conv2d(cin, w, 3, 1, 1))
                seq.add_module(str(idx + 1), nn.ReLU(True))
                idx += 2
                cin = w
            if si < 3:                                      # torchvision's indices 4, 9, 16: the pooling CLOSES slices 1-3,
                seq.add_module(str(idx), nn.MaxPool2d(2, 2))   # so relu4_3 sits at 1/8 resolution (32 x 32 for 256^2 images)
                idx += 1
            slices.append(seq)
        self.slice1, self.slice2, self.slice3, self.slice4 = slices
        for p in self.parameters():
            p.requires_grad = False
        self._cache_key, self._cache_val = None, None

    def forward(self, x):
        key = (id(x), x._version, x.data_ptr(), tuple(x.shape))
        if key == self._cache_key:
            return self._cache_val
        h1 = self.slice1(x)
        h2 = self.slice2(h1)
        h3 = self.slice3(h2)
        h4 = self.slice4(h3)
        out = VggOutputs(h1, h2, h3, h4)
        self._cache_key, self._cache_val = key, out
        return out


class GeneratorStep:
    """One generator training iteration (models/IPSR.py:120-267 without the discriminators): bf16 autocast around the
    convolutions, channels_last activations, per-sample free-form masks, Adam.  ``ddp=True`` wraps the two U-Nets in
    DistributedDataParallel (NCCL gradient all-reduce overlapped with the backward; the shift layer itself needs no
    collective: images are independent)."""

    def __init__(self, device, opt=ShiftOptions, ngf=64, bf16=True, ddp=False, lr=2e-4, lambda_a=100.0, seed=0):
        self.dev = torch.device(device)
        self.opt, self.bf16, self.lambda_a = opt, bf16, lambda_a
        torch.manual_seed(seed)
        self.netP = UnetGenerator(3, 3, 8, ngf).to(self.dev).to(memory_format=torch.channels_last)
        self.netG = UnetGeneratorIPSR(6, 3, 8, opt, None, ngf).to(self.dev).to(memory_format=torch.channels_last)
        self.vgg = Vgg16Features().to(self.dev).to(memory_format=torch.channels_last).eval()
        for net in (self.netP, self.netG):
            for m in net.modules():
                if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                    nn.init.normal_(m.weight, 0.0, 0.02)
                    if m.bias is not None:
                        nn.init.zeros_(m.bias)
        self.shift_layers, self.cos_layers, self.cos2_layers = self.netG.shift_layers, self.netG.cos_layers, self.netG.cos2_layers
        self.runP, self.runG = self.netP, self.netG
        if ddp:
            from torch.nn.parallel import DistributedDataParallel as DDP
            self.runP = DDP(self.netP, device_ids=[self.dev.index], gradient_as_bucket_view=True)
            self.runG = DDP(self.netG, device_ids=[self.dev.index], gradient_as_bucket_view=True)
        params = list(self.netP.parameters()) + list(self.netG.parameters())
        self.optim = torch.optim.Adam(params, lr=lr, betas=(0.5, 0.999), fused=True)
        self.loss = None

    def _autocast(self):
        return torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.bf16)

    def set_input(self, image, mask, ref):
        """image, ref [B,3,S,S] in [-1,1]; mask [B,1,S,S] bool, True = hole (one free-form mask per sample)."""
        self.real = image
        self.mask = mask
        self.ref_img = ref
        fill = torch.tensor([2 * 123.0 / 255.0 - 1.0, 2 * 104.0 / 255.0 - 1.0, 2 * 117.0 / 255.0 - 1.0], device=image.device).view(1, 3, 1, 1)
        self.input = torch.where(mask, fill.to(image.dtype), image).contiguous(memory_format=torch.channels_last)
        for m in self.shift_layers:
            m.set_mask(mask, 3, self.opt.threshold)
        feat_mask = self.shift_layers[0].masks[0] if getattr(self.shift_layers[0], "masks", None) else self.shift_layers[0].mask
        for m in self.cos_layers + self.cos2_layers:
            # InnerCos takes ONE mask (the reference's batch shares it, models/InnerCos.py:16-21); with per-sample masks
            # the side loss uses the first sample's, like the reference would with its [1,1,S,S] mask_global
            m.mask = feat_mask.float()
        with torch.no_grad(), self._autocast():
            ref_feat = self.vgg(self.ref_img.contiguous(memory_format=torch.channels_last))
            gt_feat = self.vgg(self.real.contiguous(memory_format=torch.channels_last))
        for m in self.shift_layers:
            m.set_ref(VggOutputs(None, None, None, ref_feat.relu4_3.float().contiguous()))
        target = gt_feat.relu4_3.float().contiguous()
        for m in self.cos_layers + self.cos2_layers:
            m.set_target(target)

    def forward(self):
        with self._autocast():
            self.fake_p = self.runP(self.input)
            un = self.fake_p.float()
            self.middle = torch.where(self.mask, un, self.real)
            x6 = torch.cat([self.middle, self.input], 1).contiguous(memory_format=torch.channels_last)
            self.fake_b = self.runG(x6)
        return self.fake_b

    def optimize_parameters(self):
        self.optim.zero_grad(set_to_none=True)
        self.forward()
        l1 = (F.l1_loss(self.fake_b.float(), self.real) + F.l1_loss(self.fake_p.float(), self.real)) * self.lambda_a
        side = sum(m.loss.detach() for m in self.cos_layers + self.cos2_layers)     # models/IPSR.py:256-263: detached values
        self.loss = l1 + side
        self.loss.backward()
        self.optim.step()
        return self.loss
