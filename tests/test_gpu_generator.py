"""SURVEY.md 8(f) ranks 3-4 / BASELINE.json configs[4]: the full generator step around the shift layer.

The host networks of deepinpainting_b200.generator are stock torch.nn (cuDNN convolutions, bf16 autocast); what is checked
here is that they ARE the reference's generator -- same state_dict keys, same outputs from the same weights -- and that
the training iteration of models/IPSR.py:120-267 (generator part) runs through them with per-sample free-form masks."""
import numpy as np
import pytest
import torch

import refnet

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _freeform(B, S, seed):
    rng = np.random.default_rng(seed)
    m = np.zeros((B, 1, S, S), bool)
    for b in range(B):
        for _ in range(4):
            y, x = rng.integers(0, S - S // 4, 2)
            h, w = rng.integers(S // 16, S // 3, 2)
            m[b, 0, y:y + h, x:x + w] = True
    return torch.from_numpy(m)


def test_generators_are_the_reference_generators():
    if refnet.staged_networks_path() is None:
        pytest.skip("oracle/_ref not staged: the reference's networks.py did not travel to this box")
    from deepinpainting_b200 import generator as G
    nets = refnet.load_networks_with_dropin()
    torch.manual_seed(0)
    S = 256
    mask = torch.zeros(1, 1, S, S, dtype=torch.bool, device=DEV)
    mask[:, :, 64:192, 64:192] = True
    refG, cosR, cos2R, shiftR = nets.define_G(6, 3, 64, "unet_ipsr", refnet.Opt, mask, "instance", False, "normal", [0], 0.02)
    refP, _, _, _ = nets.define_G(3, 3, 64, "unet_256", refnet.Opt, mask, "instance", False, "normal", [0], 0.02)
    ourG = G.UnetGeneratorIPSR(6, 3, 8, G.ShiftOptions, mask, 64).to(DEV)
    ourP = G.UnetGenerator(3, 3, 8, 64).to(DEV)
    assert list(ourG.state_dict().keys()) == list(refG.state_dict().keys())
    assert list(ourP.state_dict().keys()) == list(refP.state_dict().keys())
    ourG.load_state_dict(refG.state_dict())                    # the reference's checkpoints load unchanged
    ourP.load_state_dict(refP.state_dict())
    ref_feat = torch.relu(torch.randn(2, 512, 32, 32, device=DEV)) * 2
    target = torch.relu(torch.randn(2, 512, 32, 32, device=DEV))
    for layers, cos, cos2 in ((shiftR, cosR, cos2R), (ourG.shift_layers, ourG.cos_layers, ourG.cos2_layers)):
        for m in layers:
            m.set_mask(mask, 3, refnet.Opt.threshold)
            m.set_ref(G.VggOutputs(None, None, None, ref_feat))
        for m in cos + cos2:
            m.set_mask(mask, refnet.Opt)
            m.set_target(target)
    x6 = torch.randn(2, 6, S, S, device=DEV)
    refG.eval(), ourG.eval(), refP.eval(), ourP.eval()
    with torch.no_grad():
        a, b = refG(x6.clone()), ourG(x6.clone())
        c, d = refP(x6[:, :3].clone()), ourP(x6[:, :3].clone())
    # same modules, same weights: equal up to cuDNN's choice of algorithm per module instance
    assert float((a - b).abs().max()) <= 1e-4 * float(a.abs().max()), float((a - b).abs().max())
    assert float((c - d).abs().max()) <= 1e-4 * float(c.abs().max()), float((c - d).abs().max())
    for r, o in ((cosR[0], ourG.cos_layers[0]), (cos2R[0], ourG.cos2_layers[0])):
        assert abs(float(r.loss) - float(o.loss)) <= 1e-4 * abs(float(r.loss))


def test_vgg_slices_follow_torchvision_indices():
    from deepinpainting_b200 import generator as G
    import torchvision
    tv = torchvision.models.vgg16(weights=None).features
    ours = G.Vgg16Features()
    mapped = {}
    for k, v in tv.state_dict().items():
        idx = int(k.split(".")[0])
        if idx > 22:
            continue
        sl = 1 if idx < 5 else 2 if idx < 10 else 3 if idx < 17 else 4
        mapped["slice%d.%s" % (sl, k)] = v
    assert set(mapped) == set(ours.state_dict())
    ours.load_state_dict(mapped)
    x = torch.randn(1, 3, 64, 64)
    want = tv[:23](x)
    got = ours(x)
    assert torch.equal(got.relu4_3, want) and got.relu4_3.shape == (1, 512, 8, 8)
    assert got.relu1_2.shape == (1, 64, 32, 32)                  # the reference's slices end WITH the pooling (vgg16.py:14-21)
    assert ours(x) is got                                       # cached: the reference runs VGG twice on the same batch


def test_generator_training_step_bf16_per_sample_masks():
    from deepinpainting_b200 import generator as G
    B, S = 4, 256
    step = G.GeneratorStep(DEV, bf16=True, ddp=False, seed=1)
    gen = torch.Generator().manual_seed(5)
    img = (torch.rand(B, 3, S, S, generator=gen) * 2 - 1).to(DEV)
    ref = (torch.rand(B, 3, S, S, generator=gen) * 2 - 1).to(DEV)
    mask = _freeform(B, S, 3).to(DEV)
    before = [p.detach().clone() for p in list(step.netG.parameters())[:4]]
    losses = []
    for _ in range(3):
        step.set_input(img, mask, ref)
        losses.append(float(step.optimize_parameters()))
    torch.cuda.synchronize()
    assert all(np.isfinite(losses)), losses
    assert step.fake_b.shape == (B, 3, S, S)
    assert any(not torch.equal(a, b) for a, b in zip(before, list(step.netG.parameters())[:4]))
    layer = step.shift_layers[0]
    assert layer.flag.shape == (B, 32 * 32)                      # one flag row per sample (per-sample masks)
    assert all(float(m.loss) > 0 for m in step.cos_layers + step.cos2_layers)
    assert losses[-1] < losses[0]                                # the same batch three times: the L1 loss goes down
