"""Per-image masks in ONE batched call (extension; BASELINE.json configs[4]: free-form masks, batch 32).

Images are independent (models/IPSRFunction.py:46), so a batched call with a flag row per image must reproduce,
image by image, what the shared-mask operator gives for that image alone -- bit for bit in exact mode, where no
launch parameter depends on the batch."""
import numpy as np
import pytest
import torch

from oracle import ipsr_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (there is no CPU path to fall back to)")


def _flags(rng, B, H, kinds):
    flags = np.zeros((B, H * H), np.int64)
    for b, kind in enumerate(kinds):
        f = np.zeros((H, H), np.int64)
        if kind == "empty":
            pass
        elif kind == "full":
            f[:] = 1
        elif kind == "one":
            f[H // 2, H // 3] = 1
        elif kind == "last":
            f[H - 1, H - 1] = 1
        else:
            for _ in range(int(rng.integers(1, 4))):
                y0, x0 = rng.integers(0, H - 2, 2)
                f[y0:y0 + int(rng.integers(1, H // 2 + 1)), x0:x0 + int(rng.integers(1, H // 2 + 1))] = 1
        flags[b] = f.reshape(-1)
    return flags


@pytest.mark.parametrize("C,H,signed", [(64, 16, False), (64, 16, True), (256, 32, True), (512, 16, False)])
def test_batched_ragged_masks_equal_per_image_calls(C, H, signed):
    from deepinpainting_b200 import shift_ops
    rng = np.random.default_rng(C + H + int(signed))
    kinds = ["rand", "empty", "rand", "full", "one", "rand", "last"]
    B = len(kinds)
    flags = _flags(rng, B, H, kinds)
    x = rng.standard_normal((B, C, H, H)).astype(np.float32)
    ref = np.maximum(rng.standard_normal((B, C, H, H)), 0).astype(np.float32) * 3
    if not signed:
        x = np.abs(x)
    g = rng.standard_normal((B, C, H, H)).astype(np.float32)
    xd, rd, gd = (torch.from_numpy(a).to(DEV) for a in (x, ref, g))
    mi = shift_ops.mask_index_from_flag(torch.from_numpy(flags), DEV)
    assert mi.batched and mi.M == int(flags.sum(1).max()) and mi.m_count.tolist() == flags.sum(1).tolist()
    out, saved = shift_ops.shift_forward(xd, rd, mi, need_grad=True, mode="exact")
    gin = shift_ops.shift_backward(gd, saved, 1.5)
    torch.cuda.synchronize()
    nexc = 0
    for b in range(B):
        mib = shift_ops.mask_index_from_flag(torch.from_numpy(flags[b]), DEV)
        ob, sb = shift_ops.shift_forward(xd[b:b + 1].contiguous(), rd[b:b + 1].contiguous(), mib, need_grad=True, mode="exact")
        gb = shift_ops.shift_backward(gd[b:b + 1].contiguous(), sb, 1.5)
        torch.cuda.synchronize()
        assert torch.equal(saved.ind[b:b + 1], sb.ind), b
        assert torch.equal(out[b:b + 1], ob), (b, kinds[b])
        assert torch.equal(gin[b:b + 1], gb), (b, kinds[b])
        Mb = int(flags[b].sum())
        if Mb:
            assert torch.equal(saved.wn[b, :Mb], sb.wn[0, :Mb]) and torch.equal(saved.wo[b, :Mb], sb.wo[0, :Mb])
        if sb.exc_total is not None:
            nexc += int(sb.exc_total.sum())
    if signed:
        assert nexc > 0                                    # blended attention entries did survive the int64 store


def test_batched_ragged_masks_against_oracle_auto_mode():
    """config-5-like: batch 32, 32 x 32 x 512, a different free-form mask per image, tensor path."""
    from deepinpainting_b200 import shift_ops
    rng = np.random.default_rng(55)
    B, C, H = 32, 512, 32
    flags = _flags(rng, B, H, ["rand"] * B)
    x = np.abs(rng.standard_normal((B, C, H, H))).astype(np.float32)
    ref = np.abs(rng.standard_normal((B, C, H, H))).astype(np.float32)
    g = rng.standard_normal((B, C, H, H)).astype(np.float32)
    mi = shift_ops.mask_index_from_flag(torch.from_numpy(flags), DEV)
    out, saved = shift_ops.shift_forward(torch.from_numpy(x).to(DEV), torch.from_numpy(ref).to(DEV), mi, need_grad=True)
    gin = shift_ops.shift_backward(torch.from_numpy(g).to(DEV), saved, 1.0)
    torch.cuda.synchronize()
    out, gin, ind = out.cpu().numpy(), gin.cpu().numpy(), saved.ind.cpu().numpy().astype(np.int64)
    for b in range(0, B, 5):
        o = O.shift_forward(x[b:b + 1], ref[b:b + 1], flags[b], np.float32)
        o64 = O.shift_forward(x[b:b + 1], ref[b:b + 1], flags[b], np.float64, keep_attn=False)
        safe = o64.gap[0] > 1e-4
        np.testing.assert_array_equal(ind[b][safe], o64.ind[0][safe])
        if (ind[b] == o.ind[0]).all():
            assert np.abs(out[b:b + 1] - o.out).max() <= 1e-4 * np.abs(o.out).max()
            gi = O.shift_backward(g[b:b + 1], o.attn_trunc, 1.0)
            midx = np.nonzero(flags[b])[0]
            keep = np.ones((1, C, H, H), bool)
            if len(midx) > 1:
                a = np.abs(o.attn[:, midx[1:], :])
                bad = ((a > 0.5) & (np.abs(a - np.round(a)) < 1e-4)).any(axis=1)
                keep = ~np.broadcast_to(bad[:, None, :], (1, C, H * H)).reshape(1, C, H, H)
            assert np.abs(gin[b:b + 1] - gi)[keep].max() <= 1e-4 * np.abs(gi).max()
