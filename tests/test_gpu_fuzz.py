"""Randomised parity sweep: small random shapes, masks, batch sizes and triple weights through the module API against
the numpy oracle (well-conditioned, non-negative inputs: indices, outputs and gradients are all comparable).  Fixed
seeds, so a failure reproduces."""
import collections

import numpy as np
import pytest
import torch

from oracle import ipsr_oracle as O

pytestmark = pytest.mark.gpu
Ref = collections.namedtuple("Ref", ["relu4_3"])
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (there is no CPU path to fall back to)")


def _random_mask(rng, S):
    m = np.zeros((S, S), bool)
    for _ in range(int(rng.integers(1, 5))):
        y, x = rng.integers(0, S - 8, 2)
        h, w = rng.integers(4, max(5, S // 2), 2)
        m[y:y + h, x:x + w] = True
    if rng.random() < 0.3:
        m |= rng.random((S, S)) < 0.05
    return m[None, None]


@pytest.mark.parametrize("seed", range(24))
def test_random_case_against_oracle(seed):
    from deepinpainting_b200 import shift_ops
    from deepinpainting_b200.models import IPSR_model
    rng = np.random.default_rng(4000 + seed)
    H = int(rng.choice([4, 6, 8, 12, 16, 32]))
    C = int(rng.choice([32, 64, 96, 128, 256]))
    B = int(rng.integers(1, 4))
    tw = float(rng.choice([1.0, 0.5, 2.0]))
    mode = str(rng.choice(["auto", "exact", "auto"]))
    S = H * 8
    mg = _random_mask(rng, S)
    x = np.abs(rng.standard_normal((B, C, H, H))).astype(np.float32) + 0.05
    ref = np.abs(rng.standard_normal((B, C, H, H))).astype(np.float32)
    g = rng.standard_normal((B, C, H, H)).astype(np.float32)
    if rng.random() < 0.25:                      # duplicate bank patches: exact ties, the lowest index must win
        x[:, :, 0, 1] = x[:, :, 0, 0]
        x[:, :, H - 1, H - 1] = x[:, :, 0, 0]
    if rng.random() < 0.2:                       # a zero patch: all-zero filter, score 0
        x[:, :, H // 2, H // 2] = 0
    old = shift_ops.config["correlation_mode"]
    shift_ops.config["correlation_mode"] = mode
    try:
        m = IPSR_model(5 / 16.0, 1, 1, 1, 1, tw)
        m.set_mask(torch.from_numpy(mg).to(DEV), 3, 5 / 16.0)
        m.set_ref(Ref(torch.from_numpy(ref).to(DEV)))
        xt = torch.from_numpy(x).to(DEV).requires_grad_(True)
        y = m(xt)
        y.backward(torch.from_numpy(g).to(DEV))
        torch.cuda.synchronize()
    finally:
        shift_ops.config["correlation_mode"] = old
    fm = O.cal_feat_mask(mg, 3, 5 / 16.0)[0, 0]
    flag = O.cal_mask_given_mask_thred((C, H, H), fm, 1, 1, 1)[0]
    np.testing.assert_array_equal(m.flag.cpu().numpy(), flag)
    o32 = O.shift_forward(x, ref, flag, np.float32)
    o64 = O.shift_forward(x, ref, flag, np.float64, keep_attn=False)
    ind = y.grad_fn.saved_shift.ind.cpu().numpy().astype(np.int64)
    safe = o64.gap > 1e-4
    np.testing.assert_array_equal(ind[safe], o64.ind[safe])
    tie = o64.gap == 0                           # exact ties (duplicated patches): first index, like torch.max
    np.testing.assert_array_equal(ind[tie], o64.ind[tie])
    if not (ind == o32.ind).all():
        return
    out = y.detach().cpu().numpy()
    assert np.abs(out - o32.out).max() <= 1e-4 * max(np.abs(o32.out).max(), 1e-6)
    gin = O.shift_backward(g, o32.attn_trunc, tw)
    midx = np.nonzero(flag)[0]
    N = H * H
    keep = np.ones((B, C, H, H), bool)
    if len(midx) > 1:                            # columns whose truncated weight sits on a rounding knife-edge
        a = np.abs(o32.attn[:, midx[1:], :])
        bad = ((a > 0.5) & (np.abs(a - np.round(a)) < 1e-4)).any(axis=1)
        keep = ~np.broadcast_to(bad[:, None, :], (B, C, N)).reshape(B, C, H, H)
    assert np.abs(xt.grad.cpu().numpy() - gin)[keep].max() <= 1e-4 * np.abs(gin).max()
