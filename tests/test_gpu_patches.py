"""GPU parity tests for shift_sz = k > 1 / stride = s > 1 (forward only; BASELINE.json configs[3]: 3 x 3 patches).

The reference computes the whole output for these settings (models/IPSRFunction.py:46-133) and then fails storing the
attention for its backward (:134).  The golden fixtures tests/golden/k*.npz hold that output, produced by the reference
itself (oracle/make_golden.py run_patch_case).  Tolerances as in test_gpu_parity.py: arg-max indices bit-exact wherever
the fp64 top-2 gap exceeds 1e-4, pasted features within 1e-4 relative on the well-conditioned inputs.
"""
import collections

import numpy as np
import pytest
import torch

from conftest import PATCH_CASES, golden_path
from oracle import ipsr_oracle as O

pytestmark = pytest.mark.gpu

Ref = collections.namedtuple("Ref", ["relu4_3"])
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (there is no CPU path to fall back to)")
    from deepinpainting_b200 import _lib
    _lib.load()


def _unpack_mask(z):
    S = int(z["mask_size"])
    return np.unpackbits(z["mask_global"])[: S * S].reshape(1, 1, S, S).astype(bool)


def _run_module(x, ref, mask_global, k, s, thr, mode="auto"):
    """The layer driven as models/networks.py + models/IPSR.py drive it, with opt.shift_sz = k, opt.stride = s."""
    from deepinpainting_b200 import shift_ops
    from deepinpainting_b200.models import IPSR_model
    old = shift_ops.config["correlation_mode"]
    shift_ops.config["correlation_mode"] = mode
    try:
        m = IPSR_model(5 / 16.0, 1, k, s, thr, 1.0)
        m.set_mask(torch.from_numpy(mask_global).to(DEV), 3, 5 / 16.0)
        m.set_ref(Ref(torch.from_numpy(ref).to(DEV)))
        xt = torch.from_numpy(x).to(DEV).requires_grad_(True)
        y = m(xt)
        torch.cuda.synchronize()
        return y, m, xt
    finally:
        shift_ops.config["correlation_mode"] = old


def _gap64(x, ref, k, s):
    """fp64 top-2 score gap per query patch (oracle arithmetic)."""
    B = x.shape[0]
    gaps = []
    for b in range(B):
        pat = O.extract_patches(x[b].astype(np.float64), k, s)
        patn = O.l2_normalize_patches(pat).reshape(pat.shape[0], -1)
        rp = O.extract_patches(ref[b].astype(np.float64), k, s).reshape(pat.shape[0], -1)
        S = rp @ patn.T
        part = np.partition(S, S.shape[1] - 2, axis=1)
        gaps.append(part[:, -1] - part[:, -2])
    return np.stack(gaps)


@pytest.mark.parametrize("mode", ["auto", "exact"])
@pytest.mark.parametrize("name", PATCH_CASES)
def test_patch_forward_matches_reference_golden(name, mode):
    z = np.load(golden_path(name))
    x, ref = z["x"], z["ref"]
    k, s, thr = int(z["patch"]), int(z["stride"]), int(z["mask_thred"])
    y, m, _ = _run_module(x, ref, _unpack_mask(z), k, s, thr, mode)
    np.testing.assert_array_equal(m.flag.cpu().numpy(), z["flag"])
    np.testing.assert_array_equal(m.mask_point_idx.cpu().numpy(), z["mask_point_idx"])
    ind = y.grad_fn.ind_lst.cpu().numpy().astype(np.int64)
    safe = _gap64(x, ref, k, s) > 1e-4
    assert safe.mean() > 0.9
    np.testing.assert_array_equal(ind[safe], z["ind"][safe])
    out = y.detach().cpu().numpy()
    assert out.shape == z["out"].shape
    if (ind == z["ind"]).all():
        assert np.abs(out - z["out"]).max() <= 1e-4 * np.abs(z["out"]).max()


def test_patch_backward_is_undefined_like_the_reference():
    z = np.load(golden_path(PATCH_CASES[0]))
    y, _, xt = _run_module(z["x"], z["ref"], _unpack_mask(z), int(z["patch"]), int(z["stride"]), int(z["mask_thred"]))
    with pytest.raises(NotImplementedError):
        y.backward(torch.ones_like(y))


def test_patch_geometry_that_does_not_tile_fails_loudly():
    """(H - k) % s != 0: the reference's conv-transpose output is smaller than the input and :133 raises."""
    from deepinpainting_b200 import shift_ops
    x = torch.rand(1, 32, 8, 8, device=DEV)
    mi = shift_ops.build_flags(torch.zeros(8, 8, dtype=torch.uint8, device=DEV), 3, 2, 1)
    with pytest.raises(RuntimeError):
        shift_ops.shift_forward_patches(x, x, mi, 3, 2)


@pytest.mark.parametrize("B,C,H,k,s", [(2, 128, 12, 3, 1),      # K = 1152 > 1024: wide rows (2 values per thread)
                                        (1, 160, 10, 3, 1),      # K = 1440
                                        (1, 96, 12, 4, 2),       # K = 1536, stride 2
                                        (1, 512, 8, 3, 1)])      # K = 4608 (the model's real width, 8 values per thread)
def test_wide_patch_rows_against_oracle(B, C, H, k, s):
    rng = np.random.default_rng(C + H + k)
    x = np.abs(rng.standard_normal((B, C, H, H))).astype(np.float32)
    ref = np.abs(rng.standard_normal((B, C, H, H))).astype(np.float32)
    S = H * 8
    mg = np.zeros((1, 1, S, S), bool)
    mg[:, :, S // 3:2 * S // 3, S // 4:3 * S // 4] = True
    fm = O.cal_feat_mask(mg, 3, 5 / 16.0)[0, 0]
    y, m, _ = _run_module(x, ref, mg, k, s, 1)
    o_out, o_ind = O.shift_forward_patches(x, ref, fm, k, s, 1)
    assert 1 < int(m.flag.sum()) < m.flag.numel()
    ind = y.grad_fn.ind_lst.cpu().numpy().astype(np.int64)
    safe = _gap64(x, ref, k, s) > 1e-4
    np.testing.assert_array_equal(ind[safe], o_ind[safe])
    if (ind == o_ind).all():
        assert np.abs(y.detach().cpu().numpy() - o_out).max() <= 1e-4 * np.abs(o_out).max()


def test_narrow_and_wide_routes_agree():
    """The same problem through the 1 x 1 pipeline on patch maps (K <= 1024) and through the wide kernels."""
    from deepinpainting_b200 import shift_ops
    rng = np.random.default_rng(5)
    B, C, H, k = 2, 64, 16, 3
    x = torch.from_numpy(np.abs(rng.standard_normal((B, C, H, H))).astype(np.float32)).to(DEV)
    ref = torch.from_numpy(np.abs(rng.standard_normal((B, C, H, H))).astype(np.float32)).to(DEV)
    feat = torch.zeros(H, H, dtype=torch.uint8, device=DEV)
    feat[5:11, 4:12] = 1
    mi = shift_ops.build_flags(feat, k, 1, 1)
    out_a, ind_a = shift_ops.shift_forward_patches(x, ref, mi, k, 1, mode="exact")
    old = shift_ops.PATCH_ROW_LIMIT
    shift_ops.PATCH_ROW_LIMIT = 0
    try:
        out_b, ind_b = shift_ops.shift_forward_patches(x, ref, mi, k, 1)
    finally:
        shift_ops.PATCH_ROW_LIMIT = old
    torch.cuda.synchronize()
    assert torch.equal(ind_a, ind_b)
    assert float((out_a - out_b).abs().max()) <= 1e-5 * float(out_a.abs().max())


def test_tensor_path_on_patch_maps():
    """k = 2, stride = 2 on a 32 x 32 map: P = 256 patch positions, K = 256 -> the tcgen05 path runs on the patch maps."""
    from deepinpainting_b200 import _lib, shift_ops
    rng = np.random.default_rng(9)
    B, C, H, k, s = 2, 64, 32, 2, 2
    assert _lib.load().ipsr_tensor_path_supported(C * k * k, (H // 2) ** 2) == 1
    x = np.abs(rng.standard_normal((B, C, H, H))).astype(np.float32)
    ref = np.abs(rng.standard_normal((B, C, H, H))).astype(np.float32)
    fm = np.zeros((H, H), np.uint8)
    fm[8:20, 10:26] = 1
    mi = shift_ops.build_flags(torch.from_numpy(fm).to(DEV), k, s, 1)
    out, ind = shift_ops.shift_forward_patches(torch.from_numpy(x).to(DEV), torch.from_numpy(ref).to(DEV), mi, k, s, mode="tensor")
    torch.cuda.synchronize()
    o_out, o_ind = O.shift_forward_patches(x, ref, fm, k, s, 1)
    ind = ind.cpu().numpy().astype(np.int64)
    safe = _gap64(x, ref, k, s) > 1e-4
    np.testing.assert_array_equal(ind[safe], o_ind[safe])
    if (ind == o_ind).all():
        assert np.abs(out.cpu().numpy() - o_out).max() <= 1e-4 * np.abs(o_out).max()


def test_bank_sharded_patches_single_process():
    """configs[3]: the patch bank split in column shards, the (max, idx) keys merged by MAX (what the NCCL
    all-reduce does), for both routes."""
    from deepinpainting_b200 import shift_ops
    rng = np.random.default_rng(11)
    for (B, C, H, k) in [(1, 64, 16, 3), (1, 128, 12, 3)]:
        x = torch.from_numpy(np.abs(rng.standard_normal((B, C, H, H))).astype(np.float32)).to(DEV)
        ref = torch.from_numpy(np.abs(rng.standard_normal((B, C, H, H))).astype(np.float32)).to(DEV)
        feat = torch.zeros(H, H, dtype=torch.uint8, device=DEV)
        feat[4:9, 3:10] = 1
        mi = shift_ops.build_flags(feat, k, 1, 1)
        P = mi.flag.numel()
        full, ind_full = shift_ops.shift_forward_patches(x, ref, mi, k, 1, mode="exact")
        cuts = [0, P // 3, 2 * P // 3, P]
        keys = []

        def grab(t, keys=keys):
            keys.append(t.clone())

        for r in range(3):
            shift_ops.shift_forward_patches(x, ref, mi, k, 1, mode="exact", col_begin=cuts[r], col_end=cuts[r + 1], reduce_max=grab)
        merged = torch.stack(keys).max(dim=0).values

        def put(t, merged=merged):
            t.copy_(merged)

        out, ind = shift_ops.shift_forward_patches(x, ref, mi, k, 1, mode="exact", col_begin=cuts[0], col_end=cuts[1], reduce_max=put)
        torch.cuda.synchronize()
        assert torch.equal(ind, ind_full)
        assert torch.equal(out, full)


def test_wide_patch_tensor_route_matches_exact_route_and_shards():
    """Long patch rows (K = 1152 > 1024) on the tcgen05 path -- P = 324 positions padded to 384, the padding columns masked
    in the GEMM epilogue -- against the exact fp32 route, unsharded and as three 128-column bank shards merged by MAX."""
    from deepinpainting_b200 import shift_ops
    rng = np.random.default_rng(21)
    B, C, H, k = 2, 128, 20, 3
    xs = torch.from_numpy(rng.standard_normal((B, C, H, H)).astype(np.float32)).to(DEV)
    ref = torch.from_numpy(np.abs(rng.standard_normal((B, C, H, H))).astype(np.float32)).to(DEV)
    feat = torch.zeros(H, H, dtype=torch.uint8, device=DEV)
    feat[5:12, 4:13] = 1
    mi = shift_ops.build_flags(feat, k, 1, 1)
    P = mi.flag.numel()
    assert P == 324 and C * k * k > shift_ops.PATCH_ROW_LIMIT
    # signed input, NEGATED reference: every true score is negative, so an unmasked padding column (score 0) would win
    # everywhere.  (The blend is chaotic on signed data: indices only.)
    _, ind_e = shift_ops.shift_forward_patches(xs.abs(), -ref, mi, k, 1, mode="exact")
    _, ind_t = shift_ops.shift_forward_patches(xs.abs(), -ref, mi, k, 1, mode="tensor")
    gap = _gap64(xs.abs().cpu().numpy(), -ref.cpu().numpy(), k, 1)
    safe = torch.from_numpy(gap > 1e-4)
    assert safe.float().mean() > 0.9 and int(ind_t.max()) < P
    assert torch.equal(ind_t.cpu()[safe], ind_e.cpu()[safe])
    x = xs.abs()                                                                                # well conditioned from here on
    out_e, ind_e = shift_ops.shift_forward_patches(x, ref, mi, k, 1, mode="exact")
    out_t, ind_t = shift_ops.shift_forward_patches(x, ref, mi, k, 1, mode="tensor")
    torch.cuda.synchronize()
    gap = _gap64(x.cpu().numpy(), ref.cpu().numpy(), k, 1)
    safe = torch.from_numpy(gap > 1e-4)
    assert safe.float().mean() > 0.9
    assert torch.equal(ind_t.cpu()[safe], ind_e.cpu()[safe])
    if torch.equal(ind_t, ind_e):
        assert float((out_t - out_e).abs().max()) <= 1e-4 * float(out_e.abs().max())
    cuts = [0, 128, 256, P]
    keys = []

    def grab(t, keys=keys):
        keys.append(t.clone())

    for r in range(3):
        shift_ops.shift_forward_patches(x, ref, mi, k, 1, mode="tensor", col_begin=cuts[r], col_end=cuts[r + 1], reduce_max=grab)
    merged = torch.stack(keys).max(dim=0).values

    def put(t, merged=merged):
        t.copy_(merged)

    out_s, ind_s = shift_ops.shift_forward_patches(x, ref, mi, k, 1, mode="tensor", col_begin=0, col_end=128, reduce_max=put)
    torch.cuda.synchronize()
    assert torch.equal(ind_s, ind_t)
    assert torch.equal(out_s, out_t)
    # an empty shard (more ranks than 128-column tiles) contributes identity keys
    keys.clear()
    shift_ops.shift_forward_patches(x, ref, mi, k, 1, mode="tensor", col_begin=P, col_end=P, reduce_max=grab)
    assert bool((keys[0] == -(1 << 63)).all())


def test_wide_kernels_stay_inside_their_buffers():
    """compute-sanitizer is not available on the GPU pool: every output buffer of the long-row kernels sits between two
    guard bands here, which must come back untouched (K = 1155: not a multiple of 4 -> the 4-byte cp.async path and the
    scalar tile loads; P = 100: padded to 128 rows; M = 21: a partial block of the scan)."""
    from deepinpainting_b200 import _lib
    L = _lib.load()
    dev = DEV
    gen = torch.Generator().manual_seed(2)
    B, K, P, M = 2, 1155, 100, 21
    GUARD = 4096
    st = torch.cuda.current_stream().cuda_stream

    def guarded(n, dtype, fill):
        buf = torch.full((n + 2 * GUARD,), fill, dtype=dtype, device=dev)
        return buf, buf[GUARD:GUARD + n]

    def intact(buf, n, fill):
        lo, hi = buf[:GUARD], buf[GUARD + n:]
        if buf.dtype.is_floating_point:
            return bool((lo == fill).all() and (hi == fill).all())
        return bool((lo == fill).all() and (hi == fill).all())

    rows = torch.randn(B, P, K, generator=gen).abs().to(dev)
    inv = (1.0 / (rows.norm(dim=2) + 1e-8)).contiguous()
    rmax = rows.abs().amax(dim=2).contiguous()
    rnorm = rows.norm(dim=2).contiguous()
    Kpad, Ppad = -(-K // 64) * 64, -(-P // 128) * 128
    nt = B * (Kpad // 64) * 2 * (Ppad // 128) * 16384
    tb, tiles = guarded(nt, torch.uint8, 0x5A)
    rb, rscale = guarded(B * Ppad, torch.float32, -7.0)
    nb, rnorm_pad = guarded(B * Ppad, torch.float32, -7.0)
    _lib.call("ipsr_patch_tiles", rows.data_ptr(), None, rmax.data_ptr(), rnorm.data_ptr(), 1, B, K, P, tiles.data_ptr(),
              rscale.data_ptr(), rnorm_pad.data_ptr(), st)
    torch.cuda.synchronize()
    assert intact(tb, nt, 0x5A) and intact(rb, B * Ppad, -7.0) and intact(nb, B * Ppad, -7.0)
    assert bool((rscale.view(B, Ppad)[:, P:] == 1.0).all()) and bool((rnorm_pad.view(B, Ppad)[:, P:] == 0.0).all())
    # the blocked blend: y, wn, wo, gram between guards
    ind = torch.randint(0, P, (B, P), generator=gen, dtype=torch.int32).to(dev)
    vmax = (torch.rand(B, P, generator=gen) + 0.5).to(dev)
    midx = torch.arange(10, 10 + M, dtype=torch.int32, device=dev)
    ng = L.ipsr_blend_wide_gram_floats(B, M)
    gb, gram = guarded(ng, torch.float32, -3.0)
    yb, y = guarded(B * M * K, torch.float32, -3.0)
    wb, wn = guarded(B * M, torch.float32, -3.0)
    ob, wo = guarded(B * M, torch.float32, -3.0)
    _lib.call("ipsr_blend_wide_blocked", rows.data_ptr(), inv.data_ptr(), vmax.data_ptr(), ind.data_ptr(), midx.data_ptr(), B, K, P, M,
              gram.data_ptr(), y.data_ptr(), wn.data_ptr(), wo.data_ptr(), st)
    torch.cuda.synchronize()
    assert intact(gb, ng, -3.0) and intact(yb, B * M * K, -3.0) and intact(wb, B * M, -3.0) and intact(ob, B * M, -3.0)
    assert bool(torch.isfinite(y).all()) and bool((y != -3.0).any())
    # ... and it agrees with the one-reduction-per-step kernel
    y2, wn2, wo2 = torch.empty_like(y), torch.empty_like(wn), torch.empty_like(wo)
    _lib.call("ipsr_blend_wide", rows.data_ptr(), inv.data_ptr(), vmax.data_ptr(), ind.data_ptr(), midx.data_ptr(), B, K, P, M,
              y2.data_ptr(), wn2.data_ptr(), wo2.data_ptr(), st)
    torch.cuda.synchronize()
    torch.testing.assert_close(wn, wn2, rtol=2e-5, atol=1e-6)
    torch.testing.assert_close(y, y2, rtol=2e-4, atol=1e-5)
    # winner scores: ind / vmax / keys between guards
    ind_pad = torch.zeros(B, Ppad, dtype=torch.int32, device=dev)
    ind_pad[:, :P] = ind
    ib, ind_o = guarded(B * P, torch.int32, -5)
    vb, vm_o = guarded(B * P, torch.float32, -5.0)
    kb_, keys = guarded(B * P, torch.int64, -5)
    _lib.call("ipsr_patch_winner_scores", rows.data_ptr(), rows.data_ptr(), inv.data_ptr(), ind_pad.data_ptr(), None, B, K, P,
              ind_o.data_ptr(), vm_o.data_ptr(), keys.data_ptr(), st)
    torch.cuda.synchronize()
    assert intact(ib, B * P, -5) and intact(vb, B * P, -5.0) and intact(kb_, B * P, -5)
    want = torch.einsum("bpk,bpk->bp", rows, rows[torch.arange(B)[:, None], ind.long()] * inv[torch.arange(B)[:, None], ind.long()][..., None])
    torch.testing.assert_close(vm_o.view(B, P), want, rtol=2e-5, atol=1e-5)


def test_nonparametricshift_patches_k3():
    from deepinpainting_b200.util.NonparametricShift import NonparametricShift
    rng = np.random.default_rng(3)
    C, H, k = 24, 9, 3
    img = rng.standard_normal((C, H, H)).astype(np.float32)
    P = (H - k + 1) ** 2
    enc_all, enc_nm, dec_all, dec_nm, patches_part, patches_mask = NonparametricShift().buildAutoencoder(
        torch.from_numpy(img).to(DEV), False, False, torch.arange(P), torch.tensor([0, 5, 17]), k, 1)
    pat = O.extract_patches(img, k, 1)
    np.testing.assert_array_equal(patches_part.cpu().numpy(), pat)
    np.testing.assert_array_equal(patches_mask.cpu().numpy(), pat[[0, 5, 17]])
    np.testing.assert_allclose(enc_nm.weight.detach().cpu().numpy(), O.l2_normalize_patches(pat), rtol=3e-6, atol=1e-9)
    np.testing.assert_array_equal(dec_all.weight.detach().cpu().numpy(), pat)
    assert enc_all.kernel_size == (k, k) and dec_nm.kernel_size == (k, k)


def test_config4_full_size_properties():
    """BASELINE.json configs[3] at full size on one GPU: 64 x 64 x 256 features, 3 x 3 patches (P = 3844 patch positions,
    rows of K = 2304), centre hole.  Size-independent checks: the chosen column is the fp64 arg-max on a sample of rows,
    pixels covered only by unmasked patches are the plain overlap-sum of the matched patches, everything is finite."""
    from deepinpainting_b200 import shift_ops
    gen = torch.Generator(device="cpu").manual_seed(1234)
    B, C, H, k = 1, 256, 64, 3
    x = torch.randn(B, C, H, H, generator=gen).abs().to(DEV)
    ref = (torch.relu(torch.randn(B, C, H, H, generator=gen)) * 3).to(DEV)
    feat = torch.zeros(H, H, dtype=torch.uint8, device=DEV)
    feat[H // 4:3 * H // 4, H // 4:3 * H // 4] = 1
    mi = shift_ops.build_flags(feat, k, 1, 1)
    nH = H - k + 1
    P = nH * nH
    assert mi.flag.numel() == P and mi.M == (H // 2 + k - 1) ** 2
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    shift_ops.shift_forward_patches(x, ref, mi, k, 1)          # warm-up
    t0.record()
    out, ind = shift_ops.shift_forward_patches(x, ref, mi, k, 1)
    t1.record()
    torch.cuda.synchronize()
    print("config 4 (1 image, 64x64x256, 3x3 patches) forward: %.2f ms" % t0.elapsed_time(t1))
    assert torch.isfinite(out).all()
    ind = ind.long()
    assert int(ind.min()) >= 0 and int(ind.max()) < P
    rows, inv = shift_ops.patch_rows(x, k, 1)
    rrows, _ = shift_ops.patch_rows(ref, k, 1)
    Xn = (rows[0] * inv[0][:, None]).double()
    sample = torch.arange(0, P, 61, device=DEV)
    S = rrows[0][sample].double() @ Xn.t()
    top2 = torch.topk(S, 2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 1e-4
    assert bool(clear.float().mean() > 0.8)
    assert torch.equal(ind[0][sample][clear], S.argmax(dim=1)[clear])
    # overlap-sum of the matched raw patches (ConvTranspose2d semantics, IPSRFunction.py:131) away from the hole
    pasted = rows[0][ind[0]].view(nH, nH, C, k, k)
    acc = torch.zeros(C, H, H, device=DEV)
    for dy in range(k):
        for dx in range(k):
            acc[:, dy:dy + nH, dx:dx + nH] += pasted[:, :, :, dy, dx].permute(2, 0, 1)
    far = torch.ones(H, H, dtype=torch.bool, device=DEV)
    far[H // 4 - k:3 * H // 4 + k, H // 4 - k:3 * H // 4 + k] = False
    assert float((out[0][:, far] - acc[:, far]).abs().max()) <= 1e-5 * float(acc.abs().max())
