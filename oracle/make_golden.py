"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container (``/root/reference`` must exist):

    python oracle/make_golden.py

The reference tests ``torch.cuda.is_available`` without calling it
(models/IPSRFunction.py:28,38, util/NonparametricShift.py:15, models/InnerCos.py:19), so on a
CPU-only box it needs the two-line shim below; nothing else is patched.
``MaxCoord.update_output`` is wrapped (not altered) to record the arg-max indices / values
the reference computed.  The fixtures are the parity anchor for ``oracle/ipsr_oracle.py``
and for the CUDA path; ``/root/reference`` does not exist on the GPU box.
"""
import collections
import os
import sys

import numpy as np
import torch

REF = os.environ.get("IPSR_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _import_reference():
    torch.cuda.FloatTensor = torch.FloatTensor
    torch.Tensor.cuda = lambda self, *a, **k: self
    sys.path.insert(0, REF)
    from models.IPSR_model import IPSR_model          # noqa
    from models.InnerCos import InnerCos              # noqa
    from models.InnerCos2 import InnerCos2            # noqa
    import models.IPSRFunction as F                   # noqa
    import util.util as U                             # noqa
    return IPSR_model, InnerCos, InnerCos2, F, U


def irregular_mask(size, seed):
    """Free-form mask: a few random rectangles and strokes (bool [1,1,S,S])."""
    rng = np.random.default_rng(seed)
    m = np.zeros((size, size), bool)
    for _ in range(4):
        y, x = rng.integers(0, size - size // 4, 2)
        h, w = rng.integers(size // 16, size // 3, 2)
        m[y:y + h, x:x + w] = True
    for _ in range(3):
        y = int(rng.integers(0, size - 8))
        m[y:y + int(rng.integers(4, 12)), :] |= rng.random(size) < 0.5
    return m[None, None]


def centre_mask(size):
    m = np.zeros((1, 1, size, size), bool)
    m[:, :, size // 4:size * 3 // 4, size // 4:size * 3 // 4] = True
    return m


def make_inputs(kind, B, C, H, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, C, H, H)).astype(np.float32)
    r = rng.standard_normal((B, C, H, H)).astype(np.float32)
    g = rng.standard_normal((B, C, H, H)).astype(np.float32)
    if kind == "P1":        # well conditioned: both non-negative
        x, r = np.abs(x), np.abs(r)
    elif kind == "P2":      # signed layer input, non-negative (VGG-like) reference features
        r = np.maximum(r, 0) * 3
    elif kind == "P3":      # fully signed: chaotic regime
        pass
    return x, r, g


def run_case(name, kind, B, C, H, mask_kind, triple_w=1.0, seed=0):
    IPSR_model, _, _, F, U = _import_reference()
    S = H * 8
    if mask_kind == "centre":
        mg = centre_mask(S)
    elif mask_kind == "empty":
        mg = np.zeros((1, 1, S, S), bool)
    elif mask_kind == "full":
        mg = np.ones((1, 1, S, S), bool)
    else:
        mg = irregular_mask(S, seed + 77)
    x, r, g = make_inputs(kind, B, C, H, seed)

    rec = {"ind": [], "vmax": []}
    orig = F.MaxCoord.update_output

    def spy(self, inp, sp_x, sp_y):
        o = orig(self, inp, sp_x, sp_y)
        rec["ind"].append(o[1].clone().numpy())
        rec["vmax"].append(o[2].clone().numpy())
        return o

    F.MaxCoord.update_output = spy
    try:
        m = IPSR_model(5 / 16.0, 1, 1, 1, 1, triple_w)
        fm = m.set_mask(torch.from_numpy(mg), 3, 5 / 16.0)
        m.set_ref(collections.namedtuple("R", ["relu4_3"])(torch.from_numpy(r)))
        xt = torch.from_numpy(x).clone().requires_grad_(True)
        y = m(xt)
        y.backward(torch.from_numpy(g))
        attn_trunc = y.grad_fn.ind_lst.numpy()            # [B, N(p), H, W] int64
    finally:
        F.MaxCoord.update_output = orig
    N = H * H
    at = attn_trunc.reshape(B, N, N).transpose(0, 2, 1)   # -> [B, q, p]
    nz = np.argwhere(at != 0)
    vals = at[at != 0]
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        kind=kind, triple_w=np.float32(triple_w), mask_global=np.packbits(mg), mask_size=S,
        x=x, ref=r, g=g,
        feat_mask=fm.numpy(), flag=m.flag.numpy(), mask_point_idx=m.mask_point_idx.numpy(),
        nonmask_point_idx=m.nonmask_point_idx.numpy(), flatten_offsets=m.flatten_offsets.numpy(),
        out=y.detach().numpy(), gin=xt.grad.numpy(),
        ind=np.stack(rec["ind"]), vmax=np.stack(rec["vmax"]),
        attn_trunc_nz=nz.astype(np.int32), attn_trunc_val=vals,
    )
    print(f"{name}: M={int(m.flag.sum())} out|max|={np.abs(y.detach().numpy()).max():.3g} nnz(A_trunc)={len(vals)}")


def channel_probe(C, seed=4242):
    """Fixed fp64 weights over the channels: sum_c w[c] t[b,c,q] is a per-position fingerprint of ALL channels."""
    return np.random.default_rng(seed).standard_normal(C)


def run_case_compact(name, kind, B, C, H, mask_kind, triple_w=1.0, seed=0):
    """Full-size cases (BASELINE.json configs[2]: 64 x 64 x 256; the model's own 32 x 32 x 512): the tensors are too large
    to commit, so the fixture holds the seed (the test regenerates x / ref / g with make_inputs), every arg-max index and
    row maximum, per-position channel fingerprints of out and gin in fp64, every 16th channel of out and gin, and the
    reference's truncated attention (sparse)."""
    IPSR_model, _, _, F, U = _import_reference()
    S = H * 8
    mg = centre_mask(S) if mask_kind == "centre" else irregular_mask(S, seed + 77)
    x, r, g = make_inputs(kind, B, C, H, seed)
    rec = {"ind": [], "vmax": []}
    orig = F.MaxCoord.update_output

    def spy(self, inp, sp_x, sp_y):
        o = orig(self, inp, sp_x, sp_y)
        rec["ind"].append(o[1].clone().numpy())
        rec["vmax"].append(o[2].clone().numpy())
        return o

    F.MaxCoord.update_output = spy
    try:
        m = IPSR_model(5 / 16.0, 1, 1, 1, 1, triple_w)
        fm = m.set_mask(torch.from_numpy(mg), 3, 5 / 16.0)
        m.set_ref(collections.namedtuple("R", ["relu4_3"])(torch.from_numpy(r)))
        xt = torch.from_numpy(x).clone().requires_grad_(True)
        y = m(xt)
        y.backward(torch.from_numpy(g))
        attn_trunc = y.grad_fn.ind_lst.numpy()            # [B, N(p), H, W] int64
    finally:
        F.MaxCoord.update_output = orig
    N = H * H
    at = attn_trunc.reshape(B, N, N).transpose(0, 2, 1)   # -> [B, q, p]
    nz = np.argwhere(at != 0)
    vals = at[at != 0]
    out, gin = y.detach().numpy(), xt.grad.numpy()
    w = channel_probe(C)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        compact=True, kind=kind, B=B, C=C, H=H, seed=seed, mask_kind=mask_kind, triple_w=np.float32(triple_w),
        mask_global=np.packbits(mg), mask_size=S, feat_mask=fm.numpy(), flag=m.flag.numpy(),
        ind=np.stack(rec["ind"]).astype(np.int32), vmax=np.stack(rec["vmax"]),
        out_probe=np.einsum("c,bcq->bq", w, out.reshape(B, C, N).astype(np.float64)),
        gin_probe=np.einsum("c,bcq->bq", w, gin.reshape(B, C, N).astype(np.float64)),
        out_abs=np.abs(out.reshape(B, C, N)).astype(np.float64).sum(1), gin_abs=np.abs(gin.reshape(B, C, N)).astype(np.float64).sum(1),
        out_sub=out[:, ::16].copy(), gin_sub=gin[:, ::16].copy(),
        attn_trunc_nz=nz.astype(np.int32), attn_trunc_val=vals,
    )
    print(f"{name}: M={int(m.flag.sum())} out|max|={np.abs(out).max():.3g} nnz(A_trunc)={len(vals)}")


COMPACT_CASES = [
    # name, kind, B, C, H, mask, triple_w, seed
    ("p1_c256_h64_centre_b1_compact", "P1", 1, 256, 64, "centre", 1.0, 40),
    ("p2_c256_h64_centre_b1_compact", "P2", 1, 256, 64, "centre", 1.0, 41),
    ("p1_c512_h32_centre_b1_compact", "P1", 1, 512, 32, "centre", 1.0, 42),
    ("p2_c512_h32_irr_b2_compact", "P2", 2, 512, 32, "irr", 1.0, 43),
]


class _AnyStore(dict):
    """Stands in for the reference's ``ind_lst`` LongTensor when shift_sz != 1."""

    def cuda(self, *a, **k):
        return self


def run_patch_case(name, B, C, H, k, s, thr, mask_kind, seed):
    """Forward for shift_sz = k / stride = s.  The unmodified reference computes the whole output and then fails
    when it stores the attention for backward (``ind_lst[idx] = kbar.squeeze()``, IPSRFunction.py:134: the
    LongTensor was sized for k = 1, :36).  Only that container is replaced here (a dict that accepts the store), so
    lines :46-133 run as written and the forward output is the reference's own.  There is no backward to record:
    the reference's backward indexes the attention with the k = 1 geometry (:158-163)."""
    IPSR_model, _, _, F, U = _import_reference()
    S = H * 8
    mg = centre_mask(S) if mask_kind == "centre" else irregular_mask(S, seed + 77)
    x, r, _ = make_inputs("P1", B, C, H, seed)
    rec = {"ind": [], "vmax": []}
    orig = F.MaxCoord.update_output

    def spy(self, inp, sp_x, sp_y):
        o = orig(self, inp, sp_x, sp_y)
        rec["ind"].append(o[1].clone().numpy())
        rec["vmax"].append(o[2].clone().numpy())
        return o

    real_long = torch.LongTensor

    def long_or_store(*shape):
        return _AnyStore() if len(shape) == 4 else real_long(*shape)

    F.MaxCoord.update_output = spy
    torch.LongTensor = long_or_store
    try:
        m = IPSR_model(5 / 16.0, 1, k, s, thr, 1.0)
        fm = m.set_mask(torch.from_numpy(mg), 3, 5 / 16.0)
        m.set_ref(collections.namedtuple("R", ["relu4_3"])(torch.from_numpy(r)))
        y = m(torch.from_numpy(x))
    finally:
        torch.LongTensor = real_long
        F.MaxCoord.update_output = orig
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        kind="P1", patch=k, stride=s, mask_thred=thr, mask_global=np.packbits(mg), mask_size=S,
        x=x, ref=r, feat_mask=fm.numpy(), flag=m.flag.numpy(), mask_point_idx=m.mask_point_idx.numpy(),
        out=y.detach().numpy(), ind=np.stack(rec["ind"]), vmax=np.stack(rec["vmax"]),
    )
    print(f"{name}: k={k} s={s} M={int(m.flag.sum())} P={len(m.flag)} out|max|={np.abs(y.detach().numpy()).max():.3g}")


PATCH_CASES = [
    # name, B, C, H, k, s, mask_thred, mask, seed
    ("k3_c32_h8_irr_b1", 1, 32, 8, 3, 1, 1, "irr", 21),
    ("k3_c16_h16_irr_b2_t5", 2, 16, 16, 3, 1, 5, "irr", 22),
    ("k2s2_c16_h16_irr_b1", 1, 16, 16, 2, 2, 1, "irr", 23),
    ("k4s2_c16_h12_irr_b1_t3", 1, 16, 12, 4, 2, 3, "irr", 24),
    ("k3_c64_h16_centre_b1", 1, 64, 16, 3, 1, 1, "centre", 25),
]


def run_innercos(name, B, H, seed):
    _, InnerCos, InnerCos2, _, U = _import_reference()
    rng = np.random.default_rng(seed)
    S = H * 8
    mg = irregular_mask(S, seed)

    class Opt:
        threshold = 5 / 16.0

    x1 = rng.standard_normal((B, 512, H, H)).astype(np.float32)
    x2 = rng.standard_normal((B, 1024, H, H)).astype(np.float32)
    t = rng.standard_normal((B, 512, H, H)).astype(np.float32)
    res = {}
    for crit in ("MSE", "L1"):
        for strength in (1.0, 0.7):
            a = InnerCos(crit=crit, strength=strength, skip=0)
            a.set_mask(torch.from_numpy(mg), Opt)
            a.set_target(torch.from_numpy(t))
            a(torch.from_numpy(x1))
            b = InnerCos2(crit=crit, strength=strength, skip=0)
            b.set_mask(torch.from_numpy(mg), Opt)
            b.set_target(torch.from_numpy(t))
            b(torch.from_numpy(x2))
            res[f"loss1_{crit}_{strength}"] = np.float64(a.loss.item())
            res[f"loss2_{crit}_{strength}"] = np.float64(b.loss.item())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), mask_global=np.packbits(mg), mask_size=S,
                        x1=x1.astype(np.float16), x2=x2.astype(np.float16), t=t.astype(np.float16), **res)
    print(name, res)


def run_masks(name):
    _, _, _, _, U = _import_reference()
    out = {}
    for i, (S, seed) in enumerate([(256, 1), (256, 2), (512, 3), (128, 4)]):
        for kind in ("irr", "centre"):
            mg = irregular_mask(S, seed) if kind == "irr" else centre_mask(S)
            fm = U.cal_feat_mask(torch.from_numpy(mg), 3, 5 / 16.0).numpy()
            H = S // 8
            for (k, s, thr) in [(1, 1, 1), (3, 1, 1), (3, 1, 5), (2, 2, 1)]:
                fl, nm, fo, mi = U.cal_mask_given_mask_thred(torch.zeros(4, H, H), torch.from_numpy(fm[0, 0]), k, s, thr)
                key = f"{i}_{kind}_k{k}s{s}t{thr}"
                out[key + "_flag"] = fl.numpy()
                out[key + "_nonmask"] = nm.numpy()
                out[key + "_offsets"] = fo.numpy()
                out[key + "_maskidx"] = mi.numpy()
            out[f"{i}_{kind}_mask"] = np.packbits(mg)
            out[f"{i}_{kind}_size"] = S
            out[f"{i}_{kind}_feat"] = fm
    spx, spy = U.cal_sps_for_Advanced_Indexing(5, 7)
    out["sp_x"], out["sp_y"] = spx.numpy(), spy.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, len(out), "arrays")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)
    if "--compact-only" in sys.argv:
        for case in COMPACT_CASES:
            run_case_compact(*case)
        sys.exit(0)
    for case in PATCH_CASES:
        run_patch_case(*case)
    if "--patches-only" in sys.argv:
        sys.exit(0)
    run_masks("masks")
    run_innercos("innercos_b2_h8", 2, 8, 5)
    run_case("p1_c64_h16_centre_b1", "P1", 1, 64, 16, "centre", seed=10)
    run_case("p1_c64_h16_irr_b3_tw2p5", "P1", 3, 64, 16, "irr", triple_w=2.5, seed=11)
    run_case("p1_c32_h8_empty_b2", "P1", 2, 32, 8, "empty", seed=12)
    run_case("p1_c32_h8_full_b1", "P1", 1, 32, 8, "full", seed=13)
    run_case("p2_c64_h16_irr_b2", "P2", 2, 64, 16, "irr", seed=14)
    run_case("p3_c32_h8_centre_b1", "P3", 1, 32, 8, "centre", seed=15)
    run_case("p1_c512_h16_centre_b1", "P1", 1, 512, 16, "centre", seed=16)
    run_case("p1_c256_h32_centre_b1", "P1", 1, 256, 32, "centre", seed=17)
    for case in COMPACT_CASES:
        run_case_compact(*case)
