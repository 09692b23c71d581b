"""CPU oracle for the IPSR / CSA patch-shift attention layer.

TEST INFRASTRUCTURE ONLY.  This file is a numpy restatement of the reference's
algorithm for the hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the
product (``deepinpainting_b200``) never does and fails loudly without its CUDA
library.

Parity status: the reference ships no tests, golden vectors or fixtures for this
path, so this restatement is pinned against OUTPUTS OF THE
REFERENCE ITSELF, executed on CPU in the build container by
``oracle/make_golden.py`` and committed as ``tests/golden/*.npz``
(``tests/test_oracle_golden.py`` checks every fixture).

The arithmetic of the reference lives in PyTorch (pinned torch==1.5.1, req.txt:60);
the call sites restated here are cited per function as ``file:line`` relative to
the reference tree.

All functions take / return numpy arrays.  ``dtype`` selects float32 (canonical:
the reference computes in fp32) or float64 (used to measure conditioning and the
true top-2 score gap).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

INT64_MIN = np.iinfo(np.int64).min


# --------------------------------------------------------------------------------------
# mask helpers
# --------------------------------------------------------------------------------------
def cal_feat_mask(in_mask: np.ndarray, conv_layers: int = 3, threshold: float = 5 / 16.0) -> np.ndarray:
    """util/util.py:68-84.  ``conv_layers`` chained Conv2d(1,1,4,2,1,bias=False) with every
    weight = 1/16, then ``> threshold`` -> uint8.  Input [1,1,S,S] (or [S,S]) of bool/0-1.
    Returns uint8 [1,1,S/2^L,S/2^L].  Arithmetic is exact in fp32 (dyadic weights)."""
    m = np.asarray(in_mask)
    if m.ndim == 2:
        m = m[None, None]
    assert m.ndim == 4, "mask must be 4 dimensions"
    assert m.shape[0] == 1, "the first dimension must be 1 for mask"
    cur = m[0, 0].astype(np.float32)
    for _ in range(conv_layers):
        h, w = cur.shape
        pad = np.zeros((h + 2, w + 2), np.float32)
        pad[1:-1, 1:-1] = cur
        oh, ow = (h + 2 - 4) // 2 + 1, (w + 2 - 4) // 2 + 1
        acc = np.zeros((oh, ow), np.float32)
        for dy in range(4):
            for dx in range(4):
                acc += pad[dy:dy + 2 * oh:2, dx:dx + 2 * ow:2] * np.float32(1 / 16)
        cur = acc
    return (cur > np.float32(threshold)).astype(np.uint8)[None, None]


def cal_mask_given_mask_thred(img_shape, mask: np.ndarray, patch_size: int, stride: int, mask_thred: int):
    """util/util.py:88-161.  Returns (flag[N], nonmask_point_idx, flatten_offsets, mask_point_idx),
    all int64.  ``nonmask_point_idx`` is every index 0..N-1 (the known-only branch is commented
    out at :121-129 and replaced by :137-139).  ``flatten_offsets`` follows :150-157 literally
    (it is unused by the operator, IPSRFunction.py:88-89)."""
    assert len(img_shape) == 3, "img has to be 3 dimenison!"
    mask = np.asarray(mask)
    assert mask.ndim == 2, "mask has to be 2 dimenison!"
    _, H, W = img_shape
    nH = int(math.floor((H - patch_size) / stride + 1))
    nW = int(math.floor((W - patch_size) / stride + 1))
    N = nH * nW
    flag = np.zeros(N, np.int64)
    offsets_tmp = np.zeros(N, np.int64)
    mask_idx = []
    m64 = mask.astype(np.int64)
    for i in range(N):
        h, w = i // nW, i % nW
        win = m64[h * stride:h * stride + patch_size, w * stride:w * stride + patch_size]
        if win.sum() >= mask_thred:
            mask_idx.append(i)
            flag[i] = 1
            offsets_tmp[i] = -1
    nonmask_point_idx = np.arange(N, dtype=np.int64)
    mask_point_idx = np.asarray(mask_idx, dtype=np.int64)
    # :150-157 (python negative indices wrap exactly as torch indexing does)
    flatten_all = np.zeros(N, np.int64)
    csum = np.cumsum(offsets_tmp)
    for i in range(N):
        ov = int(csum[i])
        if flag[i] == 1:
            ov += 1
        flatten_all[i + ov] = -ov
    return flag, nonmask_point_idx, flatten_all[:N].copy(), mask_point_idx


def cal_sps_for_advanced_indexing(h: int, w: int):
    """util/util.py:166-174."""
    sp_y = np.tile(np.arange(w, dtype=np.int64), h)
    sp_x = np.repeat(np.arange(h, dtype=np.int64), w)
    return sp_x, sp_y


# --------------------------------------------------------------------------------------
# patch bank (k = shift_sz, stride)
# --------------------------------------------------------------------------------------
def extract_patches(img: np.ndarray, patch_size: int = 1, stride: int = 1) -> np.ndarray:
    """util/NonparametricShift.py:59-68: unfold(1,k,s).unfold(2,k,s).permute(1,2,0,3,4) ->
    [nH*nW, C, k, k] in raster order."""
    C, H, W = img.shape
    nH = (H - patch_size) // stride + 1
    nW = (W - patch_size) // stride + 1
    out = np.empty((nH * nW, C, patch_size, patch_size), img.dtype)
    for i in range(nH):
        for j in range(nW):
            out[i * nW + j] = img[:, i * stride:i * stride + patch_size, j * stride:j * stride + patch_size]
    return out


def l2_normalize_patches(patches: np.ndarray) -> np.ndarray:
    """util/NonparametricShift.py:36-40: ``p * (1 / (p.norm(2) + 1e-8))`` per patch
    (reciprocal, then multiply)."""
    dt = patches.dtype.type
    flat = patches.reshape(patches.shape[0], -1)
    nrm = np.sqrt((flat * flat).sum(axis=1, dtype=patches.dtype))
    inv = dt(1) / (nrm + dt(1e-8))
    return (flat * inv[:, None]).reshape(patches.shape).astype(patches.dtype)


# --------------------------------------------------------------------------------------
# shift operator, k = 1 / stride = 1 (the only configuration the reference can execute)
# --------------------------------------------------------------------------------------
@dataclass
class ShiftResult:
    out: np.ndarray                    # [B,C,H,W]
    ind: np.ndarray                    # [B,N] int64   arg-max bank index per query position
    vmax: np.ndarray                   # [B,N]         max score per query position
    wn: np.ndarray                     # [B,M]         blend weight on the previous masked output (l>=1)
    wo: np.ndarray                     # [B,M]         blend weight on the matched patch (l>=1)
    attn: Optional[np.ndarray] = None  # [B,N(q),N(p)] float attention A (IPSRFunction.py kbar)
    attn_trunc: Optional[np.ndarray] = None  # [B,N,N] int64, what the reference saves (ind_lst)
    gap: Optional[np.ndarray] = None   # [B,N] top-2 score gap
    extras: dict = field(default_factory=dict)


def float_to_int64_trunc(a: np.ndarray) -> np.ndarray:
    """``LongTensor[...] = FloatTensor`` (IPSRFunction.py:36,134): C cast on x86-64 ->
    truncation toward zero; NaN, +-inf and out-of-range give INT64_MIN (cvttss2si
    'integer indefinite')."""
    a = np.asarray(a)
    bad = ~np.isfinite(a) | (np.abs(a) >= 2.0 ** 63)
    safe = np.where(bad, 0, a)
    out = np.trunc(safe).astype(np.int64)
    out[bad] = INT64_MIN
    return out


def shift_forward(x: np.ndarray, ref: np.ndarray, flag: np.ndarray, dtype=np.float32,
                  keep_attn: bool = True, with_gap: bool = True) -> ShiftResult:
    """IPSRFunction.py:13-140 with shift_sz = stride = 1.

    x   [B,C,H,W]  the layer input (bank source; patches of ALL positions form the bank)
    ref [B,C,H,W]  ``ref.relu4_3`` (query features)
    flag[N]        1 = masked position (util.cal_mask_given_mask_thred)
    """
    x = np.asarray(x, dtype)
    ref = np.asarray(ref, dtype)
    B, C, H, W = x.shape
    N = H * W
    flag = np.asarray(flag).reshape(-1)
    assert flag.shape[0] == N
    midx = np.nonzero(flag)[0]
    M = midx.shape[0]
    dt = np.dtype(dtype).type

    out = np.empty_like(x)
    ind_all = np.empty((B, N), np.int64)
    vmax_all = np.empty((B, N), dtype)
    gap_all = np.empty((B, N), dtype)
    wn_all = np.zeros((B, M), dtype)
    wo_all = np.zeros((B, M), dtype)
    attn_all = np.zeros((B, N, N), dtype) if keep_attn else None

    for b in range(B):
        X = x[b].reshape(C, N).T.copy()                       # [N,C] raster order      :54 / NPS:65-68
        R = ref[b].reshape(C, N).T.copy()                     # [N,C]                   :49
        nrm = np.sqrt((X * X).sum(axis=1, dtype=dtype))
        inv = dt(1) / (nrm + dt(1e-8))                        # NPS:40
        Xn = (X * inv[:, None]).astype(dtype)
        S = (R @ Xn.T).astype(dtype)                          # :59  conv_enc(ref): S[q,p]
        ind = S.argmax(axis=1)                                # MaxCoord.py:22 (first index on ties)
        vmax = S[np.arange(N), ind]
        if not with_gap:
            gap_all[b] = np.nan
        elif N > 1:
            part = np.partition(S, N - 2, axis=1)
            gap_all[b] = part[:, N - 1] - part[:, N - 2]
        else:
            gap_all[b] = np.inf
        A = np.zeros((N, N), dtype)                           # kbar [1,N(p),H,W] viewed [q,p]
        A[np.arange(N), ind] = 1                              # :129 (unmasked rows; masked rows overwritten)
        prev_out = None
        prev_row = None
        for l, q in enumerate(midx):                          # :82-126 raster order over masked positions
            p = ind[q]
            if l == 0:                                        # :98-101
                prev_out = X[p].copy()
                row = np.zeros(N, dtype)
                row[p] = 1
            else:                                             # :104-125
                little = X[q]
                nq = np.sqrt((little * little).sum(dtype=dtype))
                u = (little * (dt(1) / (nq + dt(1e-8)))).astype(dtype)        # :109
                a = dt((u * prev_out).sum(dtype=dtype))                        # :116 1x1 conv == dot
                v = vmax[q]                                                    # :70
                with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
                    wn = dt(a) / (dt(a) + dt(v))                               # :120
                    wo = dt(v) / (dt(a) + dt(v))                               # :121
                    prev_out = (wn * prev_out + wo * X[p]).astype(dtype)       # :122
                    row = (prev_row * wn).astype(dtype)                        # :123
                    row[p] = row[p] + wo                                       # :124
                wn_all[b, l] = wn
                wo_all[b, l] = wo
            A[q] = row
            prev_row = row
        with np.errstate(invalid="ignore", over="ignore"):
            out[b] = (A @ X).T.reshape(C, H, W)               # :131 conv_transpose with raw patches
        ind_all[b] = ind
        vmax_all[b] = vmax
        if keep_attn:
            attn_all[b] = A

    res = ShiftResult(out=out, ind=ind_all, vmax=vmax_all, wn=wn_all, wo=wo_all, attn=attn_all, gap=gap_all)
    if keep_attn:
        res.attn_trunc = float_to_int64_trunc(attn_all)       # :134 ind_lst[idx] = kbar.squeeze()
    return res


def shift_backward(grad_out: np.ndarray, attn_trunc: np.ndarray, triple_w: float, dtype=np.float32) -> np.ndarray:
    """IPSRFunction.py:144-178.  ``gin = g + triple_w * (W^T g)`` with W = float(int64 attention).
    Uses row index i*h+j (:162), i.e. assumes square maps like the reference."""
    g = np.asarray(grad_out, dtype)
    B, C, H, W = g.shape
    N = H * W
    gin = np.empty_like(g)
    for b in range(B):
        Wm = attn_trunc[b].astype(dtype)                      # :158-163  W_mat[q, p]
        G = g[b].reshape(C, N).T                              # :169  [N,C]
        with np.errstate(invalid="ignore", over="ignore"):
            weighted = (Wm.T @ G).astype(dtype)               # :169
            gin[b] = g[b] + (weighted.T.reshape(C, H, W) * np.dtype(dtype).type(triple_w))   # :172-173
    return gin


# --------------------------------------------------------------------------------------
# general patch size, FORWARD ONLY.  For k != 1 the reference computes the whole output (lines :46-133) and then
# fails storing the attention for backward (:134, a LongTensor sized for k = 1); its backward (:158-163) indexes
# with the k = 1 geometry.  The forward is pinned against the reference's own output (oracle/make_golden.py
# run_patch_case replaces only that container; tests/golden/k*.npz); the backward is undefined and not restated.
# --------------------------------------------------------------------------------------
def shift_forward_patches(x: np.ndarray, ref: np.ndarray, mask2d: np.ndarray, patch_size: int, stride: int,
                          mask_thred: int = 1, dtype=np.float32):
    """Forward for shift_sz=k, stride=s following IPSRFunction.py:54-131 literally:
    bank = all k x k patches of x (normalised, NPS:36-40) used as Conv2d(C,P,k,s) filters on
    ref (:59) -> S[p, i, j]; arg-max over p per output location (:65); blend over masked
    locations (:82-126) on [C,k,k] patches with the dot taken over the whole patch (:116);
    paste with ConvTranspose2d(P,C,k,s) whose weights are the raw patches (:131), which SUMS
    overlapping contributions.  Returns (out [B,C,H',W'], ind [B,P])."""
    x = np.asarray(x, dtype)
    ref = np.asarray(ref, dtype)
    B, C, H, W = x.shape
    k, s = patch_size, stride
    nH, nW = (H - k) // s + 1, (W - k) // s + 1
    P = nH * nW
    flag, _, _, midx = cal_mask_given_mask_thred((C, H, W), mask2d, k, s, mask_thred)
    dt = np.dtype(dtype).type
    Ho, Wo = (nH - 1) * s + k, (nW - 1) * s + k
    outs = np.zeros((B, C, Ho, Wo), dtype)
    inds = np.zeros((B, P), np.int64)
    for b in range(B):
        pat = extract_patches(x[b], k, s)                     # [P,C,k,k]
        patn = l2_normalize_patches(pat)
        rpat = extract_patches(ref[b], k, s)                  # conv windows of ref at the same grid
        S = (rpat.reshape(P, -1) @ patn.reshape(P, -1).T).astype(dtype)   # S[q,p]
        ind = S.argmax(axis=1)
        vmax = S[np.arange(P), ind]
        A = np.zeros((P, P), dtype)
        A[np.arange(P), ind] = 1
        prev_out = prev_row = None
        for l, q in enumerate(midx):
            p = ind[q]
            if l == 0:
                prev_out = pat[p].copy()
                row = np.zeros(P, dtype)
                row[p] = 1
            else:
                little = pat[q]
                nq = np.sqrt((little * little).sum(dtype=dtype))
                u = (little * (dt(1) / (nq + dt(1e-8)))).astype(dtype)
                a = dt((u * prev_out).sum(dtype=dtype))
                v = vmax[q]
                with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
                    wn = a / (a + v)
                    wo = v / (a + v)
                    prev_out = (wn * prev_out + wo * pat[p]).astype(dtype)
                    row = (prev_row * wn).astype(dtype)
                    row[p] = row[p] + wo
            A[q] = row
            prev_row = row
        # conv_transpose2d: out[c, i*s+dy, j*s+dx] += sum_p A[q=(i,j), p] * pat[p, c, dy, dx]
        with np.errstate(invalid="ignore", over="ignore"):
            mixed = (A @ pat.reshape(P, -1)).reshape(P, C, k, k).astype(dtype)
        for i in range(nH):
            for j in range(nW):
                outs[b, :, i * s:i * s + k, j * s:j * s + k] += mixed[i * nW + j]
        inds[b] = ind
    return outs, inds


# --------------------------------------------------------------------------------------
# InnerCos / InnerCos2 side loss
# --------------------------------------------------------------------------------------
def innercos_loss(x: np.ndarray, mask2d: np.ndarray, target: np.ndarray, strength: float = 1.0,
                  crit: str = "MSE", c_limit: Optional[int] = None, dtype=np.float32) -> float:
    """models/InnerCos.py:30-36 (``c_limit=None``) and models/InnerCos2.py:34-41
    (``c_limit=512``: ``in_data.narrow(1,0,512)``).  loss = mean over all elements of
    (x*mask*strength - target)^2, or of |.| when crit != 'MSE'."""
    x = np.asarray(x, dtype)
    if c_limit is not None:
        x = x[:, :c_limit]
    m = np.asarray(mask2d, dtype)
    d = x * m[None, None] * np.dtype(dtype).type(strength) - np.asarray(target, dtype)
    if crit == "MSE":
        return float((d.astype(np.float64) ** 2).mean())
    return float(np.abs(d.astype(np.float64)).mean())


# --------------------------------------------------------------------------------------
# bank-sharded (max, idx) reduction -- the one exchange step of the sharded mode
# --------------------------------------------------------------------------------------
def pack_max_idx(v: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """Order-preserving packing used for the (max, index) all-reduce, as SIGNED int64 so that
    ncclMax / gloo MAX on int64 tensors apply: high 32 bits = fp32 score mapped to an orderable
    signed int (-0 canonicalised to +0, NaN canonicalised to the quiet NaN above +inf because
    torch.max lets NaN win), low 32 bits = 0xFFFFFFFF - idx so that on equal scores the LOWEST
    index wins (MaxCoord.py:22 tie rule).  Identity of the reduction = INT64_MIN."""
    v = np.asarray(v, np.float32) + np.float32(0.0)
    bits = v.view(np.uint32).copy()
    bits[np.isnan(v)] = np.uint32(0x7FC00000)
    neg = (bits >> np.uint32(31)) == 1
    key = np.where(neg, bits ^ np.uint32(0x7FFFFFFF), bits).astype(np.uint32)
    low = (np.uint64(0xFFFFFFFF) - np.asarray(idx).astype(np.uint64))
    return ((key.astype(np.uint64) << np.uint64(32)) | low).view(np.int64)


def unpack_max_idx(packed: np.ndarray):
    u = np.ascontiguousarray(np.asarray(packed)).view(np.uint64)
    key = (u >> np.uint64(32)).astype(np.uint32)
    idx = (np.uint64(0xFFFFFFFF) - (u & np.uint64(0xFFFFFFFF))).astype(np.int64)
    neg = (key >> np.uint32(31)) == 1
    bits = np.where(neg, key ^ np.uint32(0x7FFFFFFF), key).astype(np.uint32)
    return bits.view(np.float32), idx


def centre_mask(size: int) -> np.ndarray:
    """bool [1,1,S,S] with the hole = the middle half of the image."""
    m = np.zeros((1, 1, size, size), bool)
    m[:, :, size // 4:size * 3 // 4, size // 4:size * 3 // 4] = True
    return m
