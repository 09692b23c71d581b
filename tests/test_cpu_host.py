"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol the header declares,
the ctypes mirror of the argument struct matches, host-side index logic matches the reference's
golden vectors, and the multi-GPU partitioning / exchange logic works under gloo (world_size 2)."""
import ctypes
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, SHIFT_CASES, golden_path


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ipsr_sm100.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b((?:ipsr|innercos)_[a-z0-9_]+)\s*\(", text))
    names.discard("ipsr_fwd_args")
    return names


def test_library_exports_every_declared_symbol():
    from deepinpainting_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 25
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.ipsr_version() == 200
    assert isinstance(_lib.last_error(), str)


def test_library_is_sm100a_native():
    """The shipped binary must carry sm_100a SASS with tcgen05 / TMEM / bulk-copy instructions."""
    import shutil
    import subprocess
    from deepinpainting_b200 import build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", build.build_library()], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UBLKCP"):
        assert mnemonic in sass, mnemonic


def test_shape_support_queries():
    from deepinpainting_b200 import _lib
    lib = _lib.load()
    assert lib.ipsr_tensor_path_supported(256, 1024) == 1
    assert lib.ipsr_tensor_path_supported(512, 4096) == 1
    assert lib.ipsr_tensor_path_supported(32, 64) == 0
    assert lib.ipsr_tensor_path_supported(256, 1000) == 0
    small = lib.ipsr_workspace_bytes(1, 32, 8, 8, 16, _lib.IPSR_MODE_AUTO)
    big = lib.ipsr_workspace_bytes(16, 256, 32, 32, 256, _lib.IPSR_MODE_AUTO)
    assert 0 < small < big
    assert lib.ipsr_workspace_bytes(0, 32, 8, 8, 0, 0) == 0
    # 16 images of config A must fit comfortably in HBM
    assert big < 1 << 30


def test_argument_validation_without_gpu():
    """Launchers validate before touching the device: bad arguments give negative codes + a message."""
    from deepinpainting_b200 import _lib
    lib = _lib.load()
    assert lib.ipsr_blend_scan(None, 1, 32, 4, None, None, None, None) == -1
    assert "null" in _lib.last_error()
    args = _lib.FwdArgs()
    assert lib.ipsr_shift_forward(ctypes.byref(args), None) == -1
    assert lib.ipsr_correlate_argmax_tc(1, 1, 1, 48, 100, 0, 100, 1, 1, 2, None, 1, 1, 1, None, None, None, None) == -2
    with pytest.raises(_lib.IpsrError):
        _lib.call("ipsr_maxcoord", None, 0, 0, None, None, None)


def test_struct_layout_matches_header():
    from deepinpainting_b200 import _lib
    text = open(os.path.join(ROOT, "include", "ipsr_sm100.h")).read()
    body = text[text.index("typedef struct ipsr_fwd_args"):text.index("} ipsr_fwd_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    body = body[body.index("{") + 1:]
    fields = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        names = [re.sub(r"[\s\*]", "", n).split(" ")[-1] for n in stmt.split(",")]
        names[0] = re.findall(r"([A-Za-z_0-9]+)\s*$", stmt.split(",")[0].replace("*", " "))[0]
        fields.extend(names)
    assert fields == [f[0] for f in _lib.FwdArgs._fields_]


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "deepinpainting_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", ""), os.path.join(dirpath, f)


def test_ops_refuse_cpu_tensors():
    from deepinpainting_b200 import shift_ops
    with pytest.raises(RuntimeError, match="no CPU path"):
        shift_ops.feat_mask(torch.zeros(8, 8), 3, 0.3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        shift_ops.innercos_loss(torch.zeros(1, 2, 2, 2), torch.zeros(2, 2), torch.zeros(1, 2, 2, 2), 1.0, "MSE")


def test_flatten_offsets_closed_form_matches_reference():
    from deepinpainting_b200.util.util import flatten_offsets_from_flag, cal_sps_for_Advanced_Indexing
    z = np.load(golden_path("masks"))
    n = 0
    for k in z.files:
        if k.endswith("_flag"):
            fo = flatten_offsets_from_flag(torch.from_numpy(z[k])).numpy()
            np.testing.assert_array_equal(fo, z[k.replace("_flag", "_offsets")])
            n += 1
    assert n == 32
    for name in SHIFT_CASES:
        g = np.load(golden_path(name))
        np.testing.assert_array_equal(flatten_offsets_from_flag(torch.from_numpy(g["flag"])).numpy(), g["flatten_offsets"])
    sx, sy = cal_sps_for_Advanced_Indexing(5, 7)
    np.testing.assert_array_equal(sx.numpy(), z["sp_x"])
    np.testing.assert_array_equal(sy.numpy(), z["sp_y"])


def test_module_api_surface():
    """Constructor signatures / methods the reference's networks.py and IPSR.py rely on."""
    import inspect
    from deepinpainting_b200.models import IPSR_model, IPSRFunction, InnerCos, InnerCos2
    from deepinpainting_b200.util.NonparametricShift import NonparametricShift
    from deepinpainting_b200.util.MaxCoord import MaxCoord
    from deepinpainting_b200.util import util
    m = IPSR_model(5 / 16.0, 1, 1, 1, 1, 1)
    assert list(inspect.signature(IPSR_model.__init__).parameters)[1:] == [
        "threshold", "fixed_mask", "shift_sz", "stride", "mask_thred", "triple_weight"]
    for meth in ("set_mask", "set_ref", "forward"):
        assert callable(getattr(m, meth))
    assert len(list(m.parameters())) == 0 and len(list(m.buffers())) == 0 and len(m.state_dict()) == 0
    assert repr(m) == "IPSR_model(threshold: 0.3125 ,triple_weight 1)"
    assert list(inspect.signature(IPSRFunction.forward).parameters) == [
        "ctx", "input", "mask", "ref", "shift_sz", "stride", "triple_w", "flag", "nonmask_point_idx",
        "mask_point_idx", "flatten_offsets", "sp_x", "sp_y"]
    a, b = InnerCos(strength=2, skip=0), InnerCos2(strength=1, skip=1, infe=3)
    assert repr(a) == "InnerCos(skip: True ,strength: 2)"       # inverted skip string, as the reference
    assert repr(b) == "InnerCos2(skip: False ,strength: 1)"
    x = torch.zeros(1, 4, 2, 2)
    assert b(x) is x and b.loss == 0                              # skipped layer is a pure identity
    for meth in ("set_mask", "set_target", "get_target", "forward", "backward"):
        assert callable(getattr(a, meth))
    assert list(inspect.signature(NonparametricShift.buildAutoencoder).parameters)[1:] == [
        "target_img", "normalize", "interpolate", "nonmask_point_idx", "mask_point_idx", "patch_size", "stride"]
    with pytest.raises(NotImplementedError):
        NonparametricShift()._build(1, 1, 4, torch.zeros(2, 4, 1, 1), 2, True, False)
    with pytest.raises(AssertionError, match="The first dimension"):
        MaxCoord().update_output(torch.zeros(2, 3, 2, 2), None, None)
    with pytest.raises(AssertionError, match="mask must be 4 dimensions"):
        util.cal_feat_mask(torch.zeros(4, 4), 3, 0.3)
    with pytest.raises(AssertionError, match="Mask dimension must be 2"):
        IPSRFunction.apply(torch.zeros(1, 32, 2, 2), torch.zeros(1, 2, 2), None, 1, 1, 1, None, None, None, None, None, None)


def test_shard_arithmetic():
    from deepinpainting_b200.sharding import shard_bank, shard_batch
    for B in (1, 7, 16, 64):
        for w in (1, 2, 4, 8):
            spans = [shard_batch(B, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [shard_bank(1024, 8, r) for r in range(8)] == [(128 * r, 128 * (r + 1)) for r in range(8)]
    assert shard_bank(256, 4, 3) == (256, 256)          # more ranks than tiles: empty shard
    with pytest.raises(ValueError):
        shard_bank(100, 2, 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker_exchange(rank, world, port, case):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from deepinpainting_b200.sharding import allreduce_max_keys, shard_bank, shard_batch
    from oracle import ipsr_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        z = np.load(golden_path(case))
        x, ref, flag = z["x"], z["ref"], z["flag"]
        B, C, H, W = x.shape
        N = H * W
        # ---- bank-sharded: local (max, idx) over this rank's columns -> one all-reduce MAX ----
        cb, ce = shard_bank(N, world, rank, align=N // 2)
        keys = np.empty((B, N), np.int64)
        for b in range(B):
            X = x[b].reshape(C, N).T
            R = ref[b].reshape(C, N).T
            inv = np.float32(1) / (np.sqrt((X * X).sum(1, dtype=np.float32)) + np.float32(1e-8))
            S = (R @ (X * inv[:, None]).T.astype(np.float32)).astype(np.float32)[:, cb:ce]
            li = S.argmax(1)
            keys[b] = O.pack_max_idx(S[np.arange(N), li], li + cb)
        t = torch.from_numpy(keys)
        allreduce_max_keys(t)
        v, idx = O.unpack_max_idx(t.numpy())
        full = O.shift_forward(x, ref, flag, np.float32, keep_attn=False)
        np.testing.assert_array_equal(idx, full.ind)
        np.testing.assert_allclose(v, full.vmax, rtol=2e-6)
        # ---- batch-sharded: no collective; gathering the slices reproduces the full batch ----
        b0, b1 = shard_batch(B, world, rank)
        part = O.shift_forward(x[b0:b1], ref[b0:b1], flag, np.float32, keep_attn=False).out if b1 > b0 else np.zeros((0,) + x.shape[1:], np.float32)
        outs = [None] * world
        dist.all_gather_object(outs, part)
        np.testing.assert_array_equal(np.concatenate(outs, 0), full.out)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_exchange_world2_gloo():
    """world_size-2 gloo run of the multi-GPU host logic: key packing + MAX all-reduce gives the
    global arg-max with torch.max's tie rule; batch shards concatenate to the full result."""
    import torch.multiprocessing as mp
    mp.spawn(_worker_exchange, args=(2, _free_port(), "p1_c64_h16_irr_b3_tw2p5"), nprocs=2, join=True)


def test_key_packing_order_and_ties():
    from oracle import ipsr_oracle as O
    v = np.array([0.5, 0.5, -0.0, 0.0, -1.0, np.inf, -np.inf, np.nan, 1e-38, -1e-38], np.float32)
    idx = np.arange(10)
    k = O.pack_max_idx(v, idx)
    order = np.argsort(-k, kind="stable")
    assert order[0] == 7 and order[1] == 5          # NaN wins, then +inf
    assert k[0] > k[1]                               # equal score: lower index wins
    assert k[2] > k[3]                               # -0 == +0, lower index wins
    vv, ii = O.unpack_max_idx(k)
    np.testing.assert_array_equal(ii, idx)
    np.testing.assert_array_equal(vv[[0, 1, 4, 5, 6, 8, 9]], v[[0, 1, 4, 5, 6, 8, 9]])
    assert (k > np.int64(-(1 << 63))).all()


def test_built_library_stays_current_when_the_tree_moves(tmp_path):
    """The snapshot that travels to a GPU box lives under another directory there: the staleness check must look at file
    CONTENTS only (a library that looked stale was rebuilt by every rank of a job at once)."""
    import shutil, subprocess, sys
    from deepinpainting_b200 import build as _build
    if not _build.is_current():
        pytest.skip("library not built in this tree")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dst = tmp_path / "moved"
    shutil.copytree(os.path.join(root, "deepinpainting_b200"), dst / "deepinpainting_b200",
                    ignore=shutil.ignore_patterns("obj", "__pycache__"))
    shutil.copytree(os.path.join(root, "include"), dst / "include")
    out = subprocess.run([sys.executable, "-c", "import deepinpainting_b200.build as b; print(b.PKG); print(b.is_current())"],
                         cwd=str(dst), capture_output=True, text=True, check=True).stdout.split()
    assert out[0].startswith(str(dst)) and out[1] == "True", out
