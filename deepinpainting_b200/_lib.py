"""ctypes binding of libipsr_sm100.so (see include/ipsr_sm100.h).

There is NO CPU or PyTorch fallback: if the shared library is missing it is built with nvcc
(``deepinpainting_b200.build``); if that is impossible ``load()`` raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build

IPSR_MODE_AUTO, IPSR_MODE_TENSOR, IPSR_MODE_EXACT = 0, 1, 2
MODES = {"auto": IPSR_MODE_AUTO, "tensor": IPSR_MODE_TENSOR, "exact": IPSR_MODE_EXACT}

_p = C.c_void_p
_i = C.c_int
_f = C.c_float
_i64 = C.c_int64


class FwdArgs(C.Structure):
    """Mirror of ``ipsr_fwd_args`` (include/ipsr_sm100.h)."""
    _fields_ = [
        ("x", _p), ("ref", _p), ("flag", _p), ("mask_idx", _p), ("rank", _p),
        ("B", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("M", C.c_int32),
        ("mode", C.c_int32), ("need_grad", C.c_int32),
        ("col_begin", C.c_int32), ("col_end", C.c_int32), ("stop_after_corr", C.c_int32),
        ("psplit", C.c_int32), ("exc_cap", C.c_int32),
        ("tol_rel", _f), ("tol_abs", _f),
        ("out", _p), ("ind", _p), ("wn", _p), ("wo", _p),
        ("route_ptr", _p), ("route_q", _p),
        ("exc_start", _p), ("exc_cnt", _p), ("exc_l", _p), ("exc_w", _p), ("exc_total", _p),
        ("nrecheck_out", _p), ("npass2_out", _p), ("ev_corr_begin", _p), ("ev_corr_end", _p),
        ("workspace", _p), ("workspace_bytes", C.c_size_t),
        ("mask_stride", C.c_int32), ("m_count", _p),
        ("cos_target", _p), ("cos_mask", _p), ("cos_strength", _f), ("cos_crit", C.c_int32),
        ("cos_partials", _p), ("cos_ticket", _p), ("cos_loss", _p),
    ]


# name -> (restype, argtypes); every symbol declared in include/ipsr_sm100.h
SIGNATURES = {
    "ipsr_last_error_string": (C.c_char_p, []),
    "ipsr_version": (_i, []),
    "ipsr_abi_fwd_args_bytes": (_i, []),
    "ipsr_tensor_path_supported": (_i, [_i, _i]),
    "ipsr_tensor_cascade": (_i, [_i, _i, _i]),
    "ipsr_tensor_full_passes": (_i, [_i, _i, _i]),
    "ipsr_feat_mask": (_i, [_p, _i, _i, _i, _f, _p, _p, _p]),
    "ipsr_build_flags": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "ipsr_feat_mask_batch": (_i, [_p, _i, _i, _i, _i, _f, _p, _p, _p]),
    "ipsr_build_flags_batch": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "ipsr_extract_normalize": (_i, [_p, _p, _i, _i, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "ipsr_compact_rows": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _p]),
    "ipsr_correlate_argmax_tc": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "ipsr_correlate_argmax_tc_valid": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _i, _p]),
    "ipsr_finalize_argmax_valid": (_i, [_p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _i, _i, _f, _f, _p, _p, _p, _p, _p, _p, _i,
                                        _p, _p, _p, _p, _p, _i, _p]),
    "ipsr_patch_rows_stats": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "ipsr_patch_tiles": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "ipsr_patch_recheck": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "ipsr_patch_resolve_pairs": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "ipsr_patch_winner_scores": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p]),
    "ipsr_finalize_argmax": (_i, [_p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _i, _i, _f, _f, _p, _p, _p, _p, _p, _p, _i,
                                  _p, _p, _p, _p, _p, _p]),
    "ipsr_resolve_rows": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p]),
    "ipsr_select_all_rows": (_i, [_i, _i, _p, _p, _p, _p]),
    "ipsr_correlate_argmax_fp32": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _i, _p, _p]),
    "ipsr_apply_recheck": (_i, [_p, _p, _p, _i, _i, _p, _p, _p]),
    "ipsr_pack_winner_scores": (_i, [_p, _p, _p, _p, _i, _i, _i, _p, _p]),
    "ipsr_pack_maxidx": (_i, [_p, _p, _i64, _p, _p]),
    "ipsr_unpack_maxidx": (_i, [_p, _i64, _p, _p, _p]),
    "ipsr_maxcoord": (_i, [_p, _i, _i, _p, _p, _p]),
    "ipsr_blend_stage": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "ipsr_blend_stage_with_routes": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "ipsr_scan_block_steps": (_i, [_i]),
    "ipsr_staged_block_floats": (_i, [_i]),
    "ipsr_padded_steps": (_i, [_i]),
    "ipsr_blend_scan": (_i, [_p, _i, _i, _i, _p, _p, _p, _p]),
    "ipsr_paste_with_bookkeeping": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p]),
    "ipsr_paste": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _p]),
    "ipsr_paste_loss_partials": (_i, [_i, _i, _i]),
    "ipsr_build_routes": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _p]),
    "ipsr_build_exceptions": (_i, [_p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _i, _p]),
    "ipsr_shift_bwd": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _f, _p, _p]),
    "ipsr_patch_row_len": (_i, [_i, _i]),
    "ipsr_unfold_patches": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "ipsr_fold_patches": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "ipsr_patch_rows": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "ipsr_blend_wide": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "ipsr_blend_wide_gram_floats": (_i, [_i, _i]),
    "ipsr_blend_wide_blocked": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "ipsr_fold_patch_rows": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "ipsr_shift_bwd_masks": (_i, [_p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _f, _p, _i, _p, _p]),
    "innercos_loss_fwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _f, _i, _p, _p, _p, _p]),
    "innercos_loss_bwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _f, _i, _p, _p]),
    "ipsr_workspace_bytes": (C.c_size_t, [_i, _i, _i, _i, _i, _i]),
    "ipsr_shift_forward": (_i, [C.POINTER(FwdArgs), _p]),
    "ipsr_shift_forward_finish": (_i, [C.POINTER(FwdArgs), _p]),
    "ipsr_workspace_packed": (_p, [C.POINTER(FwdArgs)]),
}

_lock = threading.Lock()
_lib = None


class IpsrError(RuntimeError):
    pass


def load(build_if_missing: bool = True):
    """Return the loaded library handle; raises if it cannot be produced."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.library_path()
        if not os.path.exists(path):
            if not build_if_missing:
                raise IpsrError("libipsr_sm100.so not found at %s and building was disabled" % path)
            path = _build.build_library()
        elif not _build.is_current():
            # a library older than csrc/ or the header: argument lists and the ipsr_fwd_args layout may have moved
            if not build_if_missing:
                raise IpsrError("libipsr_sm100.so at %s is stale (sources changed) and building was disabled" % path)
            path = _build.build_library()
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError here == ABI drift, fail loudly
            fn.restype = res
            fn.argtypes = args
        if lib.ipsr_abi_fwd_args_bytes() != C.sizeof(FwdArgs):
            raise IpsrError("ABI drift: ipsr_fwd_args is %d bytes in %s but %d bytes in the ctypes mirror"
                            % (lib.ipsr_abi_fwd_args_bytes(), path, C.sizeof(FwdArgs)))
        _lib = lib
        return _lib


def last_error() -> str:
    return load().ipsr_last_error_string().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise IpsrError("%s failed (code %d): %s" % (what or "libipsr_sm100", rc, last_error()))


def call(name: str, *args):
    """Call an int-returning entry point and raise IpsrError on a non-zero code."""
    check(getattr(load(), name)(*args), name)
