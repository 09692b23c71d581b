"""Thin tensor-level wrappers over the C ABI (``include/ipsr_sm100.h``).

PyTorch is used for device memory, streams and autograd plumbing only; every computation is a
kernel of libipsr_sm100.so.  Nothing here falls back to torch ops or to the CPU: tensors must be
CUDA tensors and the library must load.
"""
from __future__ import annotations

import ctypes as C
import threading
import weakref
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib

# --------------------------------------------------------------------------------------
# configuration knobs of the drop-in (module level, like the reference's opt object)
# --------------------------------------------------------------------------------------
config = {
    "correlation_mode": "auto",   # "auto" | "tensor" | "exact"
    "tol_rel": -1.0,              # < 0: library default (5e-5 * ||R[q]||)
    "tol_abs": -1.0,
    "psplit": 0,                  # <= 0: auto
    "wide_blend": "blocked",      # long patch rows: "blocked" (one sequential scalar per step) | "stepwise" (one reduction per step)
    "fuse_innercos": True,        # compute the InnerCos loss that follows the layer (networks.py:347) inside the paste kernel
    "exc_cap_factor": 8,          # exception POOL of the batch = factor * N * B + M * min(M, N) entries (8 bytes each): signed
                                  # inputs make a few images chaotic (tens of thousands of attention entries survive the int64
                                  # store); they borrow the room of the others, and one image can always be fully chaotic.
                                  # Only when the whole pool is exhausted does the backward of the images that found no room
                                  # replay the recurrence per column (correct, slow)
}

EXC_REPLAY = 0x3FFFFFFF           # exc_total[b] >= this: the lists of image b are unusable, the backward replays


def exc_pool_entries(B: int, N: int, M: int) -> int:
    """Capacity of the exception pool of one call (entries of 8 bytes)."""
    if M <= 1:
        return 0
    f = float(config["exc_cap_factor"])
    shared = min(int(f * N * B), 1 << 26)
    single = min(M * min(M, N), 1 << 26) if f >= 1 else 0      # (tests shrink the pool with factors < 1)
    return max(1, shared + single)



_tls = threading.local()


def request_fused_cos(req) -> None:
    """IPSR_model.forward -> IPSRFunction.forward (whose 12-argument signature is the reference's and cannot carry it)."""
    _tls.fused_request = req


def take_fused_request():
    req = getattr(_tls, "fused_request", None)
    _tls.fused_request = None
    return req


def publish_fused_loss(loss) -> None:
    _tls.fused_loss = loss


def take_fused_loss():
    loss = getattr(_tls, "fused_loss", None)
    _tls.fused_loss = None
    return loss


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _require_cuda(t: torch.Tensor, name: str, dtype=None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: deepinpainting_b200 has no CPU path" % name)
    if dtype is not None and t.dtype != dtype:
        raise TypeError("%s must have dtype %s, got %s" % (name, dtype, t.dtype))
    return t.contiguous()


# --------------------------------------------------------------------------------------
# mask helpers
# --------------------------------------------------------------------------------------
def feat_mask(mask_2d: torch.Tensor, conv_layers: int, threshold: float) -> torch.Tensor:
    """util/util.py:68-84 on device.  mask_2d: [S_h, S_w] (bool/uint8/float, non-zero = hole)."""
    m = _require_cuda(mask_2d, "mask")
    if m.dim() != 2:
        raise ValueError("feat_mask expects a 2-D mask")
    m8 = (m != 0).to(torch.uint8).contiguous()
    sh, sw = m8.shape
    out = torch.empty((sh >> conv_layers, sw >> conv_layers), dtype=torch.uint8, device=m.device)
    scratch = torch.empty((2 * max(1, (sh // 2) * (sw // 2)),), dtype=torch.int32, device=m.device)
    _lib.call("ipsr_feat_mask", m8.data_ptr(), sh, sw, int(conv_layers), float(threshold), out.data_ptr(),
              scratch.data_ptr(), _stream_ptr(m.device))
    return out


def feat_mask_batched(masks: torch.Tensor, conv_layers: int, threshold: float) -> torch.Tensor:
    """util/util.py:68-84 for a batch of masks [B, S_h, S_w] in one launch per layer -> uint8 [B, S_h>>L, S_w>>L]."""
    m = _require_cuda(masks, "mask")
    if m.dim() != 3:
        raise ValueError("feat_mask_batched expects [B, S_h, S_w]")
    m8 = (m != 0).to(torch.uint8).contiguous()
    B, sh, sw = m8.shape
    out = torch.empty((B, sh >> conv_layers, sw >> conv_layers), dtype=torch.uint8, device=m.device)
    scratch = torch.empty((2 * B * max(1, (sh // 2) * (sw // 2)),), dtype=torch.int32, device=m.device)
    _lib.call("ipsr_feat_mask_batch", m8.data_ptr(), B, sh, sw, int(conv_layers), float(threshold), out.data_ptr(),
              scratch.data_ptr(), _stream_ptr(m.device))
    return out


@dataclass
class MaskIndex:
    """Device-side flag / index vectors of one mask (util/util.py:88-147)."""
    flag: torch.Tensor       # int32 [N]
    mask_idx: torch.Tensor   # int32 [M]
    rank: torch.Tensor       # int32 [N]  position inside mask_idx or -1
    M: int
    nH: int
    nW: int
    # per-image masks (extension of the reference's one mask per batch): flag / mask_idx / rank are [B, N],
    # m_count [B] holds the masked positions per image and M their maximum
    m_count: Optional[torch.Tensor] = None

    @property
    def batched(self) -> bool:
        return self.m_count is not None


def build_flags(feat: torch.Tensor, patch: int, stride: int, mask_thred: int) -> MaskIndex:
    f = _require_cuda(feat, "mask", torch.uint8)
    if f.dim() != 2:
        raise AssertionError("mask has to be 2 dimenison!")
    H, W = f.shape
    nH, nW = (H - patch) // stride + 1, (W - patch) // stride + 1
    P = nH * nW
    dev = f.device
    flag = torch.empty(P, dtype=torch.int32, device=dev)
    midx = torch.empty(P, dtype=torch.int32, device=dev)
    rank = torch.empty(P, dtype=torch.int32, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    _lib.call("ipsr_build_flags", f.data_ptr(), H, W, int(patch), int(stride), int(mask_thred), flag.data_ptr(),
              midx.data_ptr(), rank.data_ptr(), count.data_ptr(), _stream_ptr(dev))
    M = int(count.item())            # one host sync per NEW mask, never per forward
    return MaskIndex(flag=flag, mask_idx=midx[:M].contiguous(), rank=rank, M=M, nH=nH, nW=nW)


def build_flags_batched(feats: torch.Tensor, patch: int, stride: int, mask_thred: int) -> MaskIndex:
    """One MaskIndex for a batch whose images have their OWN masks (feats uint8 [B, H, W]) in ONE launch and ONE host
    synchronisation (the largest count sizes the per-step buffers) -- not one of each per sample."""
    f = _require_cuda(feats, "masks", torch.uint8)
    if f.dim() != 3:
        raise ValueError("build_flags_batched expects [B, H, W]")
    B, H, W = f.shape
    nH, nW = (H - patch) // stride + 1, (W - patch) // stride + 1
    P = nH * nW
    dev = f.device
    flag = torch.empty((B, P), dtype=torch.int32, device=dev)
    midx = torch.zeros((B, P), dtype=torch.int32, device=dev)
    rank = torch.empty((B, P), dtype=torch.int32, device=dev)
    count = torch.empty((B,), dtype=torch.int32, device=dev)
    _lib.call("ipsr_build_flags_batch", f.data_ptr(), B, H, W, int(patch), int(stride), int(mask_thred), flag.data_ptr(),
              midx.data_ptr(), rank.data_ptr(), count.data_ptr(), _stream_ptr(dev))
    M = int(count.max().item())
    return MaskIndex(flag=flag, mask_idx=midx, rank=rank, M=M, nH=nH, nW=nW, m_count=count)


def stack_mask_indices(items) -> MaskIndex:
    """One MaskIndex for a batch whose images have their OWN masks: rows of flag / mask_idx / rank per image."""
    items = list(items)
    N = items[0].flag.numel()
    dev = items[0].flag.device
    if any(mi.flag.numel() != N or mi.batched for mi in items):
        raise ValueError("per-image masks must share the feature-map size")
    B = len(items)
    midx = torch.zeros((B, N), dtype=torch.int32, device=dev)
    for b, mi in enumerate(items):
        if mi.M:
            midx[b, :mi.M] = mi.mask_idx
    return MaskIndex(flag=torch.stack([mi.flag for mi in items]).contiguous(), mask_idx=midx,
                     rank=torch.stack([mi.rank for mi in items]).contiguous(), M=max(mi.M for mi in items),
                     nH=items[0].nH, nW=items[0].nW,
                     m_count=torch.tensor([mi.M for mi in items], dtype=torch.int32, device=dev))


_mask_registry = {}


def register_mask_index(flag: torch.Tensor, mi: MaskIndex) -> None:
    """Remember the device vectors that belong to a reference-style ``flag`` tensor (keyed by object
    identity, dropped when the tensor dies) so IPSRFunction.apply need not rebuild them."""
    key = id(flag)
    _mask_registry[key] = mi
    weakref.finalize(flag, _mask_registry.pop, key, None)


def lookup_mask_index(flag: torch.Tensor, device) -> MaskIndex:
    mi = _mask_registry.get(id(flag))
    if mi is not None and mi.flag.device == torch.device(device):
        return mi
    mi = mask_index_from_flag(flag, device)
    register_mask_index(flag, mi)
    return mi


def mask_index_from_flag(flag: torch.Tensor, device) -> MaskIndex:
    """Device vectors from a reference-style ``flag`` vector (any integer dtype, any device):
    the flag is its own 1 x N feature mask with 1 x 1 patches.  A 2-D ``flag`` [B, N] holds one mask per image."""
    if flag.dim() == 2:
        return stack_mask_indices(mask_index_from_flag(row, device) for row in flag)
    f8 = (flag.to(device=device) != 0).to(torch.uint8).reshape(1, -1).contiguous()
    return build_flags(f8, 1, 1, 1)


# --------------------------------------------------------------------------------------
# shift operator
# --------------------------------------------------------------------------------------
@dataclass
class FusedCos:
    """Request for the InnerCos side loss of the module that follows the shift layer (models/networks.py:347), computed
    in the paste kernel while the pasted tiles are still in shared memory instead of a second pass over the output."""
    target: torch.Tensor      # [B,C,H,W] fp32
    mask: torch.Tensor        # [H,W] or [N] fp32, 1 = hole
    strength: float
    crit: str                 # 'MSE' or anything else for L1, as the reference's constructor argument


@dataclass
class ShiftSaved:
    B: int
    C: int
    H: int
    W: int
    M: int
    ind: torch.Tensor
    wn: Optional[torch.Tensor]
    wo: Optional[torch.Tensor]
    mask_idx: Optional[torch.Tensor]
    route_ptr: Optional[torch.Tensor] = None
    route_q: Optional[torch.Tensor] = None
    exc_start: Optional[torch.Tensor] = None
    exc_cnt: Optional[torch.Tensor] = None
    exc_l: Optional[torch.Tensor] = None
    exc_w: Optional[torch.Tensor] = None
    exc_total: Optional[torch.Tensor] = None    # [B] entries per image (>= EXC_REPLAY: replay); view of exc_state
    exc_state: Optional[torch.Tensor] = None    # [2B + 2]: counts, bases inside the pool, pool cursor
    exc_cap: int = 0                            # pool capacity (entries, shared by the batch)
    nrecheck: Optional[torch.Tensor] = None
    npass2: Optional[torch.Tensor] = None
    m_count: Optional[torch.Tensor] = None      # per-image masks: mask_idx is [B, N], m_count [B]
    cos_loss: Optional[torch.Tensor] = None     # fused InnerCos side loss (scalar), when requested

    def exceptions_of(self, b: int):
        """(positions q_l, truncated weights) of the attention entries of image ``b`` that survive the reference's int64
        store (rows l >= 1), or None when the image's lists are unusable and the backward replays (host sync: tests)."""
        if self.exc_state is None:
            return (torch.empty(0, dtype=torch.int32, device=self.ind.device),
                    torch.empty(0, dtype=torch.float32, device=self.ind.device))
        st = self.exc_state.cpu()
        n, base = int(st[b]), int(st[self.B + b])
        if n >= EXC_REPLAY:
            return None
        return self.exc_l[base:base + n], self.exc_w[base:base + n]


def _carve(blob: torch.Tensor, layout, name):
    off, shape, dtype = layout[name]
    n = 1
    for d in shape:
        n *= d
    return blob[off:off + n * 4].view(dtype).view(shape)


class _Plan:
    """Everything of a forward call that depends only on (device, stream, shape, mask, mode): the workspace, the filled
    argument struct and the layout of the per-call blob that holds what the backward needs.  Cached per thread, so that
    a forward through the module API costs two allocations (output + blob) and ONE ctypes call."""

    def __init__(self, dev, B, Cc, H, W, mi: MaskIndex, need_grad, mode, col_begin, col_end, stop_after_corr, diagnostics):
        N, M = H * W, mi.M
        if mi.batched:
            if tuple(mi.flag.shape) != (B, N):
                raise ValueError("per-image flags are %s but the batch is %d images of %d positions" % (tuple(mi.flag.shape), B, N))
        elif mi.flag.numel() != N:
            raise ValueError("flag has %d entries but the feature map has %d positions" % (mi.flag.numel(), N))
        self.B, self.C, self.H, self.W, self.M, self.N = B, Cc, H, W, M, N
        self.mi, self.need_grad, self.diagnostics = mi, need_grad, diagnostics
        Mr = max(M, 1)
        layout, cur = {}, 0

        def add(name, shape, dtype):
            nonlocal cur
            n = 1
            for d in shape:
                n *= d
            layout[name] = (cur, shape, dtype)
            cur += (n * 4 + 255) & ~255

        add("ind", (B, N), torch.int32)
        add("wn", (B, Mr), torch.float32)
        add("wo", (B, Mr), torch.float32)
        self.exc_cap = 0
        if need_grad:
            add("route_ptr", (B, N + 1), torch.int32)
            add("route_q", (B, N), torch.int32)
            if M > 1:
                self.exc_cap = exc_pool_entries(B, N, M)
                add("exc_start", (B, N), torch.int32)
                add("exc_cnt", (B, N), torch.int32)
                add("exc_state", (2 * B + 2,), torch.int32)
                add("exc_l", (self.exc_cap,), torch.int32)
                add("exc_w", (self.exc_cap,), torch.float32)
        if diagnostics:
            add("nrecheck", (B,), torch.int32)
            add("npass2", (B,), torch.int32)
        self.layout, self.blob_bytes = layout, cur
        lib = _lib.load()
        mode_id = _lib.MODES[mode]
        self.workspace_bytes = lib.ipsr_workspace_bytes(B, Cc, H, W, M, mode_id)
        self.workspace = torch.empty((self.workspace_bytes,), dtype=torch.uint8, device=dev)
        a = _lib.FwdArgs()
        a.flag, a.mask_idx, a.rank = mi.flag.data_ptr(), _ptr(mi.mask_idx) if M else None, mi.rank.data_ptr()
        a.B, a.C, a.H, a.W, a.M = B, Cc, H, W, M
        a.mode, a.need_grad = mode_id, int(bool(need_grad))
        a.col_begin, a.col_end, a.stop_after_corr = int(col_begin), int(col_end), int(bool(stop_after_corr))
        a.exc_cap = self.exc_cap
        a.workspace, a.workspace_bytes = self.workspace.data_ptr(), self.workspace_bytes
        a.mask_stride, a.m_count = (N, mi.m_count.data_ptr()) if mi.batched else (0, None)
        self.args = a
        self._cos_scratch = None

    def cos_scratch(self, dev):
        """Partials + the (zeroed, left zeroed) ticket of the fused InnerCos loss; stream-ordered reuse across calls."""
        if self._cos_scratch is None:
            n = _lib.load().ipsr_paste_loss_partials(self.B, self.C, self.N)
            self._cos_scratch = (torch.empty(n, dtype=torch.float32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev))
        return self._cos_scratch


_plans = threading.local()
PLAN_CACHE_SIZE = 16


def _plan_for(x, mi, need_grad, mode, col_begin, col_end, stop_after_corr, diagnostics) -> _Plan:
    B, Cc, H, W = x.shape
    dev = x.device
    capturing = torch.cuda.is_current_stream_capturing()
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream, B, Cc, H, W, id(mi), bool(need_grad), mode, int(col_begin),
           int(col_end), bool(stop_after_corr), bool(diagnostics))
    cache = getattr(_plans, "cache", None)
    if cache is None:
        cache = _plans.cache = {}
    plan = None if capturing else cache.get(key)      # a workspace captured into a CUDA graph belongs to that graph
    if plan is not None and plan.mi is mi:
        return plan
    plan = _Plan(dev, B, Cc, H, W, mi, need_grad, mode, col_begin, col_end, stop_after_corr, diagnostics)
    if not capturing:
        if len(cache) >= PLAN_CACHE_SIZE:
            cache.pop(next(iter(cache)))
        cache[key] = plan
    return plan


class _Call:
    """One forward call: the output, the blob with what the backward needs, and the plan's argument struct pointed at them."""

    def __init__(self, x, ref, mi: MaskIndex, need_grad, mode, col_begin, col_end, stop_after_corr, diagnostics,
                 events=None, fused_cos: Optional[FusedCos] = None):
        plan = _plan_for(x, mi, need_grad, mode, col_begin, col_end, stop_after_corr, diagnostics)
        dev = x.device
        self.plan = plan
        self.out = torch.empty_like(x)
        blob = torch.empty((plan.blob_bytes,), dtype=torch.uint8, device=dev)
        lay = plan.layout
        s = ShiftSaved(B=plan.B, C=plan.C, H=plan.H, W=plan.W, M=plan.M, ind=_carve(blob, lay, "ind"), wn=_carve(blob, lay, "wn"),
                       wo=_carve(blob, lay, "wo"), mask_idx=mi.mask_idx, m_count=mi.m_count, exc_cap=plan.exc_cap)
        if "route_ptr" in lay:
            s.route_ptr, s.route_q = _carve(blob, lay, "route_ptr"), _carve(blob, lay, "route_q")
        if "exc_state" in lay:
            s.exc_start, s.exc_cnt = _carve(blob, lay, "exc_start"), _carve(blob, lay, "exc_cnt")
            s.exc_l, s.exc_w = _carve(blob, lay, "exc_l"), _carve(blob, lay, "exc_w")
            s.exc_state = _carve(blob, lay, "exc_state")
            s.exc_total = s.exc_state[:plan.B]
        if "nrecheck" in lay:
            s.nrecheck, s.npass2 = _carve(blob, lay, "nrecheck"), _carve(blob, lay, "npass2")
            s.npass2.zero_()
        self.saved = s
        a = plan.args
        a.x, a.ref = x.data_ptr(), ref.data_ptr()
        a.psplit = int(config["psplit"])
        a.tol_rel, a.tol_abs = float(config["tol_rel"]), float(config["tol_abs"])
        a.out, a.ind, a.wn, a.wo = self.out.data_ptr(), s.ind.data_ptr(), s.wn.data_ptr(), s.wo.data_ptr()
        a.route_ptr, a.route_q = _ptr(s.route_ptr), _ptr(s.route_q)
        a.exc_start, a.exc_cnt, a.exc_l, a.exc_w, a.exc_total = (_ptr(s.exc_start), _ptr(s.exc_cnt), _ptr(s.exc_l),
                                                               _ptr(s.exc_w), _ptr(s.exc_state))
        a.nrecheck_out = _ptr(s.nrecheck)
        a.npass2_out = _ptr(s.npass2)
        if fused_cos is not None:
            t = _require_cuda(fused_cos.target, "InnerCos target", torch.float32)
            m = _require_cuda(fused_cos.mask, "InnerCos mask", torch.float32)
            if tuple(t.shape) != tuple(x.shape) or m.numel() != plan.N:
                raise ValueError("fused InnerCos: target %s / mask %s do not match the layer output %s"
                                 % (tuple(t.shape), tuple(m.shape), tuple(x.shape)))
            partials, ticket = plan.cos_scratch(dev)
            s.cos_loss = torch.empty((), dtype=torch.float32, device=dev)
            a.cos_target, a.cos_mask = t.data_ptr(), m.data_ptr()
            a.cos_strength, a.cos_crit = float(fused_cos.strength), 0 if fused_cos.crit == "MSE" else 1
            a.cos_partials, a.cos_ticket, a.cos_loss = partials.data_ptr(), ticket.data_ptr(), s.cos_loss.data_ptr()
            self.keep_cos = (t, m)
        else:
            a.cos_target = a.cos_mask = a.cos_partials = a.cos_ticket = a.cos_loss = None
        if events is not None:                       # (begin, end) torch.cuda.Event pair, already materialised
            a.ev_corr_begin, a.ev_corr_end = events[0].cuda_event, events[1].cuda_event
        else:
            a.ev_corr_begin, a.ev_corr_end = None, None
        self.args = a
        self.workspace = plan.workspace
        self.keep = (x, ref, mi, blob)
        self.device = dev

    def packed_keys(self) -> torch.Tensor:
        """[B,N] int64 view of the exchange keys inside the workspace (bank-sharded mode)."""
        lib = _lib.load()
        addr = lib.ipsr_workspace_packed(C.byref(self.args))
        off = addr - self.workspace.data_ptr()
        n = self.saved.B * self.saved.H * self.saved.W
        return self.workspace[off:off + 8 * n].view(torch.int64).view(self.saved.B, -1)


def shift_forward(x: torch.Tensor, ref: torch.Tensor, mi: MaskIndex, need_grad: bool = True,
                  mode: Optional[str] = None, diagnostics: bool = False, events=None, fused_cos: Optional[FusedCos] = None):
    """models/IPSRFunction.py:13-140 for shift_sz = stride = 1.  Returns (out, ShiftSaved).  ``fused_cos``: also compute
    the InnerCos side loss of the output (``saved.cos_loss``) inside the paste kernel."""
    x = _require_cuda(x, "input", torch.float32)
    ref = _require_cuda(ref, "ref.relu4_3", torch.float32)
    if x.dim() != 4:
        raise AssertionError("Input Dim has to be 4")
    if ref.shape != x.shape:
        raise ValueError("ref.relu4_3 %s must have the shape of the input %s" % (tuple(ref.shape), tuple(x.shape)))
    call = _Call(x, ref, mi, need_grad, mode or config["correlation_mode"], 0, -1, False, diagnostics, events, fused_cos)
    _lib.call("ipsr_shift_forward", C.byref(call.args), _stream_ptr(x.device))
    return call.out, call.saved


def launches_per_step(C: int, N: int, M: int, need_grad: bool = True, mode: Optional[str] = None, backward: bool = True,
                      B: int = 1) -> int:
    """Number of libipsr_sm100 KERNEL launches of one forward (+ backward) -- what bench.py reports as
    gpu_launches (memsets / event records are not kernels)."""
    mode = mode or config["correlation_mode"]
    tensor = mode == "tensor" or (mode == "auto" and _lib.load().ipsr_tensor_path_supported(C, N) == 1)
    n = 1                                   # extract_normalize
    cascade = tensor and _lib.load().ipsr_tensor_cascade(B, C, N) == 1
    if tensor and not cascade:
        # small problems: without a column split the epilogue of the (single) pass makes the finalize decision itself
        tiles = B * (N // 128)
        ps = config["psplit"] if config["psplit"] > 0 else max(1, min(8, 148 // max(1, tiles), N // 128))
        n += 1 if ps == 1 else 2
    else:
        n += 4 if tensor else 1               # (correlate_tc + finalize) x 2 with the cascade | select_all_rows
    n += 1 if tensor else 2                 # whole bank, tensor mode: recheck + resolve in one launch | correlate_fp32 + resolve_rows
    if M > 0:
        n += 2                              # blend_stage (+ routes builders in the same launch) + blend_scan
    elif need_grad:
        n += 1                              # build_routes
    n += 1                                  # paste
    if need_grad and M > 1:
        n += 1                              # build_exceptions (side stream, next to the paste)
    if backward:
        n += 1                              # shift_bwd
    return n


def shift_forward_sharded(x, ref, mi: MaskIndex, col_begin: int, col_end: int, reduce_max, need_grad: bool = True,
                          mode: Optional[str] = None):
    """Bank-sharded forward: this rank correlates against bank columns [col_begin, col_end) only;
    ``reduce_max(int64 tensor)`` performs the one exchange step (all-reduce MAX, in place) of the
    packed (score, index) keys; everything after it runs replicated."""
    x = _require_cuda(x, "input", torch.float32)
    ref = _require_cuda(ref, "ref.relu4_3", torch.float32)
    call = _Call(x, ref, mi, need_grad, mode or config["correlation_mode"], col_begin, col_end, True, False)
    st = _stream_ptr(x.device)
    _lib.call("ipsr_shift_forward", C.byref(call.args), st)
    keys = call.packed_keys()
    reduce_max(keys)
    call.args.stop_after_corr = 0
    _lib.call("ipsr_shift_forward_finish", C.byref(call.args), st)
    return call.out, call.saved


# patch rows longer than this take the wide kernels (the shared-memory tiles of the 1 x 1 kernels hold <= 1024 channels)
PATCH_ROW_LIMIT = 1024


def patch_grid(H: int, W: int, patch: int, stride: int):
    """(nH, nW) patch positions (util/NonparametricShift.py:63-64)."""
    return (H - patch) // stride + 1, (W - patch) // stride + 1


def shift_forward_patches(x: torch.Tensor, ref: torch.Tensor, mi: MaskIndex, patch: int, stride: int,
                          mode: Optional[str] = None, col_begin: int = 0, col_end: int = -1, reduce_max=None):
    """models/IPSRFunction.py:46-133 for shift_sz = patch, stride = stride -- FORWARD ONLY (the reference computes
    the output and then fails at :134; its backward is undefined for these settings).  ``mi`` holds the flag
    vectors over the nH x nW patch positions (util.cal_mask_given_mask_thred with the same patch / stride).
    ``reduce_max`` (with ``col_begin`` / ``col_end``): bank-sharded mode, as in ``shift_forward_sharded``.
    Returns (out [B,C,H,W], ind [B,P] int32)."""
    x = _require_cuda(x, "input", torch.float32)
    ref = _require_cuda(ref, "ref.relu4_3", torch.float32)
    if x.dim() != 4:
        raise AssertionError("Input Dim has to be 4")
    if ref.shape != x.shape:
        raise ValueError("ref.relu4_3 %s must have the shape of the input %s" % (tuple(ref.shape), tuple(x.shape)))
    B, Cc, H, W = x.shape
    k, s = int(patch), int(stride)
    nH, nW = patch_grid(H, W, k, s)
    if nH < 1 or nW < 1 or (nH - 1) * s + k != H or (nW - 1) * s + k != W:
        # the reference's conv-transpose output no longer has the input's shape (IPSRFunction.py:133 raises)
        raise RuntimeError("shift_sz=%d / stride=%d patches do not tile a %d x %d feature map" % (k, s, H, W))
    P = nH * nW
    if mi.batched:
        raise NotImplementedError("per-image masks are implemented for shift_sz = stride = 1 only")
    if mi.flag.numel() != P:
        raise ValueError("flag has %d entries but there are %d patch positions" % (mi.flag.numel(), P))
    dev, st = x.device, _stream_ptr(x.device)
    K = Cc * k * k
    sharded = reduce_max is not None
    out = torch.empty_like(x)
    if K <= PATCH_ROW_LIMIT:
        # patch maps are feature maps with K channels on the nH x nW grid: the 1 x 1 pipeline runs on them unchanged
        Kpad = -(-K // 64) * 64 if (P % 128 == 0 and -(-K // 64) * 64 <= PATCH_ROW_LIMIT) else -(-K // 32) * 32
        cols_x = torch.empty((B, Kpad, nH, nW), dtype=torch.float32, device=dev)
        cols_r = torch.empty((B, Kpad, nH, nW), dtype=torch.float32, device=dev)
        _lib.call("ipsr_unfold_patches", x.data_ptr(), B, Cc, H, W, k, s, Kpad, cols_x.data_ptr(), st)
        _lib.call("ipsr_unfold_patches", ref.data_ptr(), B, Cc, H, W, k, s, Kpad, cols_r.data_ptr(), st)
        if sharded:
            cols_o, saved = shift_forward_sharded(cols_x, cols_r, mi, col_begin, col_end, reduce_max, need_grad=False, mode=mode)
        else:
            cols_o, saved = shift_forward(cols_x, cols_r, mi, need_grad=False, mode=mode)
        _lib.call("ipsr_fold_patches", cols_o.data_ptr(), B, Cc, H, W, k, s, Kpad, out.data_ptr(), st)
        return out, saved.ind
    # wide rows: position-major patches + norms; correlation on the tensor cores (or exact fp32 on the patch maps);
    # register-resident blend
    mode_eff = mode or config["correlation_mode"]
    rows = torch.empty((B, P, K), dtype=torch.float32, device=dev)
    inv = torch.empty((B, P), dtype=torch.float32, device=dev)
    ind = torch.empty((B, P), dtype=torch.int32, device=dev)
    vmax = torch.empty((B, P), dtype=torch.float32, device=dev)
    cb, ce = (int(col_begin), P if (col_end < 0 or (col_end == 0 and col_begin == 0)) else int(col_end)) if sharded else (0, P)
    if mode_eff != "exact":
        ind, vmax = _wide_patch_correlation_tc(x, ref, B, Cc, H, W, k, s, P, K, rows, inv, cb, ce, reduce_max if sharded else None, st)
    else:
        cols_x = torch.empty((B, K, P), dtype=torch.float32, device=dev)
        cols_r = torch.empty((B, K, P), dtype=torch.float32, device=dev)
        _lib.call("ipsr_unfold_patches", x.data_ptr(), B, Cc, H, W, k, s, K, cols_x.data_ptr(), st)
        _lib.call("ipsr_unfold_patches", ref.data_ptr(), B, Cc, H, W, k, s, K, cols_r.data_ptr(), st)
        _lib.call("ipsr_patch_rows", x.data_ptr(), B, Cc, H, W, k, s, rows.data_ptr(), inv.data_ptr(), st)
        packed = torch.empty((B, P), dtype=torch.int64, device=dev)
        rlist = torch.empty((B, P), dtype=torch.int32, device=dev)
        nrow = torch.empty((B,), dtype=torch.int32, device=dev)
        _lib.call("ipsr_select_all_rows", B, P, rlist.data_ptr(), nrow.data_ptr(), packed.data_ptr(), st)
        if cb < ce:                                   # an empty shard (more ranks than columns) contributes identity keys
            _lib.call("ipsr_correlate_argmax_fp32", cols_x.data_ptr(), cols_r.data_ptr(), inv.data_ptr(), B, K, P, cb, ce,
                      rlist.data_ptr(), nrow.data_ptr(), (P + 63) // 64, packed.data_ptr(), st)
        if sharded:
            reduce_max(packed)
        _lib.call("ipsr_unpack_maxidx", packed.data_ptr(), B * P, vmax.data_ptr(), ind.data_ptr(), st)
    M = mi.M
    y = None
    if M > 0:
        y = torch.empty((B, M, K), dtype=torch.float32, device=dev)
        wn = torch.empty((B, M), dtype=torch.float32, device=dev)
        wo = torch.empty((B, M), dtype=torch.float32, device=dev)
        if config["wide_blend"] == "blocked":
            gram = torch.empty((_lib.load().ipsr_blend_wide_gram_floats(B, M),), dtype=torch.float32, device=dev)
            _lib.call("ipsr_blend_wide_blocked", rows.data_ptr(), inv.data_ptr(), vmax.data_ptr(), ind.data_ptr(),
                      mi.mask_idx.data_ptr(), B, K, P, M, gram.data_ptr(), y.data_ptr(), wn.data_ptr(), wo.data_ptr(), st)
        else:
            _lib.call("ipsr_blend_wide", rows.data_ptr(), inv.data_ptr(), vmax.data_ptr(), ind.data_ptr(), mi.mask_idx.data_ptr(),
                      B, K, P, M, y.data_ptr(), wn.data_ptr(), wo.data_ptr(), st)
    _lib.call("ipsr_fold_patch_rows", rows.data_ptr(), _ptr(y), ind.data_ptr(), mi.rank.data_ptr(), B, Cc, H, W, k, s, M,
              out.data_ptr(), st)
    return out, ind


# error band of the three-pass fp16 split for LONG rows, relative to ||R[q]||: the operand rounding of the split (1.2e-6)
# plus the fp32 accumulation of K * 3 products in tensor memory (8e-6 covers K <= 512; it grows linearly with K at worst).
# Rows whose top-2 gap lies inside twice the band are recomputed in exact fp32.
def _wide_tol_rel(K: int) -> float:
    return 2.0 * (8e-6 * max(1.0, K / 512.0) + 1.2e-6)


def time_wide_patch_correlation(x, ref, patch: int, stride: int, col_begin: int, col_end: int, reps: int = 10):
    """Measurement helper (bench.py): the stages of the long-row tensor route timed alone with CUDA events on the current
    stream -- patch rows + statistics, operand images, the tcgen05 GEMM, finalize + resolve + winner scores."""
    B, Cc, H, W = x.shape
    nH, nW = patch_grid(H, W, patch, stride)
    P, K = nH * nW, Cc * patch * patch
    dev = x.device
    st = _stream_ptr(dev)
    rows = torch.empty((B, P, K), dtype=torch.float32, device=dev)
    inv = torch.empty((B, P), dtype=torch.float32, device=dev)
    marks = {}
    out = {"launches_per_step": 13}
    for r in range(reps + 2):
        marks.clear()
        _wide_patch_correlation_tc(x, ref, B, Cc, H, W, patch, stride, P, K, rows, inv, col_begin, col_end, None, st, marks)
        torch.cuda.synchronize(dev)
        if r >= 2:
            names = list(marks)
            for a, b in zip(names[:-1], names[1:]):
                out[b + "_ms"] = out.get(b + "_ms", 0.0) + marks[a].elapsed_time(marks[b]) / reps
    return out


def _wide_patch_correlation_tc(x, ref, B, Cc, H, W, k, s, P, K, rows, inv, cb, ce, reduce_max, st, marks=None):
    """models/IPSRFunction.py:54-65 for long patch rows on the tcgen05 path (ipsr_patch_tc.cu).  Fills ``rows`` / ``inv``
    (the patches of x) and returns (ind [B,P] int32, vmax [B,P])."""
    dev = x.device

    def mark(name):
        if marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks[name] = ev

    Kpad, Ppad = -(-K // 64) * 64, -(-P // 128) * 128
    f32 = dict(dtype=torch.float32, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    rows_r = torch.empty((B, P, K), **f32)
    mark("start")
    rnorm, rmax = torch.empty((B, P), **f32), torch.empty((B, P), **f32)
    _lib.call("ipsr_patch_rows_stats", x.data_ptr(), B, Cc, H, W, k, s, rows.data_ptr(), inv.data_ptr(), None, None, st)
    _lib.call("ipsr_patch_rows_stats", ref.data_ptr(), B, Cc, H, W, k, s, rows_r.data_ptr(), None, rnorm.data_ptr(), rmax.data_ptr(), st)
    mark("rows")
    tile_bytes = B * (Kpad // 64) * 2 * (Ppad // 128) * 16384
    x_tiles = torch.empty((tile_bytes,), dtype=torch.uint8, device=dev)
    r_tiles = torch.empty((tile_bytes,), dtype=torch.uint8, device=dev)
    rscale, rnorm_pad = torch.empty((B, Ppad), **f32), torch.empty((B, Ppad), **f32)
    _lib.call("ipsr_patch_tiles", rows.data_ptr(), inv.data_ptr(), None, None, 0, B, K, P, x_tiles.data_ptr(), None, None, st)
    _lib.call("ipsr_patch_tiles", rows_r.data_ptr(), None, rmax.data_ptr(), rnorm.data_ptr(), 1, B, K, P, r_tiles.data_ptr(),
              rscale.data_ptr(), rnorm_pad.data_ptr(), st)
    mark("tiles")
    ind_pad = torch.empty((B, Ppad), **i32)
    packed_pad = torch.empty((B, Ppad), dtype=torch.int64, device=dev)
    rlist = torch.empty((B, Ppad), **i32)
    nlist = torch.zeros((B,), **i32)
    ind = torch.empty((B, P), **i32)
    vmax = torch.empty((B, P), **f32)
    keys = torch.empty((B, P), dtype=torch.int64, device=dev)
    if cb < ce:
        # this rank's bank columns, in whole 128-column tiles (the last tile may run into the padding, masked by n_valid)
        if cb % 128 != 0 or (ce % 128 != 0 and ce != P):
            raise ValueError("bank shards of the tensor path must be 128-column aligned, got [%d, %d)" % (cb, ce))
        ce_pad = min(Ppad, -(-ce // 128) * 128)
        tiles_rows = B * (Ppad // 128)
        psplit = max(1, min(8, 148 // max(1, tiles_rows), (ce_pad - cb) // 128))
        parts = [torch.empty((psplit, B, Ppad), **f32), torch.empty((psplit, B, Ppad), **i32), torch.empty((psplit, B, Ppad), **f32),
                 torch.empty((psplit, B, Ppad), **i32), torch.empty((psplit, B, Ppad), **f32)]
        cand2, pair_list, npair = torch.empty((B, Ppad), **i32), torch.empty((B, Ppad), **i32), torch.zeros((B,), **i32)
        _lib.call("ipsr_correlate_argmax_tc_valid", r_tiles.data_ptr(), x_tiles.data_ptr(), B, Kpad, Ppad, cb, ce_pad, psplit, 3, 2,
                  None, parts[0].data_ptr(), parts[1].data_ptr(), parts[2].data_ptr(), parts[3].data_ptr(), parts[4].data_ptr(),
                  None, min(ce, P), st)
        mark("gemm")
        # rows inside the error band: exactly two candidates -> two exact dot products; three or more (ties, duplicates)
        # -> the whole bank in fp32
        _lib.call("ipsr_finalize_argmax_valid", parts[0].data_ptr(), parts[1].data_ptr(), parts[2].data_ptr(), psplit,
                  rnorm_pad.data_ptr(), rscale.data_ptr(), None, None, None, None, None, B, Ppad, _wide_tol_rel(K), 1e-12,
                  ind_pad.data_ptr(), rlist.data_ptr(), nlist.data_ptr(), packed_pad.data_ptr(), None, None, Kpad,
                  parts[3].data_ptr(), parts[4].data_ptr(), cand2.data_ptr(), pair_list.data_ptr(), npair.data_ptr(), P, st)
        _lib.call("ipsr_patch_recheck", rows.data_ptr(), rows_r.data_ptr(), inv.data_ptr(), B, K, P, cb, ce, rlist.data_ptr(),
                  nlist.data_ptr(), packed_pad.data_ptr(), st)
        _lib.call("ipsr_apply_recheck", packed_pad.data_ptr(), rlist.data_ptr(), nlist.data_ptr(), B, Ppad, ind_pad.data_ptr(), None, st)
        _lib.call("ipsr_patch_resolve_pairs", rows.data_ptr(), rows_r.data_ptr(), inv.data_ptr(), B, K, P, pair_list.data_ptr(),
                  npair.data_ptr(), cand2.data_ptr(), ind_pad.data_ptr(), packed_pad.data_ptr(), st)
        _lib.call("ipsr_patch_winner_scores", rows.data_ptr(), rows_r.data_ptr(), inv.data_ptr(), ind_pad.data_ptr(),
                  packed_pad.data_ptr(), B, K, P, ind.data_ptr(), vmax.data_ptr(), keys.data_ptr(), st)
        mark("resolve")
    else:
        keys.fill_(-(1 << 63))                        # empty shard: identity keys
    if reduce_max is not None:
        reduce_max(keys)
        _lib.call("ipsr_unpack_maxidx", keys.data_ptr(), B * P, vmax.data_ptr(), ind.data_ptr(), st)
    return ind, vmax


def patch_rows(img: torch.Tensor, patch: int, stride: int):
    """util/NonparametricShift.py:59-68: all patches of ``img`` [B,C,H,W] as rows [B, P, C*k*k] in unfold order,
    plus 1/(||patch|| + 1e-8) [B,P] (:40)."""
    img = _require_cuda(img, "target_img", torch.float32)
    B, Cc, H, W = img.shape
    nH, nW = patch_grid(H, W, patch, stride)
    rows = torch.empty((B, nH * nW, Cc * patch * patch), dtype=torch.float32, device=img.device)
    inv = torch.empty((B, nH * nW), dtype=torch.float32, device=img.device)
    _lib.call("ipsr_patch_rows", img.data_ptr(), B, Cc, H, W, int(patch), int(stride), rows.data_ptr(), inv.data_ptr(),
              _stream_ptr(img.device))
    return rows, inv


def shift_backward(grad_out: torch.Tensor, saved: ShiftSaved, triple_w: float) -> torch.Tensor:
    """models/IPSRFunction.py:144-178."""
    g = _require_cuda(grad_out, "grad_output", torch.float32)
    s = saved
    if s.route_ptr is None:
        raise RuntimeError("forward was run with need_grad=False: nothing saved for backward")
    gin = torch.empty_like(g)
    N = s.H * s.W
    _lib.call("ipsr_shift_bwd_masks", g.data_ptr(), s.B, s.C, N, s.M, s.route_ptr.data_ptr(), s.route_q.data_ptr(),
              _ptr(s.exc_start), _ptr(s.exc_cnt), _ptr(s.exc_l), _ptr(s.exc_w), _ptr(s.exc_state), s.exc_cap,
              s.ind.data_ptr(), _ptr(s.mask_idx) if s.M else None, s.wn.data_ptr(), s.wo.data_ptr(),
              float(triple_w), gin.data_ptr(), N if s.m_count is not None else 0, _ptr(s.m_count), _stream_ptr(g.device))
    return gin


# --------------------------------------------------------------------------------------
# InnerCos side loss
# --------------------------------------------------------------------------------------
def innercos_loss(x: torch.Tensor, mask_f32: torch.Tensor, target: torch.Tensor, strength: float, crit: str,
                  c_limit: Optional[int] = None) -> torch.Tensor:
    x = _require_cuda(x, "in_data", torch.float32)
    t = _require_cuda(target, "target", torch.float32)
    m = _require_cuda(mask_f32, "mask", torch.float32)
    B, Ct, H, W = x.shape
    cl = Ct if c_limit is None else min(int(c_limit), Ct)
    if tuple(t.shape) != (B, cl, H, W):
        raise ValueError("target %s does not match the masked activations %s" % (tuple(t.shape), (B, cl, H, W)))
    if m.numel() != H * W:
        raise ValueError("mask has %d entries, feature map has %d positions" % (m.numel(), H * W))
    dev = x.device
    scratch = torch.zeros(1025, dtype=torch.float32, device=dev)      # 1024 partials + the ticket word
    loss = torch.empty((), dtype=torch.float32, device=dev)
    _lib.call("innercos_loss_fwd", x.data_ptr(), m.data_ptr(), t.data_ptr(), B, Ct, cl, H * W, float(strength),
              0 if crit == "MSE" else 1, scratch.data_ptr(), scratch.data_ptr() + 4096, loss.data_ptr(), _stream_ptr(dev))
    return loss


def innercos_loss_grad(x, mask_f32, target, grad_loss, strength: float, crit: str, c_limit: Optional[int] = None):
    x = _require_cuda(x, "in_data", torch.float32)
    t = _require_cuda(target, "target", torch.float32)
    m = _require_cuda(mask_f32, "mask", torch.float32)
    gl = _require_cuda(grad_loss, "grad_loss", torch.float32).reshape(1)
    B, Ct, H, W = x.shape
    cl = Ct if c_limit is None else min(int(c_limit), Ct)
    gx = torch.empty_like(x)
    _lib.call("innercos_loss_bwd", x.data_ptr(), m.data_ptr(), t.data_ptr(), gl.data_ptr(), B, Ct, cl, H * W,
              float(strength), 0 if crit == "MSE" else 1, gx.data_ptr(), _stream_ptr(x.device))
    return gx


# --------------------------------------------------------------------------------------
# stand-alone pieces used by the API-compatibility modules and the tests
# --------------------------------------------------------------------------------------
def extract_normalize(x: torch.Tensor):
    """util/NonparametricShift.py:36-40,59-73 for one batch: returns (xt [B,N,C], inv_norm [B,N])."""
    x = _require_cuda(x, "target_img", torch.float32)
    B, Cc, H, W = x.shape
    N = H * W
    xt = torch.empty((B, N, Cc), dtype=torch.float32, device=x.device)
    inv = torch.empty((B, N), dtype=torch.float32, device=x.device)
    _lib.call("ipsr_extract_normalize", x.data_ptr(), x.data_ptr(), B, Cc, N, None, 0, inv.data_ptr(), None,
              xt.data_ptr(), None, None, None, None, None, None, None, None, _stream_ptr(x.device))
    return xt, inv


def maxcoord(scores: torch.Tensor):
    """util/MaxCoord.py:22 on a materialised [P, L] score tensor: (ind int64 [L], vmax [L])."""
    s = _require_cuda(scores, "input", torch.float32)
    P, L = s.shape
    ind = torch.empty(L, dtype=torch.int64, device=s.device)
    vmax = torch.empty(L, dtype=torch.float32, device=s.device)
    _lib.call("ipsr_maxcoord", s.data_ptr(), P, L, ind.data_ptr(), vmax.data_ptr(), _stream_ptr(s.device))
    return ind, vmax
