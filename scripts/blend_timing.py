"""Time ipsr_blend_stage and ipsr_blend_scan alone (CUDA events): python scripts/blend_timing.py B C H"""
import sys, os, ctypes as C_, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepinpainting_b200 import _lib
B, C, H = (int(v) for v in sys.argv[1:4])
N = H * H
dev = torch.device("cuda")
m = torch.zeros(H, H, dtype=torch.int32); m[H // 4:3 * H // 4, H // 4:3 * H // 4] = 1
mask_idx = m.reshape(-1).nonzero().reshape(-1).int().to(dev)
M = mask_idx.numel()
gen = torch.Generator().manual_seed(7)
xt = torch.rand(B, N, C, generator=gen).to(dev)
r_masked = torch.rand(B, M, C, generator=gen).to(dev)
inv = (1.0 / (xt.norm(dim=2) + 1e-8)).contiguous()
ind = torch.randint(0, N, (B, N), generator=gen, dtype=torch.int32).to(dev)
T = _lib.call_value("ipsr_scan_block_steps", C) if hasattr(_lib, "call_value") else _lib.load().ipsr_scan_block_steps(C)
bf = _lib.load().ipsr_staged_block_floats(C)
nb = (M + T - 1) // T
staged = torch.zeros(B * (nb + 1) * bf, device=dev)
Mp = _lib.load().ipsr_padded_steps(M)
y = torch.zeros(B, C, Mp, device=dev); wn = torch.zeros(B, M, device=dev); wo = torch.zeros(B, M, device=dev)
st = torch.cuda.current_stream().cuda_stream
p = lambda t: C_.c_void_p(t.data_ptr())
def stage(): _lib.call("ipsr_blend_stage", p(xt), p(r_masked), p(inv), p(ind), p(mask_idx), B, C, N, M, p(staged), None, C_.c_void_p(st))
def scan(): _lib.call("ipsr_blend_scan", p(staged), B, C, M, p(y), p(wn), p(wo), C_.c_void_p(st))
def timeit(f, n=50):
    for _ in range(5): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
print("B=%d C=%d H=%d M=%d: stage %.1f us  scan %.1f us   (y checksum %.6e, wn %.6e)" % (
    B, C, H, M, timeit(stage), timeit(scan), y.double().sum().item(), wn.double().sum().item()))

