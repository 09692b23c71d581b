"""Time the backward alone (CUDA events) on bench-like data: python scripts/bwd_timing.py B C H"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from deepinpainting_b200 import shift_ops
B, C, H = (int(v) for v in sys.argv[1:4])
N = H * H
dev = torch.device("cuda")
f = np.zeros((H, H), np.int64); f[H // 4:3 * H // 4, H // 4:3 * H // 4] = 1
mi = shift_ops.mask_index_from_flag(torch.from_numpy(f.reshape(-1)), dev)
gen = torch.Generator().manual_seed(1234)
sets = []
for _ in range(4):
    x = torch.randn(B, C, H, H, generator=gen).to(dev); r = (torch.relu(torch.randn(B, C, H, H, generator=gen)) * 3).to(dev)
    g = torch.randn(B, C, H, H, generator=gen).to(dev)
    out, sv = shift_ops.shift_forward(x, r, mi, need_grad=True)
    sets.append((g, sv))
torch.cuda.synchronize()
def run(i):
    g, sv = sets[i % 4]
    return shift_ops.shift_backward(g, sv, 1.0)
for i in range(5): run(i)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(40): run(i)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 40
tot = sets[0][1].exc_total.cpu()
print("B=%d C=%d H=%d dbg=%s tile=%s: bwd %.1f us  %.0f GB/s   exc per image mean %.0f max %d" % (B, C, H, os.environ.get("IPSR_BWD_DBG"), os.environ.get("IPSR_BWD_TILE_KB"), ms * 1e3, 2 * B * C * N * 4 / ms / 1e6, tot.float().mean(), tot.max()))
