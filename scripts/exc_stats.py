"""Exception-list statistics of the synthetic bench data for a range of seeds (diagnostic)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepinpainting_b200 import shift_ops
B, C, H = 64, 256, 64
N = H * H
flag = np.zeros((H, H), np.int64); flag[H // 4:3 * H // 4, H // 4:3 * H // 4] = 1
mi = shift_ops.mask_index_from_flag(torch.from_numpy(flag.reshape(-1)), "cuda")
for seed in range(1234, 1242):
    gen = torch.Generator(device="cpu").manual_seed(seed)
    tot = []
    for s in range(2):
        x = torch.randn(B, C, H, H, generator=gen).cuda()
        ref = (torch.relu(torch.randn(B, C, H, H, generator=gen)) * 3).cuda()
        g = torch.randn(B, C, H, H, generator=gen)
        out, sv = shift_ops.shift_forward(x, ref, mi, need_grad=True)
        torch.cuda.synchronize()
        tot.append(sv.exc_total.cpu().numpy())
    t = np.concatenate(tot)
    print("seed", seed, "exc_total: mean %.0f max %d  cap %d  overflowed images %d  wn/wo finite %s" % (
        t.mean(), t.max(), sv.exc_cap, int((t > sv.exc_cap).sum()), bool(torch.isfinite(sv.wn).all() and torch.isfinite(sv.wo).all())))
