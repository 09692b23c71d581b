"""BASELINE.json configs[4]-like shift layer: batch 32, 32x32x512 features, a different free-form mask per image.
Times fwd+bwd through the module API (CUDA events): one batched call with per-image mask rows vs one call per image."""
import collections
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepinpainting_b200.models import IPSR_model  # noqa: E402

dev = "cuda:0"
B, C, H = 32, 512, 32
S = H * 8
rng = np.random.default_rng(7)
mg = np.zeros((B, 1, S, S), bool)
for b in range(B):
    for _ in range(3):
        y0, x0 = rng.integers(0, S - 64, 2)
        mg[b, 0, y0:y0 + int(rng.integers(24, 96)), x0:x0 + int(rng.integers(24, 96))] = True
Ref = collections.namedtuple("Ref", ["relu4_3"])
gen = torch.Generator().manual_seed(1)
x = torch.randn(B, C, H, H, generator=gen).to(dev)
ref = (torch.relu(torch.randn(B, C, H, H, generator=gen)) * 3).to(dev)
g = torch.randn(B, C, H, H, generator=gen).to(dev)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        fn()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps


batched = IPSR_model(5 / 16.0, 1, 1, 1, 1, 1)
batched.set_mask(torch.from_numpy(mg).to(dev), 3, 5 / 16.0)
batched.set_ref(Ref(ref))


def run_batched():
    xin = x.detach().requires_grad_(True)
    batched(xin).backward(g)


singles = []
for b in range(B):
    m = IPSR_model(5 / 16.0, 1, 1, 1, 1, 1)
    m.set_mask(torch.from_numpy(mg[b:b + 1]).to(dev), 3, 5 / 16.0)
    m.set_ref(Ref(ref[b:b + 1].contiguous()))
    singles.append(m)
xs = [x[b:b + 1].contiguous() for b in range(B)]
gs = [g[b:b + 1].contiguous() for b in range(B)]


def run_singles():
    for b in range(B):
        xin = xs[b].detach().requires_grad_(True)
        singles[b](xin).backward(gs[b])


tb, ts = timed(run_batched), timed(run_singles, 5)
print("masked positions per image: min %d max %d" % (int(batched.flag.sum(1).min()), int(batched.flag.sum(1).max())))
print("batch 32, 32x32x512, per-image masks, fwd+bwd: batched %.3f ms (%.0f images/s); one call per image %.3f ms (%.0f images/s)"
      % (tb, B / tb * 1e3, ts, B / ts * 1e3))
