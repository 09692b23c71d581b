"""Where does the host time of one fwd+bwd call through the module API go?  (cProfile over enqueue-only steps.)"""
import collections
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepinpainting_b200.models import IPSR_model  # noqa: E402

dev = "cuda:0"
B, C, H = 16, 256, 32
Ref = collections.namedtuple("Ref", ["relu4_3"])
layer = IPSR_model(5 / 16.0, 1, 1, 1, 1, 1)
S = H * 8
mg = torch.zeros(1, 1, S, S, dtype=torch.bool)
mg[:, :, S // 4:3 * S // 4, S // 4:3 * S // 4] = True
layer.set_mask(mg.to(dev), 3, 5 / 16.0)
x = torch.randn(B, C, H, H, device=dev)
ref = torch.relu(torch.randn(B, C, H, H, device=dev)) * 3
g = torch.randn(B, C, H, H, device=dev)
layer.set_ref(Ref(ref))


def step():
    xin = x.detach().requires_grad_(True)
    y = layer(xin)
    y.backward(g)
    return xin.grad


for _ in range(10):
    step()
torch.cuda.synchronize()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
t0 = time.perf_counter()
for _ in range(n):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host enqueue per fwd+bwd: %.1f us" % ((t1 - t0) / n * 1e6))
pr = cProfile.Profile()
pr.enable()
for _ in range(n):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
