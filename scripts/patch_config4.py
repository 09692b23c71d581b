"""BASELINE.json configs[3] on one GPU: 64x64x256 features, 3x3 patches (P = 3844 patch positions, rows of K = 2304),
centre hole, forward only.  Prints per-call time (CUDA events); run under ncu for the launch list:

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg4.csv \
        python scripts/patch_config4.py --iters 2
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepinpainting_b200 import shift_ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--channels", type=int, default=256)
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--patch", type=int, default=3)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    dev = "cuda:0"
    gen = torch.Generator(device="cpu").manual_seed(1234)
    B, C, H, k = a.batch, a.channels, a.size, a.patch
    x = torch.randn(B, C, H, H, generator=gen).to(dev)
    ref = (torch.relu(torch.randn(B, C, H, H, generator=gen)) * 3).to(dev)
    feat = torch.zeros(H, H, dtype=torch.uint8, device=dev)
    feat[H // 4:3 * H // 4, H // 4:3 * H // 4] = 1
    mi = shift_ops.build_flags(feat, k, 1, 1)
    shift_ops.shift_forward_patches(x, ref, mi, k, 1)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(a.iters):
        shift_ops.shift_forward_patches(x, ref, mi, k, 1)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / a.iters
    print("patch forward B=%d C=%d %dx%d k=%d: P=%d K=%d M=%d  %.3f ms / call  (%.1f images/s)"
          % (B, C, H, H, k, mi.flag.numel(), C * k * k, mi.M, ms, B / ms * 1e3))


if __name__ == "__main__":
    main()
