"""Measure the accumulation error of the tcgen05 correlation kernel alone: compare its dumped score matrix
with the fp64 value of exactly the operand split it multiplies (rounding of the operands excluded)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepinpainting_b200 import _lib
DEV = "cuda:0"
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
for (C, H, B) in ((256, 32, 2), (512, 32, 1), (256, 64, 1)):
    rng = np.random.default_rng(C + H)
    N = H * H
    x = rng.standard_normal((B, C, H, H)).astype(np.float32)
    ref = (np.maximum(rng.standard_normal((B, C, H, H)), 0) * 3).astype(np.float32)
    xd, rd = torch.from_numpy(x).to(DEV), torch.from_numpy(ref).to(DEV)
    f32 = dict(dtype=torch.float32, device=DEV)
    inv, rnorm = torch.empty(B, N, **f32), torch.empty(B, N, **f32)
    xt = torch.empty(B, N, C, **f32)
    xtiles = torch.zeros(B * C * N * 4, dtype=torch.uint8, device=DEV)
    rtiles = torch.zeros(B * C * N * 4, dtype=torch.uint8, device=DEV)
    nonfinite = torch.zeros(B, dtype=torch.int32, device=DEV)
    _lib.call("ipsr_extract_normalize", xd.data_ptr(), rd.data_ptr(), B, C, N, None, 0, inv.data_ptr(), rnorm.data_ptr(),
              xt.data_ptr(), None, xtiles.data_ptr(), rtiles.data_ptr(), nonfinite.data_ptr(), st)
    pb = torch.empty(1, B, N, **f32); ps = torch.empty(1, B, N, **f32)
    pi = torch.empty(1, B, N, dtype=torch.int32, device=DEV)
    dump = torch.full((B, N, N), float("nan"), **f32)
    _lib.call("ipsr_correlate_argmax_tc", rtiles.data_ptr(), xtiles.data_ptr(), B, C, N, 0, N, 1, pb.data_ptr(),
              pi.data_ptr(), ps.data_ptr(), dump.data_ptr(), st)
    torch.cuda.synchronize()
    S = dump.double()
    Xn = (xt * inv.unsqueeze(-1))                      # fl(X*inv) fp32, as prep computes it
    R = rd.view(B, C, N).transpose(1, 2).contiguous()
    def split(t):
        hi = t.to(torch.bfloat16).float()
        lo = (t - hi).to(torch.bfloat16).float()
        return hi.double(), lo.double()
    Xh, Xl = split(Xn); Rh, Rl = split(R)
    S3 = Rh @ Xh.transpose(1, 2) + Rh @ Xl.transpose(1, 2) + Rl @ Xh.transpose(1, 2)
    S64 = R.double() @ Xn.double().transpose(1, 2)
    rn = R.double().norm(dim=2, keepdim=True)
    acc_err = ((S - S3).abs() / rn).max().item()
    tot_err = ((S - S64).abs() / rn).max().item()
    split_err = ((S3 - S64).abs() / rn).max().item()
    print("C=%d N=%d: accumulation err/|R| max %.3e   operand-split err %.3e   total %.3e" % (C, N, acc_err, split_err, tot_err))
