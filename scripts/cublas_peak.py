"""Calibration for the tensor-pipe counter: the cuBLAS bf16 8192^3 GEMM that MEASURED_PEAKS.json's bf16 figure comes from.
Run under `ncu --set full -k regex:gemm|cutlass|nvjet -c 2` to read sm__pipe_tensor_cycles_active for the kernel that DEFINES
the peak, next to the same counter of corr_tc_kernel (north_star: tensor-pipe utilisation >= 60 %)."""
import torch

n = 8192
a = torch.randn(n, n, device="cuda", dtype=torch.bfloat16)
b = torch.randn(n, n, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    c = a @ b
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    c = a @ b
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("cuBLAS bf16 %d^3: %.3f ms  %.1f TFLOP/s" % (n, ms, 2 * n ** 3 / ms / 1e9))
