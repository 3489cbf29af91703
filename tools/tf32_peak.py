"""Measured dense TF32 tensor peak of this B200 (cuBLAS fp32 GEMM with allow_tf32), same method as the driver's
bf16 figure in MEASURED_PEAKS.json: 8192^3, best of 10 (burst) -- the denominator for kind::tf32 rooflines."""
import json
import torch

torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
a = torch.randn(n, n, device="cuda")
b = torch.randn(n, n, device="cuda")
for _ in range(3):
    a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    a @ b
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
ab, bb = a.bfloat16(), b.bfloat16()
for _ in range(3):
    ab @ bb
torch.cuda.synchronize()
best16 = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ab @ bb
    e1.record()
    torch.cuda.synchronize()
    best16 = min(best16, e0.elapsed_time(e1))
print(json.dumps({"tf32_tflops_burst": 2 * n ** 3 / best / 1e9, "bf16_tflops_burst": 2 * n ** 3 / best16 / 1e9,
                  "how": "torch.matmul 8192^3, best of 10, CUDA events"}))
