"""GEMM shape sweep through the C ABI (CUDA events, L2-warm back-to-back launches, median)."""
import os, sys, statistics, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lw-detr_b200"))
from b200 import capi
shapes = [("qkv", 51200, 576, 192, 0), ("fc1", 51200, 768, 192, 2), ("fc2", 51200, 192, 768, 0), ("proj", 51200, 192, 192, 0),
          ("med_qkv", 102400, 1152, 384, 0), ("xl_fc1", 25600, 3072, 768, 2), ("dec_l1", 9600, 2048, 256, 1)]
for name, M, N, K, act in shapes:
    A = torch.randn(M, K, device="cuda").half(); W = torch.randn(N, K, device="cuda").half() * 0.05
    b = torch.randn(N, device="cuda"); out = torch.empty(M, N, device="cuda", dtype=torch.half)
    for _ in range(3): capi.gemm(A, W, out, bias=b, act=act)
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); capi.gemm(A, W, out, bias=b, act=act); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    us = statistics.median(ts)
    print("%-8s M=%6d N=%4d K=%4d act=%d  %7.1f us  %6.1f TF/s  %6.0f GB/s" % (name, M, N, K, act, us, 2.0 * M * N * K / us / 1e6, 2.0 * (M * K + N * K + M * N) / us / 1e3))
