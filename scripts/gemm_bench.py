"""Micro-benchmark of the tcgen05 node GEMM (fp32-out test entry point) for the shapes the forward uses."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from diffndm_b200 import engine as E
dev = torch.device('cuda')
for (M, N, K) in [(22434, 512, 256), (22434, 1024, 256), (22434, 256, 512), (22434, 256, 256), (2334, 512, 256), (75000, 512, 256)]:
    a = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(N, K, device=dev).bfloat16(); b = torch.randn(N, device=dev)
    for _ in range(3): E.test_gemm(a, w, b, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): E.test_gemm(a, w, b, 0)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    fl = 2.0 * M * N * K
    by = M * K * 2 + N * K * 2 + M * N * 4
    ref_us = None
    af, wf = a, w
    for _ in range(3): torch.matmul(af, wf.T)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): torch.matmul(af, wf.T)
    e1.record(); torch.cuda.synchronize(); ref_us = e0.elapsed_time(e1) / 20 * 1e3
    print(f'M={M} N={N} K={K}: {us:7.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  {by / us / 1e3:7.1f} GB/s (min bytes)   cuBLAS bf16-out {ref_us:7.1f} us')
