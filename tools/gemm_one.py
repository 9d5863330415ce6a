#!/usr/bin/env python
"""Run one GEMM shape a few times (for ncu): python tools/gemm_one.py M N K mode[f16|gelu|resid] P"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paella_b200 import _lib, ops  # noqa: E402

M, N, K = (int(v) for v in sys.argv[1:4])
mode = sys.argv[4] if len(sys.argv) > 4 else "f16"
P = int(sys.argv[5]) if len(sys.argv) > 5 else 64
dev = "cuda"
a = torch.randn(M, K, device=dev).half()
w = (torch.randn(N, K, device=dev) / math.sqrt(K)).half()
bias = torch.randn(N, device=dev)
if mode == "gelu":
    out = torch.empty(M, N, device=dev, dtype=torch.float16)
    sq = torch.zeros(M // P, N, device=dev, dtype=torch.int64)
    run = lambda: ops.gemm_f16(a, w, _lib.EPI_GELU_F16, out, bias=bias, sqsum=sq, rows_per_sample=P)
elif mode == "resid":
    out = torch.randn(M, N, device=dev)
    film = torch.randn(M // P, 2 * N, device=dev) * 0.1
    run = lambda: ops.gemm_f16(a, w, _lib.EPI_RESID_F32, out, bias=bias, resid=out, rows_per_sample=P, film=film)
elif mode == "resid_nofilm":
    out = torch.randn(M, N, device=dev)
    run = lambda: ops.gemm_f16(a, w, _lib.EPI_RESID_F32, out, bias=bias, resid=out)
elif mode == "f32":
    out = torch.empty(M, N, device=dev)
    run = lambda: ops.gemm_f16(a, w, _lib.EPI_F32, out, bias=bias)
else:
    out = torch.empty(M, N, device=dev, dtype=torch.float16)
    run = lambda: ops.gemm_f16(a, w, _lib.EPI_F16, out, bias=bias)
for _ in range(5):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"{M}x{N}x{K} {mode}: {ms*1e3:.1f} us  {2.0*M*N*K/ms/1e9:.0f} TFLOP/s")
