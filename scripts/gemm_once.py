"""One launch each of the batched weight-gradient GEMM (MN-major x MN-major, R = 65 536 rows) and of an input-gradient
GEMM (K-major x MN-major, M = 4096 rows) with 3xTF32, for ncu captures:
    ncu --set full --clock-control none -k regex:tc_gemm_kernel -s <warm-up launches> -c 2 python scripts/gemm_once.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from marlclassification_b200 import _lib

dev = torch.device("cuda", 0)
L = _lib.lib()
g = torch.Generator(device=dev).manual_seed(0)
R, N, K = 65536, 1024, 368
dY, X = torch.randn(R, N, device=dev, generator=g), torch.randn(R, K, device=dev, generator=g)
dW = torch.zeros(N, K, device=dev)
M2, N2, K2 = 4096, 368, 1024
A2, B2 = torch.randn(M2, K2, device=dev, generator=g), torch.randn(K2, N2, device=dev, generator=g)
C2 = torch.zeros(M2, N2, device=dev)
for _ in range(3):  # launches 0..5: warm-up; 4 and 5 are the ones to capture (-s 4 -c 2)
    _lib.check(L.marlc_tc_gemm(dY.data_ptr(), N, 1, X.data_ptr(), K, 1, None, 0, None, 0, 0, None, dW.data_ptr(), K, N, K, R,
                               1, 1, 1, _lib.stream_ptr(dev)))
    _lib.check(L.marlc_tc_gemm(A2.data_ptr(), K2, 0, B2.data_ptr(), N2, 1, None, 0, None, 0, 0, None, C2.data_ptr(), N2, M2,
                               N2, K2, 0, 0, 1, _lib.stream_ptr(dev)))
torch.cuda.synchronize()
print("dW flops", 2.0 * R * N * K, "dX flops", 2.0 * M2 * N2 * K2)
