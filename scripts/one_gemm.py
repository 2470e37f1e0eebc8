import os, sys
sys.path.insert(0, "/root/repo")
import torch
from marlclassification_b200 import _lib
L = _lib.lib()
M, N, K = 128, 384, 256
A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.empty(M, N, device="cuda")
for _ in range(6):
    L.marlc_tc_gemm(A.data_ptr(), K, 0, B.data_ptr(), K, 0, None, 0, None, 0, 0, None, C.data_ptr(), N, M, N, K, 0, 0, 0, _lib.stream_ptr())
torch.cuda.synchronize()
