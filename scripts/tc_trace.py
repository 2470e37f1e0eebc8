"""In-kernel timeline of the tensor-core kernels (build with MARLC_NVCC_EXTRA=-DMARLC_TC_TRACE):
prints, for CTA (0,0,0) of a few representative launches, the cycle at which each pipeline event happens."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as ct
import torch
from marlclassification_b200 import _lib

L = _lib.lib()
dev = "cuda"
M = 128
for x3 in (0, 1):
    for (N, K) in ((384, 256), (368, 2048)):
        A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev); C = torch.empty(M, N, device=dev)
        for _ in range(3):
            L.marlc_tc_gemm(A.data_ptr(), K, 0, B.data_ptr(), K, 0, None, 0, None, 0, 0, None, C.data_ptr(), N, M, N, K, 0, 0, x3, _lib.stream_ptr())
            torch.cuda.synchronize()
    # LSTM pair at the C2 shape
    Kin, n = 368, 256
    u = torch.randn(M, Kin, device=dev)
    hs = [torch.randn(M, n, device=dev) for _ in range(2)]
    cs = [torch.randn(M, n, device=dev) for _ in range(2)]
    wih = [torch.randn(4 * n, Kin, device=dev) * 0.05 for _ in range(2)]
    whh = [torch.randn(4 * n, n, device=dev) * 0.05 for _ in range(2)]
    bih = [torch.zeros(4 * n, device=dev) for _ in range(2)]
    bhh = [torch.zeros(4 * n, device=dev) for _ in range(2)]
    cn = [torch.empty(M, n, device=dev) for _ in range(2)]
    hn = [torch.empty(M, n, device=dev) for _ in range(2)]
    gt = [torch.empty(M, 4 * n, device=dev) for _ in range(2)]
    arr = lambda ts: (ct.c_void_p * 2)(*[t.data_ptr() for t in ts])
    for _ in range(3):
        _lib.check(L.marlc_tc_lstm_pair(u.data_ptr(), M, Kin, n, arr(hs), arr(cs), arr(wih), arr(whh), arr(bih), arr(bhh),
                                        arr(cn), arr(hn), arr(gt), x3, _lib.stream_ptr()))
        torch.cuda.synchronize()
