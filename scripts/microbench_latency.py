"""Calibration: per-node latency of dependent kernels inside a CUDA graph on this GPU
(empty kernel, LN+SiLU rows, tiny tensor-core GEMM, LSTM pair).  Prints one line each."""
import ctypes as ct
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from marlclassification_b200 import _lib

L = _lib.lib()
dev = "cuda"
X3 = int(os.environ.get("X3", "0"))


def timed_graph(fn, n, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (reps * n)


st = torch.zeros(2, dtype=torch.int64, device=dev)
step = torch.zeros(1, dtype=torch.int64, device=dev)
p = torch.zeros(64, device=dev)
print(f"empty-ish kernel (adam on 64 floats = 2 launches): {timed_graph(lambda: L.marlc_adam_step(p.data_ptr(), p.data_ptr(), p.data_ptr(), p.data_ptr(), 64, 1e-3, 0.9, 0.999, 1e-8, 1.0, step.data_ptr(), _lib.stream_ptr()), 200) / 2:.2f} us per launch")

M = 128
for N in (128, 384):
    y = torch.randn(M, N, device=dev); g_ = torch.ones(N, device=dev); b_ = torch.zeros(N, device=dev); o = torch.empty_like(y)
    print(f"ln_silu_fwd [{M}x{N}]: {timed_graph(lambda: L.marlc_ln_silu(y.data_ptr(), g_.data_ptr(), b_.data_ptr(), o.data_ptr(), M, N, _lib.stream_ptr()), 200):.2f} us")
for (N, K) in ((128, 64), (128, 256), (384, 256), (1024, 368)):
    A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev); C = torch.empty(M, N, device=dev)
    f = lambda: L.marlc_tc_gemm(A.data_ptr(), K, 0, B.data_ptr(), K, 0, None, 0, None, 0, 0, None, C.data_ptr(), N, M, N, K, 0, 0, X3, _lib.stream_ptr())
    print(f"tc_gemm NT [{M}x{N}x{K}]: {timed_graph(f, 200):.2f} us")
    f2 = lambda: L.marlc_linear(A.data_ptr(), B.data_ptr(), None, C.data_ptr(), M, N, K, _lib.stream_ptr())
    print(f"simt linear [{M}x{N}x{K}]: {timed_graph(f2, 100):.2f} us")
for (Mr, N, K) in ((2048, 384, 256), (4096, 1024, 624), (65536, 384, 256)):
    A = torch.randn(Mr, K, device=dev); B = torch.randn(N, K, device=dev); C = torch.empty(Mr, N, device=dev)
    f = lambda: L.marlc_tc_gemm(A.data_ptr(), K, 0, B.data_ptr(), K, 0, None, 0, None, 0, 0, None, C.data_ptr(), N, Mr, N, K, 0, 0, X3, _lib.stream_ptr())
    t = timed_graph(f, 20)
    print(f"tc_gemm NT [{Mr}x{N}x{K}]: {t:.2f} us = {2 * Mr * N * K / t / 1e6:.1f} TFLOP/s")
