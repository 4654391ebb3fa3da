"""Warm-cache timing of single library GEMM launches inside a CUDA graph (no ncu cache flush)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from marl_b200 import _lib as L

def bench(name, fn, iters=50):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    print(f"{name:44s} {a.elapsed_time(b)/iters*1e3:8.2f} us/launch")

dev = "cuda"
# QMIX hyper GEMM: [3840,120] x [256,120]^T
M, N, S = 3840, 5, 120
wcat = torch.randn(256, S, device=dev); bcat = torch.randn(256, device=dev); wb2 = torch.randn(32, device=dev); bb2 = torch.randn(1, device=dev)
s = torch.randn(M, S, device=dev); hy = torch.empty(M, 256, device=dev)
p = L.QmixParams(wcat.data_ptr(), bcat.data_ptr(), wb2.data_ptr(), bb2.data_ptr())
bench("hyper fwd [3840x120]x[256x120]", lambda: L.call("marl_qmix_hyper_fwd", M, N, S, C.byref(p), s.data_ptr(), hy.data_ptr(), L.stream_ptr()))
dhy = torch.randn(M, 256, device=dev); gw = torch.zeros(256, S, device=dev); gb = torch.zeros(256, device=dev)
g = L.QmixGrads(gw.data_ptr(), gb.data_ptr(), wb2.data_ptr(), bb2.data_ptr())
bench("hyper wgrad [256x3840]x[3840x120]", lambda: L.call("marl_qmix_hyper_wgrad", M, N, S, s.data_ptr(), dhy.data_ptr(), C.byref(g), L.stream_ptr()))
# empty-ish kernel for launch overhead reference
x = torch.zeros(1024, device=dev)
bench("torch x.add_(1) [1024] (launch floor)", lambda: x.add_(1))
# big GEMM (cfg3-like): [19200,216] x [2048,216]^T
M2, S2, N2 = 19200, 216, 2048
w2 = torch.randn(N2, S2, device=dev); b2 = torch.randn(N2, device=dev); s2 = torch.randn(M2, S2, device=dev); y2 = torch.empty(M2, N2, device=dev)
# reuse the qmix hyper entry point: N agents such that N*32+96 == 2048 -> N = 61
p2 = L.QmixParams(w2.data_ptr(), b2.data_ptr(), wb2.data_ptr(), bb2.data_ptr())
bench("fwd [19200x216]x[2048x216] (17 GFLOP)", lambda: L.call("marl_qmix_hyper_fwd", M2, 61, S2, C.byref(p2), s2.data_ptr(), y2.data_ptr(), L.stream_ptr()), iters=10)
ref = s2 @ w2.t() + b2
print("max rel err vs torch fp32 matmul:", float((y2 - ref).abs().max() / ref.abs().max()))
