import os, sys, ctypes as C
sys.path.insert(0, "/root/repo")
import torch
from marl_b200 import _lib as L
dev = "cuda"
M, N, S = int(sys.argv[1]) if len(sys.argv) > 1 else 19200, 5, 120
wcat = torch.randn(256, S, device=dev); bcat = torch.randn(256, device=dev); wb2 = torch.randn(32, device=dev); bb2 = torch.randn(1, device=dev)
s = torch.randn(M, S, device=dev); hy = torch.empty(M, 256, device=dev)
p = L.QmixParams(wcat.data_ptr(), bcat.data_ptr(), wb2.data_ptr(), bb2.data_ptr())
for _ in range(5):
    L.call("marl_qmix_hyper_fwd", M, N, S, C.byref(p), s.data_ptr(), hy.data_ptr(), L.stream_ptr())
torch.cuda.synchronize()
lib = L.load(); lib.marl_gemm_trace.argtypes = [C.c_void_p]; lib.marl_gemm_trace.restype = C.c_int
buf = (C.c_longlong * 128)(); lib.marl_gemm_trace(C.cast(buf, C.c_void_p))
t = list(buf); t0 = t[0]
print("entry 0 | setup+wait", t[1]-t0, "| before fetch", t[5]-t0, "| after fetch issue", t[6]-t0, "| epilogue done", t[3]-t0, "| teardown", t[4]-t0)
for kt in range(8):
    b = 8 + 6*kt
    if not t[b]: break
    r = [x - t0 for x in t[b:b+6]]
    print(f"kt {kt}: top {r[0]:6d} stage-free +{r[1]-r[0]:5d} split+store +{r[2]-r[1]:5d} fence +{r[3]-r[2]:5d} sync +{r[4]-r[3]:5d} mma-issue +{r[5]-r[4]:5d}")
