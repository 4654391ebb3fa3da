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
t = list(buf); t0 = t[8]
for kt in range(8):
    b = 8 + 4*kt
    if not t[b]: break
    r = [x - t0 for x in t[b:b+4]]
    m = [t[64+2*kt]-t0, t[65+2*kt]-t0]
    print(f"kt {kt}: top {r[0]:6d} stage-free +{r[1]-r[0]:5d} split+store+fence+arrive +{r[2]-r[1]:5d} fetch +{r[3]-r[2]:5d} | mma: synced {m[0]:6d} issue +{m[1]-m[0]:5d}")
