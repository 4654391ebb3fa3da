"""One weight-gradient launch shape for ncu / timing:  wgrad_probe.py M N K [iters]   (dw[N,K] += dy[M,N]^T . x[M,K])"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from marl_b200 import _lib as L

M, N, K = (int(v) for v in sys.argv[1:4])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
torch.manual_seed(0)
x = torch.randn(M, K, device="cuda"); dy = torch.randn(M, N, device="cuda")
dw = torch.zeros(N, K, device="cuda"); db = torch.zeros(N, device="cuda")
f = lambda: L.call("marl_linear_wgrad", dy.data_ptr(), N, x.data_ptr(), K, dw.data_ptr(), K, db.data_ptr(), M, N, K, L.stream_ptr())
for _ in range(iters): f()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): f()
b.record(); torch.cuda.synchronize()
ref = dy.double().t() @ x.double() * (iters + 10)
print(f"M={M} N={N} K={K}: {a.elapsed_time(b) / 10 * 1e3:.1f} us/launch  rel err {float((dw.double() - ref).abs().max() / ref.abs().max()):.2e}")
