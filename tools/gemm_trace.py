"""Phase trace (clock64 of CTA 0) of one tgemm launch: where a persistent CTA's time goes."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from marl_b200 import _lib as L

M, N, K = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (19200, 64, 96)))
kind = sys.argv[4] if len(sys.argv) > 4 else "fwd"
dev = "cuda"
x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev)
y = torch.empty(M, N, device=dev); dy = torch.randn(M, N, device=dev); dx = torch.empty(M, K, device=dev)
dw = torch.zeros(N, K, device=dev); db = torch.zeros(N, device=dev)
def run():
    if kind == "fwd":
        L.call("marl_linear_fwd", x.data_ptr(), K, w.data_ptr(), K, b.data_ptr(), y.data_ptr(), N, M, N, K, 1, L.stream_ptr())
    elif kind == "dgrad":
        L.call("marl_linear_dgrad", dy.data_ptr(), N, w.data_ptr(), K, x.data_ptr(), K, dx.data_ptr(), K, M, N, K, L.stream_ptr())
    else:
        L.call("marl_linear_wgrad", dy.data_ptr(), N, x.data_ptr(), K, dw.data_ptr(), K, db.data_ptr(), M, N, K, L.stream_ptr())
for _ in range(3): run()
torch.cuda.synchronize()
L.call("marl_tgemm_trace", 1, None)
run()
buf = (C.c_longlong * 2048)()
L.call("marl_tgemm_trace", 0, C.cast(buf, C.c_void_p))
names = ["tma", "mma", "cvt", "epi"]
ev = []
for r in range(4):
    n = buf[r * 512 + 510]
    for i in range(n):
        ev.append((buf[r * 512 + 2 * i + 1], names[r], buf[r * 512 + 2 * i]))
ev.sort()
t0 = ev[0][0] if ev else 0
print(f"{kind} M={M} N={N} K={K}: {len(ev)} events (cycles since first; tag = ring slot use, +1000 = done)")
for t, r, tag in ev[:120]:
    print(f"{t - t0:8d}  {r}  {tag}")
