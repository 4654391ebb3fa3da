"""Accuracy + warm-cache timing of the dense primitives (marl_linear_fwd / dgrad / wgrad) against torch fp64."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from marl_b200 import _lib as L

dev = "cuda"
torch.manual_seed(0)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))


def check(M, N, K, ldx=None, relu=1):
    ldx = ldx or K
    xs = torch.randn(M, ldx, device=dev)
    x = xs[:, :K]
    w = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev)
    y = torch.empty(M, N, device=dev)
    f = lambda: L.call("marl_linear_fwd", xs.data_ptr(), ldx, w.data_ptr(), K, b.data_ptr(), y.data_ptr(), N, M, N, K, relu, L.stream_ptr())
    f(); torch.cuda.synchronize()
    ref = x.double() @ w.double().t() + b.double()
    if relu:
        ref = ref.clamp_min(0)
    e_f = rel(y, ref); t_f = timed(f)
    # dgrad: dx[M,K] = (dy[M,N] . w[N,K]) * (x > 0)
    dy = torch.randn(M, N, device=dev)
    dx = torch.empty(M, K, device=dev)
    g = lambda: L.call("marl_linear_dgrad", dy.data_ptr(), N, w.data_ptr(), K, xs.data_ptr(), ldx, dx.data_ptr(), K, M, N, K, L.stream_ptr())
    g(); torch.cuda.synchronize()
    refd = (dy.double() @ w.double()) * (x > 0)
    e_d = rel(dx, refd); t_d = timed(g)
    # wgrad
    dw = torch.zeros(N, K, device=dev); db = torch.zeros(N, device=dev)
    h = lambda: L.call("marl_linear_wgrad", dy.data_ptr(), N, xs.data_ptr(), ldx, dw.data_ptr(), K, db.data_ptr(), M, N, K, L.stream_ptr())
    h(); torch.cuda.synchronize()
    refw = dy.double().t() @ x.double(); refb = dy.double().sum(0)
    e_w = rel(dw, refw); e_b = rel(db, refb)
    dw2 = torch.zeros(N, K, device=dev); db2 = torch.zeros(N, device=dev)
    L.call("marl_linear_wgrad", dy.data_ptr(), N, xs.data_ptr(), ldx, dw2.data_ptr(), K, db2.data_ptr(), M, N, K, L.stream_ptr())
    torch.cuda.synchronize()
    det = bool(torch.equal(dw, dw2) and torch.equal(db, db2))
    t_w = timed(h)
    fl = 2.0 * M * N * K
    print(f"M={M:6d} N={N:4d} K={K:4d} ld={ldx:4d} | fwd {e_f:.1e} {t_f:7.1f}us {fl/t_f/1e6:6.1f}TF | dgrad {e_d:.1e} {t_d:7.1f}us | "
          f"wgrad {e_w:.1e} db {e_b:.1e} {t_w:7.1f}us bitwise-repeatable={det}", flush=True)
    return max(e_f, e_d, e_w, e_b)


if __name__ == "__main__":
    worst = 0.0
    for (M, N, K, ld) in [(19200, 64, 96, 96), (19200, 192, 64, 64), (19200, 64, 192, 192), (3840, 256, 120, 120), (19200, 64, 80, 80),
                          (18, 64, 8, 8), (1000, 40, 36, 36), (153600, 64, 152, 152), (19200, 2048, 216, 216), (19200, 256, 328, 328),
                          (5760, 64, 1232, 1232), (300, 130, 100, 104)]:
        worst = max(worst, check(M, N, K, ld))
    print("worst rel err", worst)
