export MARL_B200_TGEMM=1
for d in 0 1 4 5 7; do echo "== debug=$d"; MARL_TGEMM_DEBUG=$d python - <<'PY'
import sys; sys.path.insert(0,'.')
import torch
from marl_b200 import _lib as L
sys.path.insert(0,'tools')
dev='cuda'
def timed(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/iters*1e3
for (M,N,K) in [(19200,192,64),(19200,64,96)]:
    x=torch.randn(M,K,device=dev); w=torch.randn(N,K,device=dev); b=torch.randn(N,device=dev); y=torch.empty(M,N,device=dev)
    f=lambda: L.call("marl_linear_fwd", x.data_ptr(), K, w.data_ptr(), K, b.data_ptr(), y.data_ptr(), N, M, N, K, 1, L.stream_ptr())
    print(M,N,K, "fwd %.1f us" % timed(f))
PY
done
