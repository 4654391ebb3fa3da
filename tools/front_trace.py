"""Phase trace (globaltimer ns, CTA 0) of the fused input-layer kernel inside one eager train step."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from bench import make_args, SHAPE
from marl_b200 import _lib as L
from marl_b200.algorithm.q_learner import QLearner
from marl_b200.controller.share_params import SharedMAC
from marl_b200.synthetic import synthetic_batch

args = make_args("qmix")
torch.manual_seed(0)
learner = QLearner(SharedMAC(args), args)
hb = synthetic_batch(0, **SHAPE)
db = {k: torch.as_tensor(v, device="cuda") for k, v in hb.items()}
db = {k: (v.to(torch.int64) if k == "u" else v.to(torch.float32)).contiguous() for k, v in db.items()}
db["max_episode_len"] = SHAPE["T"]
learner._use_graph = False
for i in range(4):
    learner.train(db, i)
torch.cuda.synchronize()
L.call("marl_tgemm_trace", 1, None)
learner.train(db, 10)
buf = (C.c_longlong * 2048)()
L.call("marl_tgemm_trace", 0, C.cast(buf, C.c_void_p))
names = ["conv", "mma ", "E   ", "F   "]
ev = []
for r in range(4):
    n = buf[r * 512 + 510]
    for i in range(n):
        ev.append((buf[r * 512 + 2 * i + 1], names[r], buf[r * 512 + 2 * i]))
ev.sort()
t0 = ev[0][0] if ev else 0
print(f"{len(ev)} events (ns since first).  prod: g = stage acquired, 1000+g = stage handed over; mma: g = k-tile issue, 2000+j / 3000+j = GEMM 2 of tile j"
      " begin / issued; E: it = acc1 ready, 1000+it = x in TMEM, 2000+it = x stored; F: it = acc2 ready, 1000+it = gi stored; 9000.. = setup / end")
for t, r, tag in ev[:int(sys.argv[1]) if len(sys.argv) > 1 else 200]:
    print(f"{t - t0:8d}  {r}  {tag}")
