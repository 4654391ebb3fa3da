"""clock64 phase trace of the forward recurrence (steps 8..15 of the first segment; CTA 0 and CTA 100, chain 0, thread 0)."""
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
names = {0: "loop top", 1: "after wait+barrier", 2: "after FMAs", 3: "after refill issue", 4: "after reduce+shuffle", 5: "after gates+stores"}
for base, cta in ((0, 0), (512, 100)):
    n = buf[base + 510]
    print(f"CTA {cta}: R={buf[base + 508]} nrows={buf[base + 509]}  ({n} stamps; first segment + continued segment share the buffer)")
    prev = None
    for i in range(min(n, 60)):
        tag, t = buf[base + 2 * i], buf[base + 2 * i + 1]
        d = "" if prev is None else f"{t - prev:6d}"
        print(f"   step {tag // 100:3d}  {names[tag % 100]:24s} {d}")
        prev = t
