"""globaltimer phase trace of one warp of the fused QMIX mixing kernel (needs a -DMARL_MIX_TRACE build: tools/ab_build.py mixtrace -DMARL_MIX_TRACE)."""
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
names = {31: "  hidden rows staged", 32: "  heads computed (q, qn, qt written)", 0: "kernel entry", 1: "after pdl wait", 2: "heads staged (+ barrier)", 3: "selection done (hidden loads, heads, scan)", 4: "eval forward",
         5: "target forward + TD", 6: "dhy stores + dq", 7: "dq_dense / dhext", 8: "block reduction, end"}
for base, blk in ((0, 0), (512, 300)):
    n = buf[base + 510]
    print(f"block {blk}: {n} stamps")
    prev = None
    for i in range(n):
        tag, t = buf[base + 2 * i], buf[base + 2 * i + 1]
        print(f"   {str(names.get(tag, tag)):45s} {'' if prev is None else f'+{t - prev:6d} ns'}")
        prev = t
