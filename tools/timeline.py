"""Multi-stream timeline of ONE eager train step (library profiler events; the host is run ahead with a spin
kernel so the events bracket device time only).  usage: timeline.py [alg] [B]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from bench import make_args, SHAPE
from marl_b200 import _lib as L
from marl_b200.algorithm.q_learner import QLearner
from marl_b200.controller.share_params import SharedMAC
from marl_b200.synthetic import synthetic_batch

alg = sys.argv[1] if len(sys.argv) > 1 else "qmix"
args = make_args(alg)
torch.manual_seed(0)
learner = QLearner(SharedMAC(args), args)
SHAPE = dict(SHAPE); SHAPE["B"] = int(sys.argv[2]) if len(sys.argv) > 2 else SHAPE["B"]
hb = synthetic_batch(0, **SHAPE)
db = {k: torch.as_tensor(v, device="cuda") for k, v in hb.items()}
db = {k: (v.to(torch.int64) if k == "u" else v.to(torch.float32)).contiguous() for k, v in db.items()}
db["max_episode_len"] = SHAPE["T"]
learner._use_graph = False
for i in range(5):
    learner.train(db, i)
torch.cuda.synchronize()
L.profile(True)
L.call("marl_spin_us", 3000, L.stream_ptr())
learner.train(db, 10)
rows = L.profile_timeline()
L.profile_collect()
L.profile(False)
t0 = rows[0][1]
for name, s, e in sorted(rows, key=lambda r: r[1]):
    print(f"{s - t0:9.1f} {e - t0:9.1f} {e - s:8.1f}  {name}")
