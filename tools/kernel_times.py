"""Per-kernel device time of one train step (eager pass, library profiler) + graph-replay step time."""
import os, sys, time
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
hb = synthetic_batch(0, **SHAPE)
db = {k: torch.as_tensor(v, device="cuda") for k, v in hb.items()}
db = {k: (v.to(torch.int64) if k == "u" else v.to(torch.float32)).contiguous() for k, v in db.items()}
db["max_episode_len"] = SHAPE["T"]
for i in range(5):
    loss = learner.train(db, i)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
K = 50
for i in range(K):
    learner.train(db, 5 + i)
b.record(); torch.cuda.synchronize()
print(f"graph replay: {a.elapsed_time(b)/K*1e3:.1f} us/step, loss {loss:.6f}")
learner._use_graph = False
L.profile(True)
P = 10
for i in range(P):
    learner.train(db, 100 + i)
prof = L.profile_collect()
tot = sum(ms for _, ms in prof.values())
for k, (c, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:28s} {c/P:5.1f} launches  {ms/P*1e3:8.1f} us/step  {100*ms/tot:5.1f}%")
print(f"sum {tot/P*1e3:.1f} us/step")
