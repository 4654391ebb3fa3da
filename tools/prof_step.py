"""A few eager QMIX train steps at the 2s3z shape -- the command profiled by ncu (no CPU legs)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from bench import make_args, SHAPE
from marl_b200.algorithm.q_learner import QLearner
from marl_b200.controller.share_params import SharedMAC
from marl_b200.synthetic import synthetic_batch

alg = sys.argv[1] if len(sys.argv) > 1 else "qmix"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
args = make_args(alg)
args.cuda_graph = False
torch.manual_seed(0)
learner = QLearner(SharedMAC(args), args)
hb = synthetic_batch(0, **SHAPE)
db = {k: torch.as_tensor(v, device="cuda") for k, v in hb.items()}
db = {k: (v.to(torch.int64) if k == "u" else v.to(torch.float32)).contiguous() for k, v in db.items()}
db["max_episode_len"] = SHAPE["T"]
for i in range(steps):
    print(i, learner.train(db, i))
