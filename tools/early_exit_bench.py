"""Step time with and without args.early_exit on ragged batches (lengths U{T/2..T}) at several batch sizes; the two settings
alternate (two rounds each) so that clock / box drift shows up as a difference between rounds, not between settings."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from marl_b200.synthetic import synthetic_batch
from marl_b200.common.arguments import default_args
from marl_b200.controller.share_params import SharedMAC
from marl_b200.algorithm.q_learner import QLearner

for B, steps in ((32, 300), (128, 100), (256, 60), (1024, 20)):
    hb = synthetic_batch(0, B, 120, 5, 11, 80, 120)
    db = {k: torch.as_tensor(v, device="cuda") for k, v in hb.items()}
    db = {k: (v.to(torch.int64) if k == "u" else v.to(torch.float32)).contiguous() for k, v in db.items()}
    db["max_episode_len"] = 120
    learners = {}
    for early in (False, True):
        args = default_args(alg="qmix", n_agents=5, n_actions=11, obs_shape=80, state_shape=120, episode_limit=120, map="2s3z", early_exit=early)
        torch.manual_seed(0)
        learners[early] = QLearner(SharedMAC(args), args)
        for i in range(5): loss = learners[early].train(db, i)
    for rnd in range(2):
        for early in (False, True):
            learner = learners[early]
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(steps): loss = learner.train(db, 5 + i)
            b.record(); torch.cuda.synchronize()
            print(f"B={B:5d} round {rnd} early_exit={early!s:5}: {a.elapsed_time(b) / steps * 1e3:9.1f} us/step  loss {loss:.5f}", flush=True)
