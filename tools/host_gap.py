"""Host / synchronisation share of a train step at config 2: wall time per train() call vs back-to-back replays of the same CUDA graph."""
import os, sys, time, cProfile, pstats, io
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from bench import make_args, SHAPE
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
for i in range(10): learner.train(db, i)
torch.cuda.synchronize()
n = 300
t0 = time.perf_counter()
for i in range(n): learner.train(db, 20 + i)
t1 = time.perf_counter()
g = [v for v in learner._graphs.values() if isinstance(v, tuple)][0][0]
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
a.record()
for i in range(n): g.replay()
b.record(); torch.cuda.synchronize()
print(f"train() wall {1e6 * (t1 - t0) / n:.1f} us/step; graph replays back to back {1e3 * a.elapsed_time(b) / n:.1f} us/step")
pr = cProfile.Profile(); pr.enable()
for i in range(n): learner.train(db, 400 + i)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500])
