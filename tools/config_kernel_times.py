"""Per-kernel device time (library profiler, eager pass) of one train step of a BASELINE config:  config_kernel_times.py 3s5z [alg]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from marl_b200 import _lib as L
from marl_b200.synthetic import CONFIGS, synthetic_batch
from marl_b200.common.arguments import default_args
from marl_b200.controller.share_params import SharedMAC
from marl_b200.algorithm.q_learner import QLearner
from marl_b200.algorithm.qtran_learner import QTRANLearner

name = sys.argv[1] if len(sys.argv) > 1 else "3s5z"
c = dict(CONFIGS[name]); alg = sys.argv[2] if len(sys.argv) > 2 else c["alg"]
args = default_args(alg=alg, n_agents=c["N"], n_actions=c["A"], obs_shape=c["O"], state_shape=c["S"], episode_limit=c["T"], map=name)
args.cuda_graph = False
args.early_exit = os.environ.get('MARL_EARLY', '0') == '1'
torch.manual_seed(0)
mac = SharedMAC(args)
learner = QTRANLearner(mac, args) if alg == "qtran_base" else QLearner(mac, args)
hb = synthetic_batch(0, c["B"], c["T"], c["N"], c["A"], c["O"], c["S"])
db = {k: torch.as_tensor(v, device="cuda") for k, v in hb.items()}
db = {k: (v.to(torch.int64) if k == "u" else v.to(torch.float32)).contiguous() for k, v in db.items()}
db["max_episode_len"] = c["T"]
for i in range(3): learner.train(db, i)
torch.cuda.synchronize()
L.profile(True)
P = 3
for i in range(P): learner.train(db, 10 + i)
rows = L.profile_timeline()
prof = L.profile_collect()
L.profile(False)
tot = sum(ms for _, ms in prof.values())
for k, (cnt, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:28s} {cnt/P:5.1f} launches  {ms/P*1e3:9.1f} us/step  {100*ms/tot:5.1f}%")
print(f"sum {tot/P*1e3:.1f} us/step")
# the longest launches of the last step
last = rows[-(len(rows) // P):] if rows else []
t0 = last[0][1] if last else 0
full = os.environ.get("MARL_FULL_TIMELINE", "0") == "1"      # every launch of the step in start order instead of the 14 longest
for n, s, e in (sorted(last, key=lambda r: r[1]) if full else sorted(last, key=lambda r: -(r[2]-r[1]))[:14]):
    print(f"   {s - t0:9.1f} {e - t0:9.1f} {e - s:8.1f}  {n}")
