"""Graph-replay step time of every BASELINE config (device-resident batch)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from marl_b200.synthetic import CONFIGS, synthetic_batch
from marl_b200.common.arguments import default_args
from marl_b200.controller.share_params import SharedMAC
from marl_b200.algorithm.q_learner import QLearner
from marl_b200.algorithm.qtran_learner import QTRANLearner

names = sys.argv[1:] or ["matrix_game", "2s3z:vdn", "2s3z", "3s5z", "27m_vs_30m", "matrix_game_4096"]
for spec in names:
    spec, _, bsz = spec.partition("@")
    name, _, alg = spec.partition(":")
    c = dict(CONFIGS[name]); alg = alg or c["alg"]
    if bsz:
        c["B"] = int(bsz)
    args = default_args(alg=alg, n_agents=c["N"], n_actions=c["A"], obs_shape=c["O"], state_shape=c["S"], episode_limit=c["T"], map=name)
    torch.manual_seed(0)
    mac = SharedMAC(args)
    learner = QTRANLearner(mac, args) if alg == "qtran_base" else QLearner(mac, args)
    hb = synthetic_batch(0, c["B"], c["T"], c["N"], c["A"], c["O"], c["S"])
    db = {k: torch.as_tensor(v, device="cuda") for k, v in hb.items()}
    db = {k: (v.to(torch.int64) if k == "u" else v.to(torch.float32)).contiguous() for k, v in db.items()}
    db["max_episode_len"] = c["T"]
    for i in range(4): loss = learner.train(db, i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 20
    a.record()
    for i in range(K): learner.train(db, 4 + i)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / K
    print(f"{name:18s} {alg:11s} B={c['B']:5d} T={c['T']:4d} N={c['N']:3d}: {ms*1e3:10.1f} us/step  {c['B']/ms*1e3:12.0f} episode-samples/s  loss {loss:.5f}  mem {torch.cuda.max_memory_allocated()/2**30:.2f} GiB")
