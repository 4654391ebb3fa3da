"""Prints the parity error budget at a BASELINE shape: CUDA vs fp32 oracle vs fp64 oracle (GPU box)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from marl_b200.synthetic import synthetic_batch
from oracle import marl_oracle as MO
from tests import parity_util as PU

alg = sys.argv[1] if len(sys.argv) > 1 else "qmix"
B, T, N, A, O, S = (int(x) for x in (sys.argv[2:8] if len(sys.argv) > 7 else (32, 120, 5, 11, 80, 120)))
kw = {}
if alg == "qplex" and len(sys.argv) > 8:
    kw = dict(num_kernel=int(sys.argv[8]))
args = PU.make_args(alg, N, A, O, S, T, **kw)
args.cuda_graph = False
learner, st32 = PU.build_pair(args)
st64 = MO.LearnerState(st32.cfg, PU.export_params(learner), dtype=torch.float64)
batch = synthetic_batch(0, B, T, N, A, O, S)
for step in range(2):
    loss = learner.train({k: v.copy() for k, v in batch.items()}, step)
    l32, i32 = MO.train_step(st32, batch, step)
    l64, i64 = MO.train_step(st64, batch, step)
    print(f"step {step}: loss mine {loss:.8f} o32 {l32:.8f} o64 {l64:.8f}  |mine-o64|/o64 {abs(loss-l64)/abs(l64):.2e} |o32-o64|/o64 {abs(l32-l64)/abs(l64):.2e}")
    print(f"   grad_norm mine {learner.last['grad_norm']:.6f} o32 {i32['grad_norm']:.6f} o64 {i64['grad_norm']:.6f}")
    ws = learner.last["ws"]
    if step == 0:
        names = [("q_evals", ws["q"][0]), ("hidden_evals", ws["hidden"][0])]
        if alg != "qtran_base":
            names += [("q_targets", ws["q"][1]), ("q_tot", ws["q_tot"]), ("q_tot_target", ws["q_tot_t"])]
        for key, mine in names:
            print(f"   {key:14s} mine-o32 {PU.rel_err(mine.reshape(-1), i32[key].reshape(-1)):.2e} mine-o64 {PU.rel_err(mine.reshape(-1), i64[key].reshape(-1)):.2e} o32-o64 {PU.rel_err(i32[key].reshape(-1), i64[key].reshape(-1)):.2e}")
        if i32.get("a_star") is not None:
            nb, hard = PU.argmax_mismatches(ws["a_star"], i64["q_evals_next"], i64["a_star"].squeeze(3))
            nb2, hard2 = PU.argmax_mismatches(i32["a_star"].squeeze(3), i64["q_evals_next"], i64["a_star"].squeeze(3))
            print(f"   argmax vs o64: mine {nb} mismatches ({hard} hard); o32 {nb2} ({hard2} hard) of {i64['a_star'].numel()}")
    mods = PU.module_groups(learner)
    for g, k, p64 in st64.flat_params():
        p32 = st32.params[g][k]
        mine = dict(mods[g].named_parameters())[k]
        g64 = i64["clipped_grads"][f"{g}.{k}"]
        if g64 is None:
            continue
        g32 = i32["clipped_grads"][f"{g}.{k}"]
        print(f"   {g+'.'+k:34s} grad: mine-o64 {PU.rel_err(mine.grad, g64):.2e} o32-o64 {PU.rel_err(g32, g64):.2e} | param: mine-o64 {PU.rel_err(mine, p64):.2e} o32-o64 {PU.rel_err(p32, p64):.2e}")
