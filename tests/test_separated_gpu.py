"""SeparatedMAC (share_params.py:389-610) and the module-surface train step on the GPU."""
import numpy as np
import pytest
import torch

from marl_b200.synthetic import synthetic_batch
from oracle import marl_oracle as MO
from tests import golden_util as GU
from tests import parity_util as PU

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _separated(z, alg):
    from marl_b200.controller.share_params import SeparatedMAC
    N, A, O, S, T = (int(x) for x in z["meta/dims"])
    args = PU.make_args(alg, N, A, O, S, T, reuse_network=False)
    mac = SeparatedMAC(args)
    for n, net in enumerate(mac.agent):
        net.load_state_dict(GU.group(z, f"init/agent.{n}"))
    return args, mac


@pytest.mark.parametrize("alg", ["qmix", "vdn"])
def test_separated_mac_forward_matches_reference_golden(alg):
    """Per-agent networks: get_current_q_values, then get_next_q_values continuing from the carried hidden state and
    returning the per-agent LIST of final hidden states (share_params.py:508-590), against the UNMODIFIED reference."""
    z = GU.load(f"separated_{alg}")
    args, mac = _separated(z, alg)
    batch = GU.batch_of(z)
    B, L_ = batch["o"].shape[0], int(np.asarray(z["step0/q_evals"]).shape[1])
    assert isinstance(mac.agent, list) and len(mac.parameters()) == 8 * args.n_agents
    mac.init_hidden(B)
    with torch.no_grad():
        q, hid = mac.get_current_q_values(batch, L_)
        qn, hlist = mac.get_next_q_values(batch, L_)
    assert PU.rel_err(q, z["step0/q_evals"]) < TOL and PU.rel_err(hid, z["step0/hidden_evals"]) < TOL
    assert PU.rel_err(qn, z["step0/q_evals_next"]) < TOL
    assert isinstance(hlist, list) and len(hlist) == args.n_agents
    assert PU.rel_err(torch.stack(hlist), z["step0/next_hidden_list"]) < TOL
    assert tuple(mac.hidden_states.shape) == (B, args.n_agents, 64)


@pytest.mark.parametrize("alg", ["qmix", "vdn"])
def test_learner_with_separated_mac_vs_oracle(alg):
    """QLearner.train with one network per agent, three steps incl. a target sync, against the oracle restatement (whose
    forward is pinned to the reference; the reference's own train() cannot run with this controller, see the fixture)."""
    from marl_b200.algorithm.q_learner import QLearner
    z = GU.load(f"separated_{alg}")
    args, mac = _separated(z, alg)
    args.target_update_cycle = 2
    learner = QLearner(mac, args)
    learner.mixer.load_state_dict(GU.group(z, "init/mixer"))
    learner._update_targets()
    N, A, O, S, T = (int(x) for x in z["meta/dims"])
    cfg = MO.make_cfg(alg=alg, n_agents=N, n_actions=A, obs_shape=O, state_shape=S, episode_limit=T, separated=True,
                      reuse_network=False, target_update_cycle=2)
    params = {f"agent.{n}": GU.group(z, f"init/agent.{n}") for n in range(N)}
    params["mixer"] = GU.group(z, "init/mixer")
    st = MO.LearnerState(cfg, params)
    batch = GU.batch_of(z)
    for step in range(4):
        loss = learner.train({k: v.copy() for k, v in batch.items()}, step)
        oloss, info = MO.train_step(st, batch, step)
        assert abs(loss - oloss) <= (TOL if step == 0 else 5e-5) * abs(oloss), (step, loss, oloss)
        if step == 0:
            for n, net in enumerate(learner.eval_net.agent):
                for k, p in net.named_parameters():
                    assert PU.rel_err(p.grad, info["clipped_grads"][f"agent.{n}.{k}"]) < 5e-5, (n, k)
    for n, (net, tnet) in enumerate(zip(learner.eval_net.agent, learner.target_net.agent)):
        for k, p in net.named_parameters():
            assert PU.params_close(p, st.params[f"agent.{n}"][k], args.lr, 4), (n, k)
        for k, p in tnet.named_parameters():            # synced at step 2, one more update since
            assert PU.params_close(p, st.target[f"agent.{n}"][k], args.lr, 4), (n, k)


@pytest.mark.parametrize("alg", ["vdn", "qmix", "qplex"])
def test_module_surface_step_agrees_with_fused_step(alg):
    """The fused train step (hand-derived backward kernels) against the same step through the drop-in modules' autograd
    Functions: two independent implementations over the same library, same losses and gradients."""
    batch = synthetic_batch(5, 8, 20, 4, 6, 12, 10)
    out = {}
    for through_modules in (False, True):
        kw = dict(num_kernel=2, adv_hypernet_embed=8, hypernet_embed=8) if alg == "qplex" else {}
        args = PU.make_args(alg, 4, 6, 12, 10, 20, train_through_modules=through_modules, **kw)
        learner, _ = PU.build_pair(args)
        losses = [learner.train({k: v.copy() for k, v in batch.items()}, i) for i in range(2)]
        out[through_modules] = (losses, learner._flat.grad.clone())
    (lf, gf), (lm, gm) = out[False], out[True]
    assert abs(lf[0] - lm[0]) <= 2e-6 * abs(lf[0]) and abs(lf[1] - lm[1]) <= 5e-5 * abs(lf[1]), (lf, lm)
    assert PU.rel_err(gm, gf) < 5e-5
