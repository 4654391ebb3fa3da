"""TEST INFRASTRUCTURE ONLY -- regenerates tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (``python oracle/make_golden.py``); needs /root/reference,
which does not exist on the GPU box -- hence the committed fixtures.  The reference is
imported with the two shims of SURVEY.md appendix B (gym stub, np.float/np.long) and
driven through its own public surface: ``SharedMAC``, ``QLearner.train`` /
``QTRANLearner.train``, ``TwoAgentsMatrixGame``.  Each fixture stores inputs (batch,
initial weights), the loss returned by every ``train()`` call, the clipped gradients
left in ``p.grad`` after the first step, the final eval/target weights, and the
step-0 intermediates obtained by replaying the reference's own methods in the order
``train()`` calls them (q_learner.py:95-114).
"""
import copy
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch as th

REF = os.environ.get("MARL_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.modules.setdefault("gym", types.SimpleNamespace(Env=object))
np.float = float
np.long = np.int64

from common.arguments import get_mixer_args  # noqa: E402
from controller.share_params import SharedMAC  # noqa: E402
from algorithm.q_learner import QLearner  # noqa: E402
from algorithm.qtran_learner import QTRANLearner  # noqa: E402
from env.single_state_matrix_game import TwoAgentsMatrixGame  # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from marl_b200.synthetic import synthetic_batch  # noqa: E402

OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
PAYOFF1 = [[8, -12, -12], [-12, 0, 0], [-12, 0, 0]]          # matrix_game_test.py:43-45


def ref_args(alg, N, A, O, S, T, **kw):
    a = SimpleNamespace(RTW=False, alg=alg, map="synthetic", last_action=True, reuse_network=True,
                        gamma=0.99, optimizer="RMS", model_dir="/tmp/marl_golden_model", result_dir="/tmp",
                        cuda=False, load_model=False, evaluate=False, evaluate_epoch=0, replay_dir="", n_episodes=1)
    get_mixer_args(a)
    a.n_agents, a.n_actions, a.obs_shape, a.state_shape, a.episode_limit = N, A, O, S, T
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def dump_sd(prefix, sd, out):
    for k, v in sd.items():
        out[f"{prefix}/{k}"] = v.detach().cpu().numpy().copy()


def run_case(name, alg, batch, N, A, O, S, T, n_steps=3, seed=0, **kw):
    th.manual_seed(seed)
    args = ref_args(alg, N, A, O, S, T, **kw)
    mac = SharedMAC(args)
    learner = QTRANLearner(mac, args) if alg == "qtran_base" else QLearner(mac, args)
    out = {"meta/alg": np.array(alg), "meta/optimizer": np.array(args.optimizer),
           "meta/dims": np.array([N, A, O, S, T]), "meta/n_steps": np.array(n_steps),
           "meta/double_q": np.array(int(args.double_q)), "meta/lr": np.array(args.lr),
           "meta/target_update_cycle": np.array(args.target_update_cycle)}
    for k in ("num_kernel", "adv_hypernet_embed", "hypernet_embed", "qtran_hidden_dim", "hyper_hidden_dim"):
        out[f"meta/{k}"] = np.array(getattr(args, k))
    out["meta/two_hyper_layers"] = np.array(int(args.two_hyper_layers))
    out["meta/adv_hypernet_layers"] = np.array(int(getattr(args, "adv_hypernet_layers", 3)))
    for k, v in batch.items():
        out[f"batch/{k}"] = np.asarray(v)
    dump_sd("init/agent", mac.agent.state_dict(), out)
    dump_sd("init/mixer", learner.mixer.state_dict(), out)
    if alg == "qtran_base":
        dump_sd("init/v", learner.v.state_dict(), out)
        dump_sd("init/q_sum_mixer", learner.q_sum_mixer.state_dict(), out)

    # --- step-0 intermediates, replaying q_learner.py:95-114 / qtran_learner.py:95-114 on a deep copy
    probe = copy.deepcopy(learner)
    b = {k: np.array(v) for k, v in batch.items()}
    b, L = probe.get_max_episode_len(b)
    for k in b:
        b[k] = th.tensor(b[k], dtype=th.long if k == "u" else th.float32)
    B = b["o"].shape[0]
    with th.no_grad():
        probe.eval_net.init_hidden(B)
        q_evals, hid = probe.eval_net.get_current_q_values(b, L)
        probe.target_net.init_hidden(B)
        q_targets, hid_t = probe.target_net.get_next_q_values(b, L)
        q_targets[b["avail_u_next"] == 0.0] = -9999999
        out["step0/L"] = np.array(L)
        out["step0/q_evals"] = q_evals.numpy().copy()
        out["step0/hidden_evals"] = hid.numpy().copy()
        out["step0/q_targets"] = q_targets.numpy().copy()
        if alg != "qtran_base" and args.double_q:
            q_en, _ = probe.eval_net.get_next_q_values(b, L)      # no init_hidden: carried hidden
            q_en[b["avail_u_next"] == 0] = -9999999
            out["step0/q_evals_next"] = q_en.numpy().copy()
            out["step0/cur_max_actions"] = th.argmax(q_en, dim=3).numpy().copy()
        if alg in ("vdn", "qmix"):
            qc = th.gather(q_evals, 3, b["u"]).squeeze(3)
            out["step0/q_tot"] = probe.mixer(qc, b["s"]).numpy().copy()
        if alg == "qtran_base":
            out["step0/hidden_targets"] = hid_t.numpy().copy()

    # --- the real thing
    losses = []
    for step in range(n_steps):
        losses.append(learner.train({k: np.array(v) for k, v in batch.items()}, step))
        if step == 0:
            names = [("agent", k) for k, _ in mac.agent.named_parameters()]
            names += [("mixer", k) for k, _ in learner.mixer.named_parameters()]
            if alg == "qtran_base":
                names += [("v", k) for k, _ in learner.v.named_parameters()]
                names += [("q_sum_mixer", k) for k, _ in learner.q_sum_mixer.named_parameters()]
            assert len(names) == len(learner.params)
            for (g, k), p in zip(names, learner.params):
                if p.grad is not None:
                    out[f"clipped_grad/{g}/{k}"] = p.grad.detach().numpy().copy()
    out["loss"] = np.array(losses, dtype=np.float64)
    dump_sd("final/agent", mac.agent.state_dict(), out)
    dump_sd("final/mixer", learner.mixer.state_dict(), out)
    dump_sd("final_target/agent", learner.target_net.agent.state_dict(), out)
    dump_sd("final_target/mixer", learner.target_mixer.state_dict(), out)
    if alg == "qtran_base":
        dump_sd("final/v", learner.v.state_dict(), out)
    if N == 2 and A == 3 and O == 1 and S == 1 and alg != "qtran_base":
        qt, qi, qj = learner.get_q_and_q_tot_table()                 # q_learner.py:211-262
        out["table/q_tot"], out["table/q_i"], out["table/q_j"] = qt, qi, qj
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: losses={losses} -> {os.path.getsize(path) / 1024:.0f} KiB")


def env_case():
    """TwoAgentsMatrixGame golden: get_episodes() + step() for all joint actions, 3 payoff tables."""
    tables = [PAYOFF1, [[8, -12, -12], [-12, 6, 0], [-12, 0, 6]], [[8, 3, 2], [-12, -13, -14], [-12, -13, -14]]]
    out = {}
    for i, tab in enumerate(tables):
        env = TwoAgentsMatrixGame(tab)
        for k, v in env.get_episodes().items():
            out[f"t{i}/episodes/{k}"] = np.asarray(v)
        rew = np.zeros((3, 3))
        for a0 in range(3):
            for a1 in range(3):
                env.reset()
                r, term, info = env.step([a0, a1])
                assert term is True and info == {}
                rew[a0, a1] = r
        out[f"t{i}/step_reward"] = rew
        out[f"t{i}/payoff"] = np.array(tab, dtype=np.float64)
        out[f"t{i}/obs"] = np.array(env.get_obs())
        out[f"t{i}/state"] = np.array(env.get_state())
        out[f"t{i}/avail"] = np.array(env.get_avail_actions())
        out[f"t{i}/replay_len"] = np.array(len(env.replay))
        info = env.get_env_info()
        out[f"t{i}/env_info"] = np.array([info[k] for k in ("n_actions", "n_agents", "state_shape", "obs_shape", "episode_limit")])
    np.savez_compressed(os.path.join(OUT, "matrix_game_env.npz"), **out)
    print("matrix_game_env ok")


def choose_action_case():
    """SharedMAC.choose_action (share_params.py:37-72) driven like rollout.py:60-76 for 6 steps on the literal 3s5z
    observations / availability masks the reference keeps in test_file/choose_action_test.py:7-208: greedy steps
    (epsilon 0) and exploring steps (epsilon 0.5) under a fixed numpy seed.  Stores inputs, actions and the hidden state
    the controller carries after every step."""
    src = open(os.path.join(REF, "test_file", "choose_action_test.py")).read()
    i, j = src.index("obs = ["), src.index("avail_actions = ")
    obs0 = np.stack(eval(src[i + 6:j].strip(), {"np": np}))                       # [8, 128] float32 literals
    avail0 = np.array(eval(src[j + len("avail_actions = "):src.index("\n\n", j)]))   # [8, 14]
    N, O = obs0.shape
    A = avail0.shape[1]
    th.manual_seed(0)
    args = ref_args("qmix", N, A, O, 216, 150)
    mac = SharedMAC(args)
    out = {"meta/dims": np.array([N, A, O]), "meta/seed": np.array(7)}
    dump_sd("init/agent", mac.agent.state_dict(), out)
    n_steps = 6
    eps = np.array([0.0, 0.0, 0.5, 0.5, 0.0, 0.5])
    rng = np.random.RandomState(99)
    obs = np.stack([obs0 * (1.0 + 0.05 * t) + 0.01 * rng.randn(N, O).astype(np.float32) for t in range(n_steps)]).astype(np.float32)
    avail = np.stack([np.roll(avail0, t, axis=0) for t in range(n_steps)]).astype(np.float64)
    avail[:, :, 0] = np.where(avail.sum(-1) == 0, 1.0, avail[:, :, 0])            # at least one action everywhere
    avail[3, 2] = 0; avail[3, 2, 6] = 1                                              # a single available action
    np.random.seed(7)
    mac.init_hidden(1)
    last = np.zeros((N, A))
    actions = np.zeros((n_steps, N), dtype=np.int64)
    hidden = np.zeros((n_steps, N, 64), dtype=np.float32)
    with th.no_grad():
        for t in range(n_steps):
            for a in range(N):
                act = int(mac.choose_action(obs[t, a], last[a], a, avail[t, a], eps[t]))
                actions[t, a] = act
                last[a] = np.eye(A)[act]
            hidden[t] = mac.hidden_states[0].numpy()
    out.update({"obs": obs, "avail": avail, "eps": eps, "actions": actions, "hidden": hidden})
    np.savez_compressed(os.path.join(OUT, "choose_action_3s5z.npz"), **out)
    print("choose_action_3s5z: actions", actions.tolist())


def checkpoint_case():
    """The trained 2s3z checkpoints the reference ships (model/{vdn,qplex,qtran_base}/2s3z) as realistic-valued weights:
    copied as data fixtures, plus the losses of two reference train() steps on the seed-0 2s3z-shaped batch."""
    dst = os.path.join(OUT, "ckpt")
    os.makedirs(dst, exist_ok=True)
    picks = {"vdn": ("2", ["rnn"]), "qplex": ("3", ["rnn", "mixer"]), "qtran_base": ("3", ["rnn", "mixer", "v"])}
    batch = synthetic_batch(0, 32, 120, 5, 11, 80, 120)
    out = {}
    for alg, (num, parts) in picks.items():
        for part in parts:
            # (the shipped pickles carry CUDA storages: re-saved as CPU state_dicts, values untouched)
            sd = th.load(os.path.join(REF, "model", alg, "2s3z", f"{num}_{part}_net_params.pkl"), map_location="cpu", weights_only=True)
            th.save({k: v.clone() for k, v in sd.items()}, os.path.join(dst, f"{alg}_{part}.pkl"))
        th.manual_seed(0)
        args = ref_args(alg, 5, 11, 80, 120, 120)
        mac = SharedMAC(args)
        learner = QTRANLearner(mac, args) if alg == "qtran_base" else QLearner(mac, args)
        mac.agent.load_state_dict(th.load(os.path.join(dst, f"{alg}_rnn.pkl"), weights_only=True))
        if "mixer" in parts:
            learner.mixer.load_state_dict(th.load(os.path.join(dst, f"{alg}_mixer.pkl"), weights_only=True))
        if "v" in parts:
            learner.v.load_state_dict(th.load(os.path.join(dst, f"{alg}_v.pkl"), weights_only=True))
        learner.target_net.load_state(mac)
        learner.target_mixer.load_state_dict(learner.mixer.state_dict())
        losses = [learner.train({k: np.array(v) for k, v in batch.items()}, step) for step in range(2)]
        out[f"{alg}/loss"] = np.array(losses, dtype=np.float64)
        print("checkpoint", alg, losses)
    np.savez_compressed(os.path.join(OUT, "checkpoint_losses.npz"), **out)


def seeds_case():
    """QMIX at the full 2s3z shape: per-step losses of 10 reference train() steps (seed 0) and of 3 steps for data seeds
    1..4 (1 and 10 steps, seeds 0-4).  Only the losses are stored (the inputs are regenerated from the seed)."""
    out = {}
    for seed in range(5):
        n_steps = 10 if seed == 0 else 3
        th.manual_seed(seed)
        args = ref_args("qmix", 5, 11, 80, 120, 120)
        mac = SharedMAC(args)
        learner = QLearner(mac, args)
        for k, v in mac.agent.state_dict().items():
            out[f"s{seed}/agent/{k}"] = v.numpy().copy()
        for k, v in learner.mixer.state_dict().items():
            out[f"s{seed}/mixer/{k}"] = v.numpy().copy()
        batches = [synthetic_batch(100 * seed + i, 32, 120, 5, 11, 80, 120) for i in range(2)]
        losses = [learner.train({k: np.array(v) for k, v in batches[i % 2].items()}, i) for i in range(n_steps)]
        out[f"s{seed}/loss"] = np.array(losses, dtype=np.float64)
        print("seed", seed, losses)
    np.savez_compressed(os.path.join(OUT, "qmix_2s3z_seeds.npz"), **out)


def rollout_case():
    """The UNMODIFIED reference RolloutWorker (rollout.py:30-173) playing 7 greedy (evaluate=True) episodes of the
    multi-step test game of tests/envs.py: ragged lengths, step-dependent availability, padding to episode_limit.
    Pins the episode layout, the padding rules (rollout.py:122-133) and the greedy actions for the batched worker."""
    from rollout import RolloutWorker
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    from tests.envs import CountdownGameHost
    env = CountdownGameHost()
    info = env.get_env_info()
    th.manual_seed(5)
    args = ref_args("qmix", info["n_agents"], info["n_actions"], info["obs_shape"], info["state_shape"], info["episode_limit"])
    args.epsilon, args.anneal_epsilon, args.min_epsilon, args.epsilon_anneal_scale = 0.5, 0.01, 0.02, "step"
    mac = SharedMAC(args)
    worker = RolloutWorker(env, mac, args)
    n = 7
    np.random.seed(3)
    episodes, rewards, wins, steps = worker.generate_episodes(n_episodes=n, evaluate=True)
    out = {"meta/dims": np.array([info["n_agents"], info["n_actions"], info["obs_shape"], info["state_shape"], info["episode_limit"]]),
           "meta/n": np.array(n), "rewards": np.array(rewards, dtype=np.float64), "steps": np.array(steps)}
    dump_sd("init/agent", mac.agent.state_dict(), out)
    for k, v in episodes.items():
        out[f"episodes/{k}"] = np.asarray(v, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "rollout_multistep.npz"), **out)
    lens = (1 - out["episodes/padded"][:, :, 0]).sum(1)
    print("rollout_multistep: steps", steps, "lengths", lens.tolist(), "rewards", rewards)


def main():
    os.makedirs(OUT, exist_ok=True)
    def qplex_layers_case():
        tiny = dict(N=3, A=4, O=5, S=6, T=6)
        tb = synthetic_batch(0, 4, tiny["T"], tiny["N"], tiny["A"], tiny["O"], tiny["S"])
        for nl in (1, 2):
            run_case(f"tiny_qplex_layers{nl}", "qplex", tb, **tiny, target_update_cycle=2, num_kernel=2, adv_hypernet_embed=8,
                     hypernet_embed=8, adv_hypernet_layers=nl)

    def separated_case():
        """SeparatedMAC (share_params.py:389-610: one RNNQNet per agent, reuse_network=False as runner.py:24-26 builds it)
        under the UNMODIFIED QLearner for 3 QMIX steps (< target_update_cycle: the reference's own target sync calls
        ``other_mac.agent.state_dict()`` on a list and cannot run with this controller)."""
        from controller.share_params import SeparatedMAC
        tiny = dict(N=3, A=4, O=5, S=6, T=6)
        tb = synthetic_batch(2, 4, tiny["T"], tiny["N"], tiny["A"], tiny["O"], tiny["S"])
        for alg in ("qmix", "vdn"):
            th.manual_seed(0)
            args = ref_args(alg, tiny["N"], tiny["A"], tiny["O"], tiny["S"], tiny["T"], reuse_network=False)
            mac = SeparatedMAC(args)
            learner = QLearner(mac, args)
            out = {"meta/alg": np.array(alg), "meta/dims": np.array([tiny[k] for k in ("N", "A", "O", "S", "T")])}
            for k, v in tb.items():
                out[f"batch/{k}"] = np.asarray(v)
            for n, ag in enumerate(mac.agent):
                dump_sd(f"init/agent.{n}", ag.state_dict(), out)
            dump_sd("init/mixer", learner.mixer.state_dict(), out)
            probe = copy.deepcopy(learner)
            b = {k: np.array(v) for k, v in tb.items()}
            b, Lc = probe.get_max_episode_len(b)
            for k in b:
                b[k] = th.tensor(b[k], dtype=th.long if k == "u" else th.float32)
            with th.no_grad():
                probe.eval_net.init_hidden(4)
                q_evals, hid = probe.eval_net.get_current_q_values(b, Lc)
                q_next, hid_list = probe.eval_net.get_next_q_values(b, Lc)          # carried hidden; returns the per-agent LIST
                out["step0/q_evals"], out["step0/hidden_evals"] = q_evals.numpy().copy(), hid.numpy().copy()
                out["step0/q_evals_next"] = q_next.numpy().copy()
                out["step0/next_hidden_list"] = np.stack([h.numpy() for h in hid_list])
            # train(): with the installed torch (2.11) the reference's per-agent in-place hidden-state write-back
            # (share_params.py:538-539) invalidates the autograd graph and loss.backward() raises -- recorded, so that
            # the fixture documents that only the FORWARD surface of this controller can be pinned to the reference
            try:
                losses = [learner.train({k: np.array(v) for k, v in tb.items()}, step) for step in range(3)]
                out["loss"] = np.array(losses, dtype=np.float64)
            except RuntimeError as e:
                losses = None
                out["train_error"] = np.array(str(e)[:200])
            np.savez_compressed(os.path.join(OUT, f"separated_{alg}.npz"), **out)
            print("separated", alg, losses if losses is not None else "train() raised: " + str(out["train_error"])[:80])

    extra = {"separated": separated_case, "qplex_layers": qplex_layers_case, "rollout": rollout_case, "choose_action": choose_action_case, "checkpoints": checkpoint_case, "seeds": seeds_case}
    picked = [k for k in sys.argv[1:] if k in extra]
    if picked:                              # round-2 fixtures (the round-1 fixtures stay as committed)
        for k in picked:
            extra[k]()
        return
    tiny = dict(N=3, A=4, O=5, S=6, T=6)
    tb = synthetic_batch(0, 4, tiny["T"], tiny["N"], tiny["A"], tiny["O"], tiny["S"])
    small_qplex = dict(num_kernel=2, adv_hypernet_embed=8, hypernet_embed=8)
    if "two_hyper" in sys.argv[1:]:        # only the case added for SURVEY 8(f) N4 (the other fixtures stay as committed)
        run_case("tiny_qmix_two_hyper", "qmix", tb, **tiny, target_update_cycle=2, two_hyper_layers=True, hyper_hidden_dim=16)
        return
    env_case()
    run_case("tiny_vdn_rms", "vdn", tb, **tiny, target_update_cycle=2)
    run_case("tiny_qmix_rms", "qmix", tb, **tiny, target_update_cycle=2)
    run_case("tiny_qmix_adam", "qmix", tb, **tiny, optimizer="Adam", target_update_cycle=2)
    run_case("tiny_qmix_nodouble", "qmix", tb, **tiny, double_q=False, n_steps=2)
    run_case("tiny_qplex_rms", "qplex", tb, **tiny, target_update_cycle=2, **small_qplex)
    run_case("tiny_qtran_rms", "qtran_base", tb, **tiny, target_update_cycle=2, qtran_hidden_dim=16)
    # config 1: QMIX on the matrix game's fixed 9-episode batch (matrix_game_test.py:77-93)
    mg = TwoAgentsMatrixGame(PAYOFF1).get_episodes()
    run_case("matrix_qmix_rms", "qmix", mg, N=2, A=3, O=1, S=1, T=1, n_steps=5, lr=1e-3)
    run_case("matrix_vdn_rms", "vdn", mg, N=2, A=3, O=1, S=1, T=1, n_steps=3, lr=1e-3)
    # one ragged mid-size case: first episode shorter than the limit -> truncation L < T
    rb = synthetic_batch(3, 5, 9, 2, 5, 7, 4, full_length_first=False, min_len=2)
    run_case("ragged_qmix_rms", "qmix", rb, N=2, A=5, O=7, S=4, T=9, n_steps=2)
    run_case("tiny_qmix_two_hyper", "qmix", tb, **tiny, target_update_cycle=2, two_hyper_layers=True, hyper_hidden_dim=16)


if __name__ == "__main__":
    main()
