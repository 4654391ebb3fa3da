"""Batched rollout on the matrix game: actions, rewards and the epsilon schedule must be what the
reference's sequential RolloutWorker + SharedMAC.choose_action produce under the same numpy seed, and the emitted
episodes must train the learner through the device replay buffer."""
import numpy as np
import pytest
import torch

from marl_b200.common.replaybuffer import ReplayBuffer
from marl_b200.env.single_state_matrix_game import BatchedMatrixGame
from marl_b200.rollout import BatchedRolloutWorker
from oracle import rollout_oracle as RO
from tests import golden_util as GU
from tests import parity_util as PU

pytestmark = pytest.mark.gpu
PAYOFF1 = [[8, -12, -12], [-12, 0, 0], [-12, 0, 0]]


def _setup(n_envs, **kw):
    args = PU.make_args("qmix", 2, 3, 1, 1, 1, **kw)
    learner, st = PU.build_pair(args)
    env = BatchedMatrixGame(PAYOFF1, n_envs, keep_r64=True)
    return args, learner, st, env


@pytest.mark.parametrize("scale", ["step", "episode"])
@pytest.mark.parametrize("epsilon", [1.0, 0.3, 0.0])
def test_actions_bit_exact_with_sequential_reference_rollout(scale, epsilon):
    n = 257
    args, learner, st, env = _setup(n, epsilon=epsilon, epsilon_anneal_scale=scale)
    args.anneal_epsilon = 0.01
    worker = BatchedRolloutWorker(env, learner.eval_net, args)
    np.random.seed(7)
    episodes, rewards, wins, steps = worker.generate_episodes()
    np.random.seed(7)
    agent = {k: v.detach().cpu() for k, v in learner.eval_net.agent.state_dict().items()}
    want_u, want_r, want_eps = RO.generate_episodes(agent, PAYOFF1, n, epsilon, 0.01, args.min_epsilon, scale)
    got_u = episodes["u"].reshape(n, 2).cpu().numpy()
    assert np.array_equal(got_u, want_u)
    assert np.array_equal(env.r64.cpu().numpy(), want_r)                     # float64 rewards exactly as the env returns them
    assert np.array_equal(rewards.cpu().numpy(), want_r.astype(np.float32))
    assert worker.epsilon == want_eps and steps == n and wins == [False] * n
    # the rest of the record is the rollout layout (rollout.py:122-149): zeros obs/state, all available, terminated, not padded
    assert float(episodes["o"].abs().max()) == 0.0 and float(episodes["s_next"].abs().max()) == 0.0
    assert float(episodes["terminated"].min()) == 1.0 and float(episodes["padded"].max()) == 0.0
    oh = torch.nn.functional.one_hot(episodes["u"].reshape(n, 2), 3).to(torch.float32)
    assert torch.equal(episodes["u_onehot"].reshape(n, 2, 3), oh)


def test_evaluate_is_greedy_and_keeps_epsilon():
    args, learner, st, env = _setup(64, epsilon=0.9)
    worker = BatchedRolloutWorker(env, learner.eval_net, args)
    np.random.seed(0)
    episodes, _, _, _ = worker.generate_episodes(evaluate=True)
    u = episodes["u"].reshape(64, 2)
    assert bool((u == u[0]).all()) and worker.epsilon == 0.9                 # same state everywhere -> same greedy joint action


def test_rollout_to_buffer_to_learner_matches_oracle():
    from oracle import marl_oracle as MO
    n = 512
    args, learner, st, env = _setup(n, epsilon=1.0, buffer_size=1024)
    worker = BatchedRolloutWorker(env, learner.eval_net, args)
    buf = ReplayBuffer(args)
    np.random.seed(3)
    for _ in range(3):                                                       # 3 x 512 episodes into 1024 slots: wraps
        episodes, _, _, _ = worker.generate_episodes()
        buf.store_episode(episodes)
    assert buf.current_size == 1024 and buf.current_idx == 512
    host = {k: v.cpu().numpy().astype(np.float64) for k, v in buf.buffers.items()}
    for step in range(3):
        view = buf.sample(128)
        batch = {k: host[k][view.idx_host] for k in host}
        loss = learner.train(view, step)
        oloss, _ = MO.train_step(st, batch, step)
        assert abs(loss - oloss) <= 1e-5 * abs(oloss)


def test_device_rng_mode_explores_at_the_requested_rate():
    args, learner, st, env = _setup(20000, epsilon=0.25)
    worker = BatchedRolloutWorker(env, learner.eval_net, args)
    torch.manual_seed(0)
    greedy, _, _, _ = worker.generate_episodes(evaluate=True)
    g = greedy["u"].reshape(-1, 2).clone()
    worker.epsilon, args.anneal_epsilon, worker.anneal_epsilon = 0.25, 0.0, 0.0
    ep, _, _, _ = worker.generate_episodes(rng="device")
    frac = float((ep["u"].reshape(-1, 2) != g).float().mean())               # explore (0.25) and draw another action (2/3)
    assert abs(frac - 0.25 * 2 / 3) < 0.02


# ------------------------------------------------------------------------------------------ multi-step episodes
KEYS11 = ("o", "s", "u", "r", "avail_u", "o_next", "s_next", "avail_u_next", "u_onehot", "padded", "terminated")


def _multistep_worker(z, n):
    from marl_b200.controller.share_params import SharedMAC
    from marl_b200.rollout import BatchedRolloutWorker
    from tests import parity_util as PU
    from tests.envs import CountdownGameBatched
    N, A, O, S, T = (int(x) for x in z["meta/dims"])
    args = PU.make_args("qmix", N, A, O, S, T)
    args.epsilon, args.anneal_epsilon, args.min_epsilon, args.epsilon_anneal_scale = 0.5, 0.01, 0.02, "step"
    mac = SharedMAC(args)
    mac.agent.load_state_dict(GU.group(z, "init/agent"))
    env = CountdownGameBatched(n, n_agents=N, n_actions=A, episode_limit=T)
    return BatchedRolloutWorker(env, mac, args), args


def test_multistep_batched_rollout_reproduces_reference_worker_golden():
    """Greedy (evaluate=True) episodes of the ragged multi-step test game: the batched worker against the episode
    batch the UNMODIFIED reference RolloutWorker produced (rollout.py:30-173): all 11 keys incl. the padding
    (zeros, padded = 1, terminated = 1), o_next / s_next / avail_u_next of the last real step, rewards, step count."""
    z = GU.load("rollout_multistep")
    n = int(z["meta/n"])
    worker, _ = _multistep_worker(z, n)
    ep, rewards, wins, steps = worker.generate_episodes(evaluate=True)
    for k in KEYS11:
        assert np.array_equal(ep[k].cpu().numpy().astype(np.float64), z[f"episodes/{k}"]), k
    assert np.array_equal(rewards.cpu().numpy(), z["rewards"]) and steps == int(z["steps"])


@pytest.mark.parametrize("n", [7, 64])
def test_multistep_batched_rollout_exploring_matches_sequential_oracle(n):
    """Exploring episodes under the host-supplied draws of the RNG contract: actions bit-exact against the sequential
    restatement (which is pinned to the reference worker by tests/test_rollout_oracle_cpu.py), every key equal, and no
    unavailable action is ever taken; the result goes through store_episode -> sample -> train."""
    from oracle import rollout_oracle as RO
    from tests.envs import CountdownGameHost
    z = GU.load("rollout_multistep")
    worker, args = _multistep_worker(z, n)
    N, A, T = args.n_agents, args.n_actions, args.episode_limit
    rng = np.random.RandomState(n)
    draws = (rng.rand(n, T, N), rng.rand(n, T, N))
    ep, rewards, _, steps = worker.generate_episodes(draws=draws)
    oep, orew, osteps = RO.rollout_multistep(GU.group(z, "init/agent"), CountdownGameHost(n_agents=N, n_actions=A, episode_limit=T), n,
                                             0.5, 0.01, 0.02, "step", draws=draws)
    assert np.array_equal(ep["u"].cpu().numpy().astype(np.float64), oep["u"])          # actions: bit-exact
    for k in KEYS11:
        assert np.array_equal(ep[k].cpu().numpy().astype(np.float64), oep[k]), k
    assert np.array_equal(rewards.cpu().numpy(), np.array(orew)) and steps == osteps
    taken = np.take_along_axis(oep["avail_u"], oep["u"].astype(np.int64), axis=3)[..., 0]
    assert np.all(taken[oep["padded"][:, :, 0] == 0] == 1)
    explored = (draws[0] < 0.5)[oep["padded"][:, :, 0] == 0].mean()
    assert explored > 0.2                                        # the draws really made agents explore
    # and the batch is consumable: device replay buffer -> learner
    from marl_b200.algorithm.q_learner import QLearner
    from marl_b200.common.replaybuffer import ReplayBuffer
    args.buffer_size = 128
    buf = ReplayBuffer(args)
    buf.store_episode(ep)
    learner = QLearner(worker.mac, args)
    np.random.seed(0)
    loss = learner.train(buf.sample(min(n, 16)), 0)
    assert np.isfinite(loss)
