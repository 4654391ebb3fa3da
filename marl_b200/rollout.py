"""Batched data collection, the many-environments counterpart of ``rollout.py:3-173``.

The reference's ``RolloutWorker.generate_episodes`` plays ONE environment, asks ``mac.choose_action`` for one
agent at a time (a batch-1 network forward each, ``rollout.py:60-76``) and concatenates episodes with
``np.concatenate``.  ``BatchedRolloutWorker`` plays every instance of a batched environment
(``BatchedMatrixGame``) at once: one agent forward for all (env, agent) rows, one epsilon-greedy selection
launch, one environment step launch; the episode batch comes out on the device in the ReplayBuffer layout
(``rollout.py:135-149``) and can go straight into ``marl_b200.common.replaybuffer.ReplayBuffer``.

RNG contract.  ``rng="numpy"`` draws, on the host, exactly what the reference would draw and in its order --
per episode, per agent: one ``np.random.uniform()``, then ``np.random.choice(available)`` only when exploring
(``share_params.py:67-68``) -- so a seeded run takes bit-identical actions; none of these draws depends on the
Q-values, which is what makes hoisting them out of the per-agent loop exact.  ``rng="device"`` draws with
``torch`` on the GPU for throughput (same distribution, different stream).  Epsilon follows the reference's
schedule (``rollout.py:47-49, 103-104, 166-167``): annealed per step or per episode, carried across calls.
"""
from __future__ import annotations

import numpy as np
import torch as th

from . import _lib as L


class BatchedRolloutWorker:
    def __init__(self, env, mac, args):
        self.env = env
        self.mac = mac
        self.episode_limit = args.episode_limit
        self.n_actions = args.n_actions
        self.n_agents = args.n_agents
        self.state_shape = args.state_shape
        self.obs_shape = args.obs_shape
        self.args = args
        self.epsilon = args.epsilon
        self.anneal_epsilon = args.anneal_epsilon
        self.min_epsilon = args.min_epsilon
        if self.episode_limit != 1:
            raise NotImplementedError("the batched environment in scope (single_state_matrix_game) has one-step episodes")

    def _epsilons(self, n, evaluate):
        """Per-episode epsilon, advanced exactly like n sequential episodes of the reference would."""
        eps = np.empty(n, dtype=np.float64)
        cur = self.epsilon
        for e in range(n):
            epsilon = 0 if evaluate else cur
            if self.args.epsilon_anneal_scale == 'episode':
                epsilon = epsilon - self.anneal_epsilon if epsilon > self.min_epsilon else epsilon
            eps[e] = epsilon                                   # the single step of the episode acts with this value
            if self.args.epsilon_anneal_scale == 'step':
                epsilon = epsilon - self.anneal_epsilon if epsilon > self.min_epsilon else epsilon
            if not evaluate:
                cur = epsilon
        return eps, cur

    def generate_episodes(self, n_episodes=None, evaluate=False, random_select=False, rng="numpy"):
        """Plays env.n_envs one-step episodes.  Returns (episodes, episode_rewards, win_tags, steps) like
        rollout.py:173; `episodes` is a dict of CUDA tensors [n_envs, 1, ...] owned by the environment."""
        n, N, A = self.env.n_envs, self.n_agents, self.n_actions
        if n_episodes not in (None, n):
            raise ValueError("a batched worker plays exactly env.n_envs episodes per call")
        if random_select:
            raise NotImplementedError("random_select is not used by the in-scope drivers")
        dev = self.mac.device
        eps, eps_after = self._epsilons(n, evaluate)
        avail_idx = np.arange(A)                               # every action of the matrix game is available
        if rng == "numpy":
            explore = np.zeros((n, N), dtype=np.uint8)
            rand_a = np.zeros((n, N), dtype=np.int64)
            for e in range(n):                                 # the reference's draw order: episode-major, agent-minor
                for a in range(N):
                    if np.random.uniform() < eps[e]:
                        explore[e, a] = 1
                        rand_a[e, a] = np.random.choice(avail_idx)
            explore_d = th.from_numpy(explore).to(dev, non_blocking=True)
            rand_d = th.from_numpy(rand_a).to(dev, non_blocking=True)
        else:
            eps_d = th.from_numpy(eps).to(dev).to(th.float32).unsqueeze(1)
            explore_d = (th.rand(n, N, device=dev) < eps_d).to(th.uint8).contiguous()
            rand_d = th.randint(0, A, (n, N), device=dev, dtype=th.int64)
        # the worker's view of the game at the only step: obs = get_obs() (zeros), no last action, fresh hidden state
        obs = th.zeros(n, 1, N, self.obs_shape, device=dev)
        last = th.zeros(n, 1, N, A, device=dev)
        self.mac.init_hidden(n)
        q, _ = self.mac.get_current_q_values({"o": obs, "u_onehot": last}, 1)          # [n, 1, N, A]
        actions = th.empty(n, N, dtype=th.int64, device=dev)
        L.call("marl_epsgreedy_select", n * N, A, q.contiguous().data_ptr(), None, explore_d.data_ptr(), rand_d.data_ptr(),
               actions.data_ptr(), None, L.stream_ptr())
        episodes = self.env.step(actions)
        self._keep = (explore_d, rand_d, q)
        if not evaluate:
            self.epsilon = eps_after
        rewards = episodes["r"].reshape(n)
        return episodes, rewards, [False] * n, n
