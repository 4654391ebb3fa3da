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
        # one-step matrix game: the fused record-writing env kernel; anything else that speaks the batched protocol
        # (reset / get_obs / get_state / get_avail_actions / step(actions, active)) goes through the generic multi-step loop
        self._generic = not (hasattr(env, "payoff_table") and self.episode_limit == 1)

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

    def generate_episodes(self, n_episodes=None, evaluate=False, random_select=False, rng="numpy", draws=None):
        """Plays env.n_envs one-step episodes.  Returns (episodes, episode_rewards, win_tags, steps) like
        rollout.py:173; `episodes` is a dict of CUDA tensors [n_envs, 1, ...] owned by the environment."""
        n, N, A = self.env.n_envs, self.n_agents, self.n_actions
        if n_episodes not in (None, n):
            raise ValueError("a batched worker plays exactly env.n_envs episodes per call")
        if self._generic:
            return self._generate_multistep(evaluate, random_select, draws)
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

    # ---- multi-step episodes (rollout.py:52-149 for every instance at once) ------------------------------------------
    def _generate_multistep(self, evaluate, random_select, draws):
        """All env.n_envs instances step together until each has terminated or hit ``episode_limit``.

        Per step: ONE agent forward over the (instance, agent) rows with the carried hidden state and the last actions
        (share_params.py:37-60), ONE ``marl_epsgreedy_select`` launch with the availability mask (``-inf``, first
        maximum; share_params.py:62-72), one ``env.step``.  Finished instances are frozen and their remaining steps are
        the reference's padding (rollout.py:122-133): zeros everywhere, ``padded = 1``, ``terminated = 1``; the trailing
        observation / state / availability read after an instance's last step fill ``o_next`` / ``s_next`` /
        ``avail_u_next`` (rollout.py:106-120).

        RNG contract: ``draws=(explore_u, choice_u)``, two [n, episode_limit, N] arrays of uniforms in [0, 1) (host
        numpy or device tensors); agent a of instance e explores at step t iff explore_u[e, t, a] < epsilon_t and then
        takes the floor(choice_u * n_available)-th available action -- the same rule as oracle.rollout_oracle, so a
        seeded run is bit-identical to the sequential restatement.  Anything else draws the uniforms on the device.
        Every instance starts the call from the worker's current epsilon (they run side by side) and anneals it per
        step; afterwards the worker's epsilon has moved by the total number of environment steps, as n sequential
        episodes would have moved it."""
        if random_select:
            raise NotImplementedError("random_select is not used by the in-scope drivers")
        env, mac, dev = self.env, self.mac, self.mac.device
        n, N, A, O, S, T = env.n_envs, self.n_agents, self.n_actions, self.obs_shape, self.state_shape, self.episode_limit
        f = lambda *shape: th.zeros(*shape, dtype=th.float32, device=dev)
        ep = dict(o=f(n, T, N, O), s=f(n, T, S), u=th.zeros(n, T, N, 1, dtype=th.int64, device=dev), r=f(n, T, 1),
                  o_next=f(n, T, N, O), s_next=f(n, T, S), avail_u=f(n, T, N, A), avail_u_next=f(n, T, N, A),
                  u_onehot=f(n, T, N, A), padded=th.ones(n, T, 1, device=dev), terminated=th.ones(n, T, 1, device=dev))
        if isinstance(draws, (tuple, list)):
            explore_u, choice_u = (th.as_tensor(d, dtype=th.float64, device=dev) for d in draws)
        else:
            explore_u = th.rand(n, T, N, dtype=th.float64, device=dev)
            choice_u = th.rand(n, T, N, dtype=th.float64, device=dev)
        eps = 0.0 if evaluate else float(self.epsilon)
        if self.args.epsilon_anneal_scale == 'episode':
            eps = eps - self.anneal_epsilon if eps > self.min_epsilon else eps
        env.reset()
        hidden = th.zeros(n * N, self.args.rnn_hidden_dim, device=dev)
        last = f(n, N, A)
        active = th.ones(n, dtype=th.bool, device=dev)
        rewards = th.zeros(n, dtype=th.float64, device=dev)
        actions = th.empty(n, N, dtype=th.int64, device=dev)
        onehot = th.empty(n, N, A, dtype=th.float32, device=dev)
        steps_tot, prev_active = 0, None
        with th.no_grad():
            for t in range(T + 1):
                obs, state, avail = env.get_obs(), env.get_state(), env.get_avail_actions()
                if prev_active is not None:          # what the instances that acted at t-1 see afterwards
                    m = prev_active
                    ep["o_next"][:, t - 1] = th.where(m.view(n, 1, 1), obs, ep["o_next"][:, t - 1])
                    ep["s_next"][:, t - 1] = th.where(m.view(n, 1), state, ep["s_next"][:, t - 1])
                    ep["avail_u_next"][:, t - 1] = th.where(m.view(n, 1, 1), avail, ep["avail_u_next"][:, t - 1])
                if t == T:
                    break
                n_active = int(active.sum().item())
                if n_active == 0:
                    break
                q, _, hidden = mac.agent.unroll(obs.unsqueeze(1).contiguous(), last.unsqueeze(1).contiguous(), hidden, 0)
                explore = (explore_u[:, t] < eps).to(th.uint8).contiguous()
                cnt = avail.sum(-1).to(th.float64)
                kth = th.minimum((choice_u[:, t] * cnt).floor(), cnt - 1)
                rand_a = ((avail.cumsum(-1).to(th.float64) == (kth + 1).unsqueeze(-1)) & (avail > 0)).to(th.uint8).argmax(-1)
                L.call("marl_epsgreedy_select", n * N, A, q.reshape(n * N, A).contiguous().data_ptr(), avail.contiguous().data_ptr(),
                       explore.data_ptr(), rand_a.contiguous().data_ptr(), actions.data_ptr(), onehot.data_ptr(), L.stream_ptr())
                reward, term = env.step(actions, active)
                a3, a2 = active.view(n, 1, 1), active.view(n, 1)
                ep["o"][:, t] = th.where(a3, obs, ep["o"][:, t])
                ep["s"][:, t] = th.where(a2, state, ep["s"][:, t])
                ep["avail_u"][:, t] = th.where(a3, avail, ep["avail_u"][:, t])
                ep["u"][:, t] = th.where(a3, actions.unsqueeze(-1), ep["u"][:, t])
                ep["u_onehot"][:, t] = th.where(a3, onehot, ep["u_onehot"][:, t])
                ep["r"][:, t] = th.where(a2, reward.to(th.float32).view(n, 1), ep["r"][:, t])
                ep["terminated"][:, t] = th.where(a2, term.to(th.float32).view(n, 1), ep["terminated"][:, t])
                ep["padded"][:, t] = th.where(a2, th.zeros_like(ep["padded"][:, t]), ep["padded"][:, t])
                rewards += th.where(active, reward.to(th.float64), th.zeros_like(rewards))
                last = th.where(a3, onehot, last)
                steps_tot += n_active
                prev_active = active
                active = active & ~term
                if self.args.epsilon_anneal_scale == 'step':
                    eps = eps - self.anneal_epsilon if eps > self.min_epsilon else eps
        if not evaluate:
            cur = float(self.epsilon)
            if self.args.epsilon_anneal_scale == 'step':
                for _ in range(min(steps_tot, 1 << 20)):
                    if not cur > self.min_epsilon:
                        break
                    cur -= self.anneal_epsilon
            else:
                for _ in range(n):
                    cur = cur - self.anneal_epsilon if cur > self.min_epsilon else cur
            self.epsilon = cur
        mac.hidden_states = hidden
        return ep, rewards, [False] * n, steps_tot
