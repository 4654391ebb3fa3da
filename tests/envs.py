"""TEST INFRASTRUCTURE -- a small multi-step cooperative game with ragged episode lengths and step-dependent action
availability, in two forms with identical arithmetic:

* ``CountdownGameHost``: the SMAC-style single-environment interface the reference's ``RolloutWorker`` drives
  (``rollout.py:30-173``: reset / get_obs / get_state / get_avail_actions / get_avail_agent_actions / step);
  ``reset()`` walks through instances 0, 1, 2, ... so that n sequential episodes play n different games;
* ``CountdownGameBatched``: all instances at once on the GPU, the protocol ``BatchedRolloutWorker`` expects.

Every observation / state value is a multiple of 1/8, rewards are multiples of 1/2: exact in fp32 and float64.
"""
import numpy as np
import torch


class _Rules:
    def __init__(self, n_agents=3, n_actions=5, episode_limit=6):
        self.N, self.A, self.T = n_agents, n_actions, episode_limit
        self.O, self.S = 4, 3

    @staticmethod
    def _ar(xp, n, like):
        return xp.arange(n, device=like.device) if xp is torch else xp.arange(n)

    # all rules on integer arrays: c [n] step counter, p [n, N] per-agent position, act [n, N]
    def obs(self, xp, c, p):
        a = self._ar(xp, self.N, p).reshape(1, self.N)
        cc = c.reshape(-1, 1) + 0 * p
        return xp.stack([cc, p, a + 1 + 0 * p, (cc * p) % 5], -1) / 8.0

    def state(self, xp, c, p):
        return xp.stack([c, p.sum(-1), (p * p).sum(-1) % 7], -1) / 8.0

    def avail(self, xp, c, p):
        a = self._ar(xp, self.N, p).reshape(1, self.N, 1)
        k = self._ar(xp, self.A, p).reshape(1, 1, self.A)
        blocked = ((c.reshape(-1, 1, 1) + a + k + p[..., None]) % 3 == 0) & (k != 0)
        return 1 - blocked * 1

    def reward(self, act, p_after):
        return ((act.sum(-1) + p_after.sum(-1)) % 4) * 0.5 - 0.5

    def done(self, c_after, p_after):
        return ((p_after.sum(-1) + 2 * c_after) % 5 == 0) | (c_after >= self.T)


class CountdownGameHost(_Rules):
    def __init__(self, **kw):
        super().__init__(**kw)
        self.instance = -1
        self.c = self.p = None

    def get_env_info(self):
        return {"n_actions": self.A, "n_agents": self.N, "state_shape": self.S, "obs_shape": self.O, "episode_limit": self.T}

    def reset(self):
        self.instance += 1
        self.c = np.zeros(1, dtype=np.int64)
        self.p = ((self.instance + np.arange(self.N)) % self.A).reshape(1, self.N).astype(np.int64)

    def get_obs(self):
        return [row for row in self.obs(np, self.c, self.p)[0]]

    def get_state(self):
        return self.state(np, self.c, self.p)[0]

    def get_avail_actions(self):
        return [list(row) for row in self.avail(np, self.c, self.p)[0]]

    def get_avail_agent_actions(self, agent_id):
        return list(self.avail(np, self.c, self.p)[0][agent_id])

    def step(self, actions):
        act = np.array([int(a) for a in actions], dtype=np.int64).reshape(1, self.N)
        self.p = (self.p + act) % self.A
        self.c = self.c + 1
        return float(self.reward(act, self.p)[0]), bool(self.done(self.c, self.p)[0]), {}

    def close(self):
        pass

    def save_replay(self):
        pass


class CountdownGameBatched(_Rules):
    def __init__(self, n_envs, device="cuda", **kw):
        super().__init__(**kw)
        self.n_envs, self.device = n_envs, torch.device(device)
        self.n_agents, self.n_actions, self.obs_shape, self.state_shape, self.episode_limit = self.N, self.A, self.O, self.S, self.T
        self.c = self.p = None

    def reset(self):
        e = torch.arange(self.n_envs, device=self.device).reshape(-1, 1)
        self.c = torch.zeros(self.n_envs, dtype=torch.int64, device=self.device)
        self.p = (e + torch.arange(self.N, device=self.device).reshape(1, -1)) % self.A

    def get_obs(self):
        return self.obs(torch, self.c, self.p).to(torch.float32)

    def get_state(self):
        return self.state(torch, self.c, self.p).to(torch.float32)

    def get_avail_actions(self):
        return self.avail(torch, self.c, self.p).to(torch.float32)

    def step(self, actions, active):
        """Advances the ACTIVE instances only; returns (reward [n] float64, terminated [n] bool)."""
        act = actions.to(torch.int64)
        live = active.reshape(-1, 1)
        self.p = torch.where(live, (self.p + act) % self.A, self.p)
        self.c = torch.where(active, self.c + 1, self.c)
        return self.reward(act, self.p).to(torch.float64), self.done(self.c, self.p)
