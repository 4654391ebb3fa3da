"""Single-state cooperative matrix game, drop-in for ``env/single_state_matrix_game.py:5-120``.

``TwoAgentsMatrixGame`` keeps the reference's SMAC-style single-environment API (host-side
bookkeeping: replay log, env_info, the fixed 9-episode ``get_episodes()`` batch).
``BatchedMatrixGame`` is the B200 form of the same game: ``n_envs`` independent instances
stepped by ONE kernel launch (``marl_matrix_game_step``) that writes the episode records
straight into device memory in the ReplayBuffer layout the learner consumes.
"""
from __future__ import annotations

import ctypes as C
import datetime

import numpy as np
import torch as th

from .. import _lib as L


class TwoAgentsMatrixGame:
    def __init__(self, payoff_table, replay_dir='./replay_dir'):
        self.payoff_table = np.array(payoff_table, dtype=np.float64)
        self._init_replay()
        self.current_episode = 0
        self.replay_dir = replay_dir
        self.n_actions = 3
        self.n_agents = 2
        self.state_shape = 1
        self.obs_shape = 1
        self.episode_limit = 1
        self.env_info = {"n_actions": self.n_actions, "n_agents": self.n_agents, "state_shape": self.state_shape,
                         "obs_shape": self.obs_shape, "episode_limit": self.episode_limit,
                         "env_name": "SingleStateMatrixGame"}

    def step(self, actions):
        """actions: list of the two agents' action ids -> (reward float64, terminated, info)."""
        reward = self.payoff_table[actions[0], actions[1]]
        log = self.replay[self.current_episode]
        log["obs"].append([1., 1.])
        log["state"].append(1.)
        log["actions"].append(actions)
        log["reward"].append(reward)
        log["episode_length"] += 1
        return reward, True, {}

    def get_obs(self):
        return [np.array([0.]), np.array([0.])]

    def get_state(self):
        return np.array([0.])

    def get_avail_actions(self):
        return [np.array([1, 1, 1]), np.array([1, 1, 1])]

    def get_avail_agent_actions(self, agent_id):
        return np.array([1, 1, 1])

    def reset(self):
        if self.replay[self.current_episode]["episode_length"] != 0:
            self.replay.append({"obs": [], "state": [], "actions": [], "reward": [], "episode_length": 0})
            self.current_episode += 1

    def close(self):
        self._init_replay()
        self.current_episode = 0

    def _init_replay(self):
        self.replay = [{"obs": [], "state": [], "actions": [], "reward": [], "episode_length": 0}]

    def save_replay(self):
        stamp = datetime.datetime.today().strftime('%Y-%m-%d_%H %M %S')
        np.save(self.env_info["env_name"] + stamp, np.array(self.replay, dtype=object), allow_pickle=True)

    def get_env_info(self):
        return self.env_info

    def get_episodes(self):
        """The 9 joint actions as a ready training batch (o = s = 1, agent-0 action varies slowest)."""
        n_ep, T, N, A = self.payoff_table.size, self.episode_limit, self.n_agents, self.n_actions
        joint = np.stack(np.meshgrid(*[np.arange(A)] * N, indexing="ij"), axis=-1).reshape(-1, N)
        u = joint.reshape(n_ep, T, N, 1).astype(np.int64)
        u_onehot = (u == np.arange(A).reshape(1, 1, 1, A)).astype(np.int64)
        ones_o = np.ones((n_ep, T, N, self.obs_shape), dtype=np.float64)
        ones_s = np.ones((n_ep, T, self.state_shape), dtype=np.float64)
        avail = np.ones((n_ep, T, N, A))
        return dict(o=ones_o.copy(), s=ones_s.copy(), u=u, r=self.payoff_table.reshape(n_ep, T, 1).copy(),
                    avail_u=avail.copy(), o_next=ones_o.copy(), s_next=ones_s.copy(), avail_u_next=avail.copy(),
                    u_onehot=u_onehot, padded=np.zeros((n_ep, T, 1)), terminated=np.ones((n_ep, T, 1)))


class BatchedMatrixGame:
    """n_envs TwoAgentsMatrixGame instances on one GPU, one kernel launch per step.

    ``step(actions)`` takes an int32/int64 CUDA tensor [n_envs, 2] and returns the episode batch as a
    dict of device tensors in the learner's layout (fp32, u int64, T = 1) -- the buffers are owned
    by the env and overwritten by the next step.  ``obs_value=0`` records what get_obs()/get_state()
    return (what a rollout stores); ``obs_value=1`` reproduces get_episodes().
    """

    KEYS = L.EPISODE_KEYS

    def __init__(self, payoff_table, n_envs, device="cuda", obs_value=0.0, keep_r64=False, validate=False):
        self.payoff_table = np.array(payoff_table, dtype=np.float64).reshape(3, 3)
        self.n_envs = int(n_envs)
        self.device = th.device(device)
        if self.device.type != "cuda":
            raise L.MarlLibraryError("BatchedMatrixGame needs a CUDA device: marl_b200 has no CPU path")
        self.obs_value = float(obs_value)
        self.validate = validate
        n, dev = self.n_envs, self.device
        f = lambda *s: th.empty(*s, dtype=th.float32, device=dev)
        self.buffers = dict(o=f(n, 1, 2, 1), s=f(n, 1, 1), u=th.empty(n, 1, 2, 1, dtype=th.int64, device=dev), r=f(n, 1, 1),
                            o_next=f(n, 1, 2, 1), s_next=f(n, 1, 1), avail_u=f(n, 1, 2, 3), avail_u_next=f(n, 1, 2, 3),
                            u_onehot=f(n, 1, 2, 3), padded=f(n, 1, 1), terminated=f(n, 1, 1))
        self.r64 = th.empty(n, dtype=th.float64, device=dev) if keep_r64 else None
        self._bad = th.zeros(1, dtype=th.int32, device=dev)
        self._payoff_c = (C.c_double * 9)(*self.payoff_table.reshape(-1).tolist())
        self._out = L.EpisodeF32()
        for k in self.KEYS:
            setattr(self._out, k, self.buffers[k].data_ptr())
        self.env_info = {"n_actions": 3, "n_agents": 2, "state_shape": 1, "obs_shape": 1, "episode_limit": 1,
                         "env_name": "SingleStateMatrixGame"}
        self.steps = 0

    BYTES_PER_ENV_STEP = 16 + 124   # int64 actions in + fp32 episode record out

    def get_env_info(self):
        return self.env_info

    def step(self, actions):
        if not (th.is_tensor(actions) and actions.is_cuda):
            raise L.MarlLibraryError("actions must be a CUDA tensor [n_envs, 2]")
        if actions.dtype not in (th.int32, th.int64) or actions.numel() != 2 * self.n_envs:
            raise ValueError("actions must be int32/int64 of shape [n_envs, 2]")
        if not actions.is_contiguous() or actions.data_ptr() % 16:
            actions = actions.contiguous().clone()
        sp = L.stream_ptr()
        nbytes = 8 if actions.dtype == th.int64 else 4
        if self.validate:
            self._bad.zero_()
            L.call("marl_matrix_game_validate_actions", actions.data_ptr(), nbytes, self.n_envs, self._bad.data_ptr(), sp)
            if int(self._bad.item()):
                raise IndexError("action index out of bounds for the 3x3 payoff table")
        L.call("marl_matrix_game_step", self._payoff_c, actions.data_ptr(), nbytes, self.n_envs, self.obs_value,
               C.byref(self._out), L.ptr(self.r64), sp)
        self.steps += 1
        return self.buffers
