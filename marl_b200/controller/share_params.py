"""Parameter-sharing multi-agent controller, drop-in for ``controller/share_params.py:8-182``.

Same surface as the reference's ``SharedMAC`` (``init_hidden``, ``get_current_q_values``,
``get_next_q_values``, ``choose_action``, ``parameters``, ``load_state``, ``save_models``,
``load_models``, ``cuda``) and the same stateful ``hidden_states`` semantics -- the hidden
state is carried between successive ``get_*_q_values`` calls until ``init_hidden`` -- but each
T-step unroll is ONE call into libmarl_b200 instead of a Python loop over timesteps, and the
input rows [obs | last_action | agent_id] (share_params.py:84-112) are never materialised.
"""
from __future__ import annotations

import numpy as np
import torch as th

from .. import _lib as L
from ..network.q_network import RNNQNet


def _as_f32_cuda(x, device):
    if isinstance(x, np.ndarray):
        x = th.from_numpy(np.ascontiguousarray(x))
    if not th.is_tensor(x):
        x = th.as_tensor(x)
    return x.to(device=device, dtype=th.float32, non_blocking=True).contiguous()


class SharedMAC:
    """All agents share one RNNQNet; the agent id travels as a one-hot input."""

    def __init__(self, args):
        self.n_actions = args.n_actions
        self.n_agents = args.n_agents
        self.state_shape = args.state_shape
        self.obs_shape = args.obs_shape
        self.args = args
        if not (args.last_action and args.reuse_network):
            raise NotImplementedError("libmarl_b200 implements the default last_action=True, reuse_network=True input")
        self._build_agents(self._get_input_shape())
        self.hidden_states = None   # (n_episodes, n_agents, hidden_dim)

    # ---- construction -----------------------------------------------------------------------------
    def _get_input_shape(self):
        return self.obs_shape + self.n_actions + self.n_agents    # share_params.py:114-123

    def _build_agents(self, input_shape):
        self.agent = RNNQNet(input_shape, self.args)

    @property
    def device(self):
        return self.agent.fc1.weight.device

    def cuda(self):
        self.agent.cuda()

    def parameters(self):
        return self.agent.parameters()

    def load_state(self, other_mac):
        self.agent.load_state_dict(other_mac.agent.state_dict())

    def save_models(self, path):
        th.save(self.agent.state_dict(), path)

    def load_models(self, path):
        self.agent.load_state_dict(th.load(path, map_location=self.device))

    # ---- hidden state -------------------------------------------------------------------------------
    def init_hidden(self, episode_num):
        self.hidden_states = th.zeros((episode_num, self.n_agents, self.args.rnn_hidden_dim), device=self.device)

    # ---- episode unrolls ------------------------------------------------------------------------------
    def _unroll(self, obs, onehot, max_episode_len, shift):
        dev = self.device
        if dev.type != "cuda":
            raise L.MarlLibraryError("SharedMAC needs a CUDA device: marl_b200 has no CPU path")
        obs = _as_f32_cuda(obs, dev)[:, :max_episode_len].contiguous()
        onehot = _as_f32_cuda(onehot, dev)[:, :max_episode_len].contiguous()
        h0 = self.hidden_states.reshape(-1, self.args.rnn_hidden_dim).to(dev).contiguous()
        q, hidden, h_last = self.agent.unroll(obs, onehot, h0, shift)
        self.hidden_states = h_last            # [B*N, H], as left behind by the reference's loop
        return q, hidden

    def get_current_q_values(self, batch, max_episode_len):
        """Q-values of every transition's current observation (share_params.py:125-146):
        step t consumes [o_t | u_onehot_{t-1} (zeros at t=0) | agent id]."""
        return self._unroll(batch["o"], batch["u_onehot"], max_episode_len, shift=1)

    def get_next_q_values(self, batch, max_episode_len):
        """Q-values of every transition's next observation (share_params.py:148-168):
        step t consumes [o_next_t | u_onehot_t | agent id]."""
        return self._unroll(batch["o_next"], batch["u_onehot"], max_episode_len, shift=0)

    def forward(self, ep_batch, t):
        """pymarl-style alias named by the north star: Q-values of all agents at step t of the
        current-observation stream (one step of get_current_q_values), returns [B, N, A]."""
        sl = slice(max(t - 1, 0), t + 1)
        batch = {"o": ep_batch["o"][:, sl], "u_onehot": ep_batch["u_onehot"][:, sl]}
        if t == 0:
            q, _ = self._unroll(batch["o"], batch["u_onehot"], 1, shift=1)
            return q[:, 0]
        # step t needs u_onehot[t-1]: feed it unshifted alongside o_t
        dev = self.device
        obs = _as_f32_cuda(ep_batch["o"], dev)[:, t:t + 1].contiguous()
        oh = _as_f32_cuda(ep_batch["u_onehot"], dev)[:, t - 1:t].contiguous()
        h0 = self.hidden_states.reshape(-1, self.args.rnn_hidden_dim).to(dev).contiguous()
        q, _, h_last = self.agent.unroll(obs, oh, h0, 0)
        self.hidden_states = h_last
        return q[:, 0]

    # ---- acting ---------------------------------------------------------------------------------------
    def choose_action(self, obs, last_action, agent_num, avail_actions, epsilon, evaluate=False):
        """Epsilon-greedy action of ONE agent (share_params.py:37-72); RNG order: one
        np.random.uniform(), then np.random.choice only when exploring."""
        inputs = np.asarray(obs, dtype=np.float64).copy()
        avail_actions_ind = np.nonzero(avail_actions)[0]
        agent_id = np.zeros(self.n_agents)
        agent_id[agent_num] = 1.
        inputs = np.hstack((inputs, last_action, agent_id))
        dev = self.device
        hidden_state = self.hidden_states[:, agent_num, :]
        inputs = th.tensor(inputs, dtype=th.float32).unsqueeze(0).to(dev)
        avail = th.tensor(np.asarray(avail_actions), dtype=th.float32).unsqueeze(0).to(dev)
        with th.no_grad():
            q_value, h = self.agent(inputs, hidden_state.to(dev))
            self.hidden_states[:, agent_num, :] = h
            q_value[avail == 0.0] = -float("inf")
        if np.random.uniform() < epsilon:
            action = np.random.choice(avail_actions_ind)
        else:
            action = th.argmax(q_value)
        return action
