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


class SeparatedMAC:
    """Every agent has its own RNNQNet, drop-in for ``controller/share_params.py:389-610`` (what ``runner.py:24-26``
    builds when ``reuse_network`` is False).

    Same surface and state semantics as the reference: ``agent`` is a LIST of networks, ``hidden_states`` is
    ``[episodes, n_agents, hidden]`` and is carried between successive ``get_*_q_values`` calls, the input of agent
    n at step t is ``[o_t[n] | u_onehot_{t-1}[n] (last_action) | eye(N)[n] (reuse_network)]`` (``:467-495``), and
    ``get_next_q_values`` returns the per-agent LIST of final hidden states as its second value like the reference
    does (``:590``).  Each agent's T-step unroll is one ``marl_agent_unroll_fwd`` call on its own [B, T, 1, I] input
    (N calls per unroll instead of N*T network forwards), differentiable through ``RNNQNet``'s autograd function.

    Two reference defects are NOT reproduced because they make the class unusable for training: ``load_state`` there
    calls ``other_mac.agent.state_dict()`` (an AttributeError on a list -> the first target sync raises), and its
    per-agent in-place ``hidden_states`` write-back breaks ``loss.backward()`` under current torch.  Here ``load_state``
    copies agent i to agent i (or one shared network into every agent), and the hidden state is re-assembled out of place.
    """

    def __init__(self, args):
        self.n_actions = args.n_actions
        self.n_agents = args.n_agents
        self.state_shape = args.state_shape
        self.obs_shape = args.obs_shape
        self.args = args
        self._build_agents(self._get_input_shape())
        self.hidden_states = None

    def _get_input_shape(self):                   # share_params.py:497-506
        input_shape = self.obs_shape
        if self.args.last_action:
            input_shape += self.n_actions
        if self.args.reuse_network:
            input_shape += self.n_agents
        return input_shape

    def _build_agents(self, input_shape):
        self.agent = [RNNQNet(input_shape, self.args) for _ in range(self.n_agents)]

    @property
    def device(self):
        return self.agent[0].fc1.weight.device

    def cuda(self):
        for a in self.agent:
            a.cuda()

    def parameters(self):
        params = []
        for a in self.agent:
            params += list(a.parameters())
        return params

    def load_state(self, other_mac):
        for i, a in enumerate(self.agent):
            src = other_mac.agent[i] if isinstance(other_mac.agent, (list, tuple)) else other_mac.agent
            a.load_state_dict(src.state_dict())

    def save_models(self, path):                  # share_params.py:603-605: every agent to the same path (the last one stays)
        for a in self.agent:
            th.save(a.state_dict(), path)

    def load_models(self, path):
        for a in self.agent:
            a.load_state_dict(th.load(path, map_location=self.device))

    def init_hidden(self, episode_num):
        self.hidden_states = th.zeros((episode_num, self.n_agents, self.args.rnn_hidden_dim), device=self.device)

    def _inputs(self, obs, onehot, L_, shift):
        """[B, L, N, I]: obs | last action (shifted by one step for the current-observation stream) | agent id."""
        dev = self.device
        obs = _as_f32_cuda(obs, dev)[:, :L_]
        parts = [obs]
        if self.args.last_action:
            oh = _as_f32_cuda(onehot, dev)[:, :L_]
            if shift:
                oh = th.cat([th.zeros_like(oh[:, :1]), oh[:, :-1]], dim=1)
            parts.append(oh)
        if self.args.reuse_network:
            B = obs.shape[0]
            parts.append(th.eye(self.n_agents, device=dev).view(1, 1, self.n_agents, self.n_agents).expand(B, L_, -1, -1))
        return th.cat(parts, dim=3)

    def _unroll(self, obs, onehot, max_episode_len, shift):
        if self.device.type != "cuda":
            raise L.MarlLibraryError("SeparatedMAC needs a CUDA device: marl_b200 has no CPU path")
        x = self._inputs(obs, onehot, max_episode_len, shift)
        B, L_, N, _ = x.shape
        h0 = self.hidden_states.reshape(B, N, -1).to(self.device)
        qs, hids, lasts = [], [], []
        for n, net in enumerate(self.agent):
            q, hid, h_last = net.unroll_full(x[:, :, n:n + 1].contiguous(), h0[:, n].contiguous())
            qs.append(q)
            hids.append(hid)
            lasts.append(h_last)
        self.hidden_states = th.stack(lasts, dim=1)                      # [B, N, H]
        return th.cat(qs, dim=2), th.cat(hids, dim=2), lasts

    def get_current_q_values(self, batch, max_episode_len):
        q, hid, _ = self._unroll(batch["o"], batch["u_onehot"], max_episode_len, shift=1)
        return q, hid

    def get_next_q_values(self, batch, max_episode_len):
        q, _, lasts = self._unroll(batch["o_next"], batch["u_onehot"], max_episode_len, shift=0)
        return q, lasts                                                  # the per-agent list, as share_params.py:590 returns

    def choose_action(self, obs, last_action, agent_num, avail_actions, epsilon, evaluate=False):
        """share_params.py:419-453 with the agent's own network; same RNG order as SharedMAC.choose_action."""
        inputs = np.asarray(obs, dtype=np.float64).copy()
        avail_actions_ind = np.nonzero(avail_actions)[0]
        agent_id = np.zeros(self.n_agents)
        agent_id[agent_num] = 1.
        if self.args.last_action:
            inputs = np.hstack((inputs, last_action))
        if self.args.reuse_network:
            inputs = np.hstack((inputs, agent_id))
        dev = self.device
        hidden_state = self.hidden_states[:, agent_num, :]
        inputs = th.tensor(inputs, dtype=th.float32).unsqueeze(0).to(dev)
        avail = th.tensor(np.asarray(avail_actions), dtype=th.float32).unsqueeze(0).to(dev)
        with th.no_grad():
            q_value, h = self.agent[agent_num](inputs, hidden_state.to(dev))
            self.hidden_states[:, agent_num, :] = h
            q_value[avail == 0.0] = -float("inf")
        if np.random.uniform() < epsilon:
            action = np.random.choice(avail_actions_ind)
        else:
            action = th.argmax(q_value)
        return action
