"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's sequential rollout on the matrix game.

Follows ``rollout.py:30-173`` (episode loop, epsilon schedule ``:47-49, 103-104, 166-167``, episode layout
``:122-149``) around ``controller/share_params.py:37-72`` (``choose_action``: batch-1 agent forward on
``[obs | last_action | agent_id]``, ``-inf`` mask, one ``np.random.uniform()`` and -- only when exploring -- one
``np.random.choice`` per agent) and ``env/single_state_matrix_game.py:27-55`` (reward = payoff[a0, a1], one-step
episodes, obs = state = 0, all actions available).  Only tests may import this module.
"""
import numpy as np
import torch

from . import marl_oracle as MO


def generate_episodes(agent_params, payoff, n_episodes, epsilon, anneal_epsilon, min_epsilon, anneal_scale="step",
                      evaluate=False, n_agents=2, n_actions=3, obs_shape=1, hidden=64):
    payoff = np.asarray(payoff, dtype=np.float64).reshape(3, 3)
    p = {k: v.detach() for k, v in agent_params.items()}
    actions = np.zeros((n_episodes, n_agents), dtype=np.int64)
    rewards = np.zeros(n_episodes, dtype=np.float64)
    cur = epsilon
    with torch.no_grad():
        for e in range(n_episodes):
            h = torch.zeros(1, n_agents, hidden)                       # mac.init_hidden(1), rollout.py:45
            eps = 0 if evaluate else cur
            if anneal_scale == "episode":
                eps = eps - anneal_epsilon if eps > min_epsilon else eps
            obs = [np.zeros(obs_shape) for _ in range(n_agents)]       # env.get_obs(), single_state_matrix_game.py:42-45
            last = np.zeros((n_agents, n_actions))
            for a in range(n_agents):
                agent_id = np.zeros(n_agents)
                agent_id[a] = 1.0
                x = torch.tensor(np.hstack((obs[a], last[a], agent_id)), dtype=torch.float32).unsqueeze(0)
                q, hn = MO.agent_step(p, x, h[:, a, :])
                h[:, a, :] = hn
                avail = np.ones(n_actions)
                q[torch.tensor(avail).unsqueeze(0) == 0.0] = -float("inf")
                if np.random.uniform() < eps:                          # share_params.py:67-70
                    act = int(np.random.choice(np.nonzero(avail)[0]))
                else:
                    act = int(torch.argmax(q))
                actions[e, a] = act
                last[a] = np.eye(n_actions)[act]
            rewards[e] = payoff[actions[e, 0], actions[e, 1]]
            if anneal_scale == "step":
                eps = eps - anneal_epsilon if eps > min_epsilon else eps
            if not evaluate:
                cur = eps
    return actions, rewards, cur
