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


def choose_action_sequence(agent_params, obs, avail, eps, n_agents, n_actions, hidden=64):
    """``SharedMAC.choose_action`` (share_params.py:37-72) called agent by agent for every step of ONE episode, the way
    rollout.py:60-76 drives it: obs [T, N, O], avail [T, N, A], eps [T]; numpy's global RNG in the reference's order
    (one uniform per agent, one choice only when exploring).  Returns actions [T, N] and the carried hidden state
    after every step [T, N, hidden]."""
    p = {k: v.detach() for k, v in agent_params.items()}
    T = obs.shape[0]
    h = torch.zeros(1, n_agents, hidden)
    last = np.zeros((n_agents, n_actions))
    actions = np.zeros((T, n_agents), dtype=np.int64)
    hid = np.zeros((T, n_agents, hidden), dtype=np.float32)
    with torch.no_grad():
        for t in range(T):
            for a in range(n_agents):
                agent_id = np.zeros(n_agents)
                agent_id[a] = 1.0
                x = torch.tensor(np.hstack((obs[t, a], last[a], agent_id)), dtype=torch.float32).unsqueeze(0)
                q, hn = MO.agent_step(p, x, h[:, a, :])
                h[:, a, :] = hn
                q[torch.tensor(avail[t, a], dtype=torch.float32).unsqueeze(0) == 0.0] = -float("inf")
                if np.random.uniform() < eps[t]:
                    act = int(np.random.choice(np.nonzero(avail[t, a])[0]))
                else:
                    act = int(torch.argmax(q))
                actions[t, a] = act
                last[a] = np.eye(n_actions)[act]
            hid[t] = h[0].numpy()
    return actions, hid


def rollout_multistep(agent_params, env, n_episodes, epsilon, anneal_epsilon, min_epsilon, anneal_scale="step", evaluate=False,
                      draws=None, hidden=64):
    """``RolloutWorker.generate_episodes`` (rollout.py:30-173) on a multi-step SMAC-style host environment: per-step
    epsilon (``:47-49, 103-104``), availability masking in ``choose_action``, the trailing observation / state /
    availability read after the last step (``:106-120``), zero padding to ``episode_limit`` with ``padded = 1`` and
    ``terminated = 1`` (``:122-133``), the episode layout (``:135-149``).

    ``draws=None`` consumes numpy's global RNG in the reference's order.  ``draws=(explore_u, choice_u)`` (arrays
    [n_episodes, episode_limit, n_agents] of uniforms in [0, 1)) is the batched worker's RNG contract: agent a of
    episode e explores at step t iff explore_u[e, t, a] < epsilon_t and then takes the floor(choice_u * n_available)-th
    available action; every episode starts from the same epsilon (the instances run side by side)."""
    info = env.get_env_info()
    N, A, O, S, T = (info[k] for k in ("n_agents", "n_actions", "obs_shape", "state_shape", "episode_limit"))
    p = {k: v.detach() for k, v in agent_params.items()}
    keys = ("o", "s", "u", "r", "avail_u", "o_next", "s_next", "avail_u_next", "u_onehot", "padded", "terminated")
    out = {k: [] for k in keys}
    rewards, steps_tot, cur = [], 0, epsilon
    with torch.no_grad():
        for e in range(n_episodes):
            env.reset()
            h = torch.zeros(1, N, hidden)
            eps = 0 if evaluate else (cur if draws is None else epsilon)
            if anneal_scale == "episode":
                eps = eps - anneal_epsilon if eps > min_epsilon else eps
            last = np.zeros((N, A))
            o, u, r, s, avail_u, u_onehot, terminate, padded = [], [], [], [], [], [], [], []
            terminated, step, ep_reward = False, 0, 0.0
            while not terminated and step < T:
                obs, state, avail = env.get_obs(), env.get_state(), env.get_avail_actions()
                actions, onehots = [], []
                for a in range(N):
                    agent_id = np.zeros(N)
                    agent_id[a] = 1.0
                    x = torch.tensor(np.hstack((obs[a], last[a], agent_id)), dtype=torch.float32).unsqueeze(0)
                    q, hn = MO.agent_step(p, x, h[:, a, :])
                    h[:, a, :] = hn
                    av = np.asarray(avail[a])
                    q[torch.tensor(av, dtype=torch.float32).unsqueeze(0) == 0.0] = -float("inf")
                    idx = np.nonzero(av)[0]
                    if draws is None:
                        act = int(np.random.choice(idx)) if np.random.uniform() < eps else int(torch.argmax(q))
                    else:
                        eu, cu = draws[0][e, step, a], draws[1][e, step, a]
                        act = int(idx[min(int(cu * len(idx)), len(idx) - 1)]) if eu < eps else int(torch.argmax(q))
                    actions.append(act)
                    onehots.append(np.eye(A)[act])
                    last[a] = np.eye(A)[act]
                reward, terminated, _ = env.step(actions)
                o.append(obs); s.append(state); u.append(np.reshape(actions, [N, 1])); u_onehot.append(onehots)
                avail_u.append(avail); r.append([reward]); terminate.append([terminated]); padded.append([0.0])
                ep_reward += reward
                step += 1
                if anneal_scale == "step":
                    eps = eps - anneal_epsilon if eps > min_epsilon else eps
            o.append(env.get_obs()); s.append(env.get_state())
            o_next, s_next, o, s = o[1:], s[1:], o[:-1], s[:-1]
            avail_u.append([env.get_avail_agent_actions(a) for a in range(N)])
            avail_u_next, avail_u = avail_u[1:], avail_u[:-1]
            for _ in range(step, T):
                o.append(np.zeros((N, O))); u.append(np.zeros([N, 1])); s.append(np.zeros(S)); r.append([0.0])
                o_next.append(np.zeros((N, O))); s_next.append(np.zeros(S)); u_onehot.append(np.zeros((N, A)))
                avail_u.append(np.zeros((N, A))); avail_u_next.append(np.zeros((N, A))); padded.append([1.0]); terminate.append([1.0])
            ep = dict(o=o, s=s, u=u, r=r, avail_u=avail_u, o_next=o_next, s_next=s_next, avail_u_next=avail_u_next,
                      u_onehot=u_onehot, padded=padded, terminated=terminate)
            for k in keys:
                out[k].append(np.array(ep[k], dtype=np.float64))
            rewards.append(ep_reward)
            steps_tot += step
            if not evaluate and draws is None:
                cur = eps
    return {k: np.stack(v) for k, v in out.items()}, rewards, steps_tot
