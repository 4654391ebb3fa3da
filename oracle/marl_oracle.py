"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the value-factorisation learner step.

A functional (stateless-module) restatement, in plain PyTorch on the CPU, of the
reference hot path.  Parameters travel as ``{group: {state_dict_key: tensor}}`` so
that weights can be exchanged with the reference (goldens) and with the CUDA
product (parity tests) by name.  Gradients come from autograd, so the oracle checks
the hand-derived backward kernels of the product independently.

Pinned by ``tests/golden/*.npz`` (made by ``oracle/make_golden.py`` from the
unmodified reference imported from /root/reference; see tests/test_oracle_golden.py).

Reference lines restated (paths relative to /root/reference):
  * agent   network/q_network.py:16-21, controller/share_params.py:84-168
  * VDN     network/mixer.py:15-16          * QMIX  network/mixer.py:57-80
  * QPLEX   network/mixer.py:149-171,211-288
  * QTRAN   network/mixer.py:378-388,411-418
  * learner algorithm/q_learner.py:49-179, algorithm/qtran_learner.py:71-200
  * clip / RMSprop / Adam: torch/nn/utils/clip_grad.py, torch/optim/{rmsprop,adam}.py
"""
from __future__ import annotations

import copy
import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

NEG_BIG = -9999999.0      # q_learner.py:105,112,126 ; qtran_learner.py:106
NEG_QTRAN_EVAL = -999999.0  # qtran_learner.py:105

DEFAULTS = dict(
    separated=False,   # SeparatedMAC (share_params.py:389-610): one network per agent
     # common/arguments.py:86-147 (get_mixer_args) + :31-34
    rnn_hidden_dim=64, qmix_hidden_dim=32, two_hyper_layers=False, hyper_hidden_dim=64,
    qtran_hidden_dim=64, lr=5e-4, target_update_cycle=200, lambda_opt=1, lambda_nopt=1,
    grad_norm_clip=10, adv_hypernet_embed=64, num_kernel=10, adv_hypernet_layers=3,
    weighted_head=True, hypernet_embed=64, is_minus_one=True, double_q=True,
    gamma=0.99, optimizer="RMS", last_action=True, reuse_network=True, alg="qmix",
)


def make_cfg(**kw):
    d = dict(DEFAULTS)
    d.update(kw)
    return SimpleNamespace(**d)


# --------------------------------------------------------------------------------------
# parameter construction (same shapes / key names / default initialisers as the
# reference's nn.Linear / nn.GRUCell modules, created in the same order)
# --------------------------------------------------------------------------------------
def _linear(sd, name, n_in, n_out):
    m = torch.nn.Linear(n_in, n_out)
    sd[name + ".weight"] = m.weight.detach().clone()
    sd[name + ".bias"] = m.bias.detach().clone()


def init_agent(cfg):
    """network/q_network.py:8-14 ; input width per share_params.py:114-123."""
    sd = {}
    n_in = cfg.obs_shape + (cfg.n_actions if getattr(cfg, "last_action", True) else 0) + \
        (cfg.n_agents if getattr(cfg, "reuse_network", True) else 0)          # share_params.py:114-123 / :497-506
    H = cfg.rnn_hidden_dim
    _linear(sd, "fc1", n_in, H)
    g = torch.nn.GRUCell(H, H)
    for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
        sd["rnn." + k] = getattr(g, k).detach().clone()
    _linear(sd, "fc2", H, cfg.n_actions)
    return sd


def init_mixer(cfg, alg=None):
    alg = alg or cfg.alg
    sd = {}
    S, N, A, H = cfg.state_shape, cfg.n_agents, cfg.n_actions, cfg.rnn_hidden_dim
    if alg == "vdn":
        return sd
    if alg == "qmix":  # mixer.py:45-55
        E = cfg.qmix_hidden_dim
        if getattr(cfg, "two_hyper_layers", False):  # mixer.py:36-43
            hh = cfg.hyper_hidden_dim
            _linear(sd, "hyper_w1.0", S, hh)
            _linear(sd, "hyper_w1.2", hh, N * E)
            _linear(sd, "hyper_w2.0", S, hh)
            _linear(sd, "hyper_w2.2", hh, E)
        else:
            _linear(sd, "hyper_w1", S, N * E)
            _linear(sd, "hyper_w2", S, E)
        _linear(sd, "hyper_b1", S, E)
        _linear(sd, "hyper_b2.0", S, E)
        _linear(sd, "hyper_b2.2", E, 1)
        return sd
    if alg == "qplex":  # mixer.py:200-208, 110-145 (3-layer extractors)
        he, ae = cfg.hypernet_embed, cfg.adv_hypernet_embed
        _linear(sd, "hyper_w_final.0", S, he)
        _linear(sd, "hyper_w_final.2", he, N)
        _linear(sd, "V.0", S, he)
        _linear(sd, "V.2", he, N)
        nl = cfg.adv_hypernet_layers          # mixer.py:115-145: 1, 2 or 3 Linear layers per extractor
        assert nl in (1, 2, 3)
        for k in range(cfg.num_kernel):
            for name, n_in, n_out in (("key", S, 1), ("agents", S, N), ("action", S + N * A, N)):
                pre = f"si_weight.{name}_extractors.{k}"
                if nl == 1:
                    _linear(sd, pre, n_in, n_out)
                else:
                    widths = [n_in] + [ae] * (nl - 1) + [n_out]
                    for i in range(nl):
                        _linear(sd, f"{pre}.{2 * i}", widths[i], widths[i + 1])
        return sd
    if alg == "qtran_base":  # mixer.py:364-375
        ae, qh = H + A, cfg.qtran_hidden_dim
        _linear(sd, "hidden_action_encoding.0", ae, ae)
        _linear(sd, "hidden_action_encoding.2", ae, ae)
        _linear(sd, "q.0", S + A + H, qh)
        _linear(sd, "q.2", qh, qh)
        _linear(sd, "q.4", qh, 1)
        return sd
    raise ValueError("Mixer {} not recognised.".format(alg))


def init_qtran_v(cfg):  # mixer.py:398-409
    sd = {}
    H, qh = cfg.rnn_hidden_dim, cfg.qtran_hidden_dim
    _linear(sd, "hidden_encoding.0", H, H)
    _linear(sd, "hidden_encoding.2", H, H)
    _linear(sd, "v.0", cfg.state_shape + H, qh)
    _linear(sd, "v.2", qh, qh)
    _linear(sd, "v.4", qh, 1)
    return sd


# --------------------------------------------------------------------------------------
# networks
# --------------------------------------------------------------------------------------
def agent_step(p, x, h):
    """q_network.py:16-21 : fc1 -> relu -> GRUCell -> fc2 (gate order r,z,n)."""
    x = F.relu(F.linear(x, p["fc1.weight"], p["fc1.bias"]))
    h = torch.gru_cell(x, h, p["rnn.weight_ih"], p["rnn.weight_hh"], p["rnn.bias_ih"], p["rnn.bias_hh"])
    return F.linear(h, p["fc2.weight"], p["fc2.bias"]), h


def gru_cell_explicit(x, h, w_ih, w_hh, b_ih, b_hh):
    """torch/nn/modules/rnn.py GRUCell equations written out (cross-check of gru_cell)."""
    gi = F.linear(x, w_ih, b_ih)
    gh = F.linear(h, w_hh, b_hh)
    H = h.shape[-1]
    r = torch.sigmoid(gi[:, :H] + gh[:, :H])
    z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    return (1 - z) * n + z * h


def unroll(p, obs, last_onehot, h, cfg):
    """share_params.py:125-168 : T-step unroll; rows are (b, n) with n minor.

    obs [B,T,N,O]; last_onehot [B,T,N,A] is the one-hot fed at each step (already
    shifted for the current-obs stream, share_params.py:96-101).  Returns
    q [B,T,N,A], hidden [B,T,N,H] (h after each step), final h [B*N,H].
    """
    B, T, N, _ = obs.shape
    eye = torch.eye(N, dtype=obs.dtype).unsqueeze(0).expand(B, -1, -1)
    separated = isinstance(p, (list, tuple))        # SeparatedMAC: agent n's rows go through network n (share_params.py:527-533)
    qs, hs = [], []
    for t in range(T):
        parts = [obs[:, t].reshape(B * N, -1)]
        if getattr(cfg, "last_action", True):
            parts.append(last_onehot[:, t].reshape(B * N, -1))
        if getattr(cfg, "reuse_network", True):
            parts.append(eye.reshape(B * N, -1))
        x = torch.cat(parts, dim=1)
        if separated:
            xv, hv = x.view(B, N, -1), h.reshape(B, N, -1)
            outs = [agent_step(p[n], xv[:, n], hv[:, n]) for n in range(N)]
            q = torch.stack([o[0] for o in outs], dim=1).reshape(B * N, -1)
            h = torch.stack([o[1] for o in outs], dim=1).reshape(B * N, -1)
        else:
            q, h = agent_step(p, x, h.reshape(B * N, -1))
        qs.append(q.view(B, N, -1))
        hs.append(h.view(B, N, -1))
    return torch.stack(qs, 1), torch.stack(hs, 1), h


def shift_onehot(u_onehot):
    """share_params.py:96-99 : last action at t is u_onehot[t-1], zeros at t=0."""
    z = torch.zeros_like(u_onehot[:, :1])
    return torch.cat([z, u_onehot[:, :-1]], dim=1)


def vdn_mix(q):  # mixer.py:16
    return q.sum(dim=2, keepdim=True)


def qmix_mix(p, q, s, cfg):
    """mixer.py:57-80; hyper_w1 / hyper_w2 are one Linear (:44-47) or Linear-ReLU-Linear (two_hyper_layers, :36-43)."""
    B = q.shape[0]
    N, E = cfg.n_agents, cfg.qmix_hidden_dim
    q = q.reshape(-1, 1, N)
    s = s.reshape(-1, cfg.state_shape)

    def hyper(name):
        if name + ".0.weight" in p:
            return F.linear(F.relu(F.linear(s, p[name + ".0.weight"], p[name + ".0.bias"])), p[name + ".2.weight"], p[name + ".2.bias"])
        return F.linear(s, p[name + ".weight"], p[name + ".bias"])

    w1 = torch.abs(hyper("hyper_w1")).view(-1, N, E)
    b1 = F.linear(s, p["hyper_b1.weight"], p["hyper_b1.bias"]).view(-1, 1, E)
    hid = F.elu(torch.bmm(q, w1) + b1)
    w2 = torch.abs(hyper("hyper_w2")).view(-1, E, 1)
    b2 = F.linear(F.relu(F.linear(s, p["hyper_b2.0.weight"], p["hyper_b2.0.bias"])),
                  p["hyper_b2.2.weight"], p["hyper_b2.2.bias"]).view(-1, 1, 1)
    return (torch.bmm(hid, w2) + b2).view(B, -1, 1)


def _mlp(p, prefix, x, n_layers):
    for i in range(n_layers):
        x = F.linear(x, p[f"{prefix}.{2 * i}.weight"], p[f"{prefix}.{2 * i}.bias"])
        if i + 1 < n_layers:
            x = F.relu(x)
    return x


def qplex_lambda(p, s, actions, cfg):
    """DMAQ_SI_Weight.forward, mixer.py:149-171."""
    N = cfg.n_agents
    s = s.reshape(-1, cfg.state_shape)
    data = torch.cat([s, actions.reshape(-1, N * cfg.n_actions)], dim=1)
    tot = 0
    nl = cfg.adv_hypernet_layers

    def ext(name, k, x):
        pre = f"si_weight.{name}_extractors.{k}"
        return F.linear(x, p[pre + ".weight"], p[pre + ".bias"]) if nl == 1 else _mlp(p, pre, x, nl)

    for k in range(cfg.num_kernel):
        key = torch.abs(ext("key", k, s)).repeat(1, N) + 1e-10
        ag = torch.sigmoid(ext("agents", k, s))
        ac = torch.sigmoid(ext("action", k, data))
        tot = tot + key * ag * ac
    return tot


def qplex_mix(p, q, s, cfg, actions=None, max_q_i=None, is_v=False):
    """DMAQer.forward, mixer.py:257-288 (+calc_v :218-220, calc_adv :232-247)."""
    B, N = q.shape[0], cfg.n_agents
    sf = s.reshape(-1, cfg.state_shape)
    q = q.reshape(-1, N)
    w = torch.abs(_mlp(p, "hyper_w_final", sf, 2)).view(-1, N) + 1e-10
    v = _mlp(p, "V", sf, 2).view(-1, N)
    if cfg.weighted_head:
        q = w * q + v
    if is_v:
        return q.sum(dim=-1).view(B, -1, 1)
    mq = max_q_i.reshape(-1, N)
    if cfg.weighted_head:
        mq = w * mq + v
    adv = (q - mq).detach()
    lam = qplex_lambda(p, sf, actions, cfg).view(-1, N)
    y = (adv * (lam - 1.0)).sum(dim=1) if cfg.is_minus_one else (adv * lam).sum(dim=1)
    return y.view(B, -1, 1)


def qtran_q(p, s, hidden, actions, cfg):
    """QtranQBase.forward, mixer.py:378-388 -> [B*T, 1]."""
    B, T, N, _ = actions.shape
    ha = torch.cat([hidden, actions], dim=-1).reshape(B * T * N, -1)
    enc = _mlp(p, "hidden_action_encoding", ha, 2).reshape(B * T, N, -1).sum(dim=-2)
    return _mlp(p, "q", torch.cat([s.reshape(B * T, -1), enc], dim=-1), 3)


def qtran_v(p, s, hidden, cfg):
    """QtranV.forward, mixer.py:411-418 -> [B*T, 1]."""
    B, T, N, H = hidden.shape
    enc = _mlp(p, "hidden_encoding", hidden.reshape(-1, H), 2).reshape(B * T, N, -1).sum(dim=-2)
    return _mlp(p, "v", torch.cat([s.reshape(B * T, -1), enc], dim=-1), 3)


# --------------------------------------------------------------------------------------
# learner
# --------------------------------------------------------------------------------------
def max_episode_len(terminated, episode_limit):
    """q_learner.py:49-61 : 1 + the largest first-terminated index; limit if none."""
    L = 0
    term = np.asarray(terminated)
    for b in range(term.shape[0]):
        for t in range(episode_limit):
            if term[b, t, 0] == 1:
                if t + 1 >= L:
                    L = t + 1
                break
    return L if L > 0 else episode_limit


def to_tensors(batch, L, dtype):
    """q_learner.py:63-78 : slice to L, float cast, u -> int64 (truncation)."""
    out = {}
    for k, v in batch.items():
        v = np.asarray(v)[:, :L]
        out[k] = torch.tensor(v, dtype=torch.long) if k == "u" else torch.tensor(v, dtype=dtype)
    return out


class LearnerState:
    """All mutable state of one learner: eval/target params, optimiser moments."""

    def __init__(self, cfg, params=None, dtype=torch.float32):
        self.cfg, self.dtype = cfg, dtype
        if params is None:
            if getattr(cfg, "separated", False):
                params = {f"agent.{n}": init_agent(cfg) for n in range(cfg.n_agents)}
                params["mixer"] = init_mixer(cfg)
            else:
                params = {"agent": init_agent(cfg), "mixer": init_mixer(cfg)}
            if cfg.alg == "qtran_base":
                params["v"] = init_qtran_v(cfg)
                params["q_sum_mixer"] = init_mixer(cfg, "qmix")  # qtran_learner.py:37 (never used)
        self.params = {g: {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in sd.items()}
                       for g, sd in params.items()}
        self.target = {g: {k: v.detach().clone() for k, v in self.params[g].items()}
                       for g in self.params if g == "mixer" or g.startswith("agent")}
        self.opt = {}      # (group, key) -> dict of moment tensors
        self.steps = 0

    def flat_params(self):
        return [(g, k, v) for g, sd in self.params.items() for k, v in sd.items()]

    def sync_targets(self):  # q_learner.py:181-184
        for g in self.target:
            for k, v in self.params[g].items():
                self.target[g][k] = v.detach().clone()

    @staticmethod
    def agent_of(P):
        """The agent parameters of a group table: one dict (SharedMAC) or the per-agent list (SeparatedMAC)."""
        if "agent" in P:
            return P["agent"]
        return [P[f"agent.{n}"] for n in range(sum(1 for g in P if g.startswith("agent.")))]


def q_learner_forward(st, batch_np):
    """q_learner.py:70-168 up to the loss. Returns dict of named intermediates."""
    cfg, P, TP = st.cfg, st.params, st.target
    L = max_episode_len(batch_np["terminated"], cfg.episode_limit)
    b = to_tensors(batch_np, L, st.dtype)
    B, N = b["o"].shape[0], cfg.n_agents
    mask = 1 - b["padded"]
    h0 = torch.zeros(B * N, cfg.rnn_hidden_dim, dtype=st.dtype)
    q_evals, hid_evals, h_last = unroll(LearnerState.agent_of(P), b["o"], shift_onehot(b["u_onehot"]), h0, cfg)
    q_chosen = torch.gather(q_evals, 3, b["u"]).squeeze(3)
    with torch.no_grad():
        q_targets, _, _ = unroll(LearnerState.agent_of(TP), b["o_next"], b["u_onehot"], h0, cfg)
        q_targets[b["avail_u_next"] == 0.0] = NEG_BIG
        if cfg.double_q:
            # NB: hidden carried over from the current-obs unroll (no init_hidden, q_learner.py:110)
            q_en, _, _ = unroll(LearnerState.agent_of(P), b["o_next"], b["u_onehot"], h_last.detach(), cfg)
            q_en[b["avail_u_next"] == 0] = NEG_BIG
            a_star = torch.argmax(q_en, dim=3, keepdim=True)
            q_tc = torch.gather(q_targets, 3, a_star).squeeze(3)
        else:
            a_star, q_en = None, None
            q_tc = q_targets.max(dim=3)[0]
    if cfg.alg == "qplex":
        v_tot = qplex_mix(P["mixer"], q_chosen, b["s"], cfg, is_v=True)
        qd = q_evals.detach().clone()
        qd[b["avail_u"] == 0] = NEG_BIG
        max_q = qd.max(dim=3)[0]
        a_tot = qplex_mix(P["mixer"], q_chosen, b["s"], cfg, actions=b["u_onehot"], max_q_i=max_q)
        q_tot = v_tot + a_tot
        with torch.no_grad():
            if cfg.double_q:
                oh = torch.zeros_like(b["u_onehot"]).scatter_(3, a_star, 1)
                vt = qplex_mix(TP["mixer"], q_tc, b["s_next"], cfg, is_v=True)
                at = qplex_mix(TP["mixer"], q_tc, b["s_next"], cfg, actions=oh,
                               max_q_i=q_targets.max(dim=3)[0])
                q_tot_t = vt + at
            else:
                q_tot_t = qplex_mix(TP["mixer"], q_tc, b["s_next"], cfg, is_v=True)
    elif cfg.alg == "qmix":
        q_tot = qmix_mix(P["mixer"], q_chosen, b["s"], cfg)
        with torch.no_grad():
            q_tot_t = qmix_mix(TP["mixer"], q_tc, b["s_next"], cfg)
    elif cfg.alg == "vdn":
        q_tot, q_tot_t = vdn_mix(q_chosen), vdn_mix(q_tc)
    else:
        raise ValueError("Mixer {} not recognised.".format(cfg.alg))
    targets = b["r"] + cfg.gamma * q_tot_t * (1 - b["terminated"])
    td = targets.detach() - q_tot
    loss = ((mask * td) ** 2).sum() / mask.sum()
    return dict(loss=loss, L=L, q_evals=q_evals, hidden_evals=hid_evals, q_targets=q_targets,
                a_star=a_star, q_chosen=q_chosen, q_targets_chosen=q_tc, q_tot=q_tot,
                q_tot_target=q_tot_t, h_last=h_last, q_evals_next=q_en)


def qtran_forward(st, batch_np):
    """qtran_learner.py:73-152."""
    cfg, P, TP = st.cfg, st.params, st.target
    L = max_episode_len(batch_np["terminated"], cfg.episode_limit)
    b = to_tensors(batch_np, L, st.dtype)
    B, N = b["o"].shape[0], cfg.n_agents
    mask = 1 - b["padded"].squeeze(-1)
    h0 = torch.zeros(B * N, cfg.rnn_hidden_dim, dtype=st.dtype)
    q_ev, hid_ev, _ = unroll(LearnerState.agent_of(P), b["o"], shift_onehot(b["u_onehot"]), h0, cfg)
    with torch.no_grad():
        q_tg, hid_tg, _ = unroll(LearnerState.agent_of(TP), b["o_next"], b["u_onehot"], h0, cfg)
        q_tg[b["avail_u_next"] == 0.0] = NEG_BIG
        oh_t = torch.zeros_like(q_tg).scatter(-1, q_tg.argmax(dim=3, keepdim=True), 1)
    q_clone = q_ev.clone()
    q_clone[b["avail_u"] == 0.0] = NEG_QTRAN_EVAL
    opt_eval = q_clone.argmax(dim=3, keepdim=True)
    oh_e = torch.zeros_like(q_clone).scatter(-1, opt_eval, 1).detach()
    joint_q = qtran_q(P["mixer"], b["s"], hid_ev, b["u_onehot"], cfg).view(B, -1)
    with torch.no_grad():
        joint_q_t = qtran_q(TP["mixer"], b["s_next"], hid_tg, oh_t, cfg).view(B, -1)
    v = qtran_v(P["v"], b["s"], hid_ev, cfg).view(B, -1)
    y = b["r"].squeeze(-1) + cfg.gamma * joint_q_t * (1 - b["terminated"].squeeze(-1))
    l_td = (((joint_q - y.detach()) * mask) ** 2).sum() / mask.sum()
    q_sum_opt = q_clone.max(dim=-1)[0].sum(dim=-1)
    joint_q_hat = qtran_q(P["mixer"], b["s"], hid_ev, oh_e, cfg).view(B, -1)
    l_opt = (((q_sum_opt - joint_q_hat.detach() + v) * mask) ** 2).sum() / mask.sum()
    q_sum_nopt = torch.gather(q_ev, -1, b["u"]).squeeze(-1).sum(dim=-1)
    nopt = (q_sum_nopt - joint_q.detach() + v).clamp(max=0)
    l_nopt = ((nopt * mask) ** 2).sum() / mask.sum()
    loss = l_td + cfg.lambda_opt * l_opt + cfg.lambda_nopt * l_nopt
    return dict(loss=loss, L=L, q_evals=q_ev, hidden_evals=hid_ev, q_targets=q_tg,
                hidden_targets=hid_tg, opt_action_eval=opt_eval, joint_q=joint_q,
                joint_q_target=joint_q_t, joint_q_hat=joint_q_hat, v=v,
                l_td=l_td, l_opt=l_opt, l_nopt=l_nopt)


def forward(st, batch_np):
    return qtran_forward(st, batch_np) if st.cfg.alg == "qtran_base" else q_learner_forward(st, batch_np)


def clip_and_step(st, grads):
    """clip_grad_norm_ (L2, max_norm) then RMSprop / Adam with torch defaults.

    grads: list aligned with st.flat_params(); None entries are skipped exactly as
    torch skips ``p.grad is None`` (QTRAN q_sum_mixer, qtran_learner.py:37-38).
    """
    cfg = st.cfg
    live = [(g, k, p, gr) for (g, k, p), gr in zip(st.flat_params(), grads) if gr is not None]
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(gr) for *_, gr in live]))
    coef = torch.clamp(cfg.grad_norm_clip / (total + 1e-6), max=1.0)
    st.steps += 1
    with torch.no_grad():
        for g, k, p, gr in live:
            gr = gr * coef
            s = st.opt.setdefault((g, k), {})
            if cfg.optimizer == "RMS":      # torch/optim/rmsprop.py (alpha .99, eps 1e-8)
                sq = s.setdefault("square_avg", torch.zeros_like(p))
                sq.mul_(0.99).addcmul_(gr, gr, value=0.01)
                p.addcdiv_(gr, sq.sqrt().add_(1e-8), value=-cfg.lr)
            elif cfg.optimizer == "Adam":   # torch/optim/adam.py (betas .9/.999, eps 1e-8)
                m = s.setdefault("exp_avg", torch.zeros_like(p))
                v2 = s.setdefault("exp_avg_sq", torch.zeros_like(p))
                s["step"] = s.get("step", 0) + 1
                m.lerp_(gr, 0.1)
                v2.mul_(0.999).addcmul_(gr, gr, value=0.001)
                bc1, bc2 = 1 - 0.9 ** s["step"], 1 - 0.999 ** s["step"]
                p.addcdiv_(m, (v2.sqrt() / math.sqrt(bc2)).add_(1e-8), value=-cfg.lr / bc1)
            else:
                raise ValueError("optimizer {} not recognised.".format(cfg.optimizer))
    return float(total), [None if gr is None else gr * coef for gr in grads]


def train_step(st, batch_np, train_step_idx, want=None):
    """One full learner step (q_learner.py:68-179). Returns (loss float, info dict)."""
    out = forward(st, {k: np.array(v) for k, v in batch_np.items()})
    plist = [p for _, _, p in st.flat_params()]
    grads = torch.autograd.grad(out["loss"], plist, allow_unused=True)
    raw = list(grads)
    total, clipped = clip_and_step(st, raw)
    if train_step_idx > 0 and train_step_idx % st.cfg.target_update_cycle == 0:
        st.sync_targets()
    info = {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}
    info["grad_norm"] = total
    info["grads"] = {f"{g}.{k}": gr for (g, k, _), gr in zip(st.flat_params(), raw)}
    info["clipped_grads"] = {f"{g}.{k}": gr for (g, k, _), gr in zip(st.flat_params(), clipped)}
    return float(out["loss"].detach()), info


def clone_state(st):
    return copy.deepcopy(st)
