"""QPLEX duplex-dueling mixer on sm_100a kernels, drop-in for ``DMAQ_SI_Weight`` / ``DMAQer``
(reference ``network/mixer.py:85-288``): same constructor, state_dict keys
(``hyper_w_final.{0,2}``, ``V.{0,2}``, ``si_weight.{key,agents,action}_extractors.k.{0,2,4}``),
same ``forward(agent_qs, states, actions=None, max_q_i=None, is_v=False)``.

The 2 + 3*num_kernel small MLPs are stored layer-concatenated in one flat buffer so that a mixer
call is six GEMM launches plus one warp-per-sample kernel (``marl_qplex_fwd`` / ``_bwd``).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from .. import _lib as L
from ..flat import FlatBuffer, FlatPackedMixin


def _mlp3(n_in, hid, n_out):
    return nn.Sequential(nn.Linear(n_in, hid), nn.ReLU(), nn.Linear(hid, hid), nn.ReLU(), nn.Linear(hid, n_out))


class DMAQ_SI_Weight(nn.Module):
    """lambda_i(tau, a): num_kernel heads of |key(s)| * sigmoid(agents(s)) * sigmoid(action([s, a]))."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.n_agents, self.n_actions = args.n_agents, args.n_actions
        self.state_dim = int(np.prod(args.state_shape))
        self.action_dim = args.n_agents * args.n_actions
        self.state_action_dim = self.state_dim + self.action_dim
        self.num_kernel = args.num_kernel
        self.layers = int(getattr(args, "adv_hypernet_layers", 1))      # the reference's getattr default (mixer.py:115)
        if self.layers not in (1, 2, 3):
            raise Exception("Error setting number of adv hypernet layers.")       # mixer.py:145
        ae = args.adv_hypernet_embed

        def ext(n_in, n_out):
            if self.layers == 1:
                return nn.Linear(n_in, n_out)
            if self.layers == 2:
                return nn.Sequential(nn.Linear(n_in, ae), nn.ReLU(), nn.Linear(ae, n_out))
            return _mlp3(n_in, ae, n_out)

        self.key_extractors, self.agents_extractors, self.action_extractors = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        for _ in range(self.num_kernel):
            self.key_extractors.append(ext(self.state_dim, 1))
            self.agents_extractors.append(ext(self.state_dim, self.n_agents))
            self.action_extractors.append(ext(self.state_action_dim, self.n_agents))


def qplex_flat_order(K, layers=3):
    """Parameter names in flat order; entries after the first of each group are packed tight.  `layers` =
    adv_hypernet_layers: the extractors' output layer is ``.4`` (3 layers), ``.2`` (2 layers) or the bare Linear (1 layer,
    where the key and agents extractors form ONE [K + K*N, S] matrix and there is no extractor hidden layer at all)."""
    ks = range(K)
    si = "si_weight."
    if layers == 1:
        return {
            "w1s": ["hyper_w_final.0.weight", "V.0.weight"],
            "b1s": ["hyper_w_final.0.bias", "V.0.bias"],
            "w3k": [f"{si}key_extractors.{k}.weight" for k in ks] + [f"{si}agents_extractors.{k}.weight" for k in ks],
            "b3k": [f"{si}key_extractors.{k}.bias" for k in ks] + [f"{si}agents_extractors.{k}.bias" for k in ks],
            "w3n": [f"{si}action_extractors.{k}.weight" for k in ks],
            "b3n": [f"{si}action_extractors.{k}.bias" for k in ks],
            "wfv": ["hyper_w_final.2.weight", "V.2.weight"],
            "bfv": ["hyper_w_final.2.bias", "V.2.bias"],
        }
    out = 2 * (layers - 1)
    groups = {
        "w1s": ["hyper_w_final.0.weight", "V.0.weight"] + [f"{si}key_extractors.{k}.0.weight" for k in ks]
               + [f"{si}agents_extractors.{k}.0.weight" for k in ks],
        "b1s": ["hyper_w_final.0.bias", "V.0.bias"] + [f"{si}key_extractors.{k}.0.bias" for k in ks]
               + [f"{si}agents_extractors.{k}.0.bias" for k in ks],
        "w1a": [f"{si}action_extractors.{k}.0.weight" for k in ks],
        "b1a": [f"{si}action_extractors.{k}.0.bias" for k in ks],
    }
    if layers == 3:
        groups["w2"] = [f"{si}{t}_extractors.{k}.2.weight" for t in ("key", "agents", "action") for k in ks]
        groups["b2"] = [f"{si}{t}_extractors.{k}.2.bias" for t in ("key", "agents", "action") for k in ks]
    groups.update({
        "w3k": [f"{si}key_extractors.{k}.{out}.weight" for k in ks],
        "b3k": [f"{si}key_extractors.{k}.{out}.bias" for k in ks],
        "w3n": [f"{si}{t}_extractors.{k}.{out}.weight" for t in ("agents", "action") for k in ks],
        "b3n": [f"{si}{t}_extractors.{k}.{out}.bias" for t in ("agents", "action") for k in ks],
        "wfv": ["hyper_w_final.2.weight", "V.2.weight"],
        "bfv": ["hyper_w_final.2.bias", "V.2.bias"],
    })
    return groups


def qplex_struct(addr_of, K, cls=L.QplexParams, layers=3):
    """addr_of(name) -> device address; fills the group pointers with the address of each group's head (groups a
    layer count does not have stay NULL)."""
    s = cls()
    for field, names in qplex_flat_order(K, layers).items():
        setattr(s, field, addr_of(names[0]))
    return s


def _layers(args):
    return int(getattr(args, "adv_hypernet_layers", 1))


def qplex_dims(args):
    return L.QplexDims(args.n_agents, args.n_actions, int(np.prod(args.state_shape)), args.hypernet_embed,
                       args.adv_hypernet_embed, args.num_kernel, int(bool(args.weighted_head)), int(bool(args.is_minus_one)),
                       _layers(args))


def qplex_workspace(M, args, device):
    N, K, he, ae = args.n_agents, args.num_kernel, args.hypernet_embed, args.adv_hypernet_embed
    ext = 0 if _layers(args) == 1 else K * ae
    f = lambda w: torch.empty(M, max(w, 1), dtype=torch.float32, device=device)
    return dict(h1=f(2 * he + 3 * ext), h2=f(3 * ext if _layers(args) == 3 else 1), o3=f(K + 2 * K * N), wv=f(2 * N))


def ws_struct(ws):
    s = L.QplexWs()
    s.h1, s.h2, s.o3, s.wv = (ws[k].data_ptr() for k in ("h1", "h2", "o3", "wv"))
    return s


class _QplexFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, s, actions, max_q, is_v, module, *params):
        M = q.shape[0]
        args = module.args
        ws = qplex_workspace(M, args, q.device)
        out = torch.empty(M, dtype=torch.float32, device=q.device)
        d = qplex_dims(args)
        names = module.flat_names()
        p = qplex_struct(dict(zip(names, (t.data_ptr() for t in params))).__getitem__, args.num_kernel, layers=_layers(args))
        wss = ws_struct(ws)
        if is_v:
            L.call("marl_qplex_fwd", M, C.byref(d), C.byref(p), q.data_ptr(), s.data_ptr(), None, None, C.byref(wss),
                   out.data_ptr(), None, None, L.stream_ptr())
        else:
            L.call("marl_qplex_fwd", M, C.byref(d), C.byref(p), q.data_ptr(), s.data_ptr(), actions.data_ptr(),
                   max_q.data_ptr(), C.byref(wss), None, out.data_ptr(), None, L.stream_ptr())
        ctx.save_for_backward(q, s, actions if actions is not None else q, max_q if max_q is not None else q, *params)
        ctx.ws, ctx.is_v, ctx.module = ws, is_v, module
        return out

    @staticmethod
    def backward(ctx, dout):
        q, s, actions, max_q, *params = ctx.saved_tensors
        module, is_v = ctx.module, ctx.is_v
        args = module.args
        M = q.shape[0]
        dev = q.device
        names = module.flat_names()
        tight = module.flat_tight()
        # gradient buffer with the parameters' own (tight) layout
        offs, off = {}, 0
        for n, t in zip(names, params):
            if n not in tight:
                off = (off + 3) // 4 * 4
            offs[n] = off
            off += t.numel()
        gbuf = torch.zeros(off + 4, dtype=torch.float32, device=dev)
        views = [gbuf[offs[n]:offs[n] + t.numel()].view(t.shape) for n, t in zip(names, params)]
        g = qplex_struct(lambda n: gbuf.data_ptr() + 4 * offs[n], args.num_kernel, L.QplexGrads, layers=_layers(args))
        p = qplex_struct(dict(zip(names, (t.data_ptr() for t in params))).__getitem__, args.num_kernel, layers=_layers(args))
        dws = qplex_workspace(M, args, dev)
        for t in dws.values():
            t.zero_()
        dq = torch.zeros(M, args.n_agents, dtype=torch.float32, device=dev)
        d = qplex_dims(args)
        dout = dout.contiguous()
        wss, dwss = ws_struct(ctx.ws), ws_struct(dws)
        if is_v:
            L.call("marl_qplex_bwd", M, C.byref(d), C.byref(p), q.data_ptr(), s.data_ptr(), None, None, C.byref(wss),
                   dout.data_ptr(), None, C.byref(dwss), dq.data_ptr(), C.byref(g), L.stream_ptr())
            return (dq, None, None, None, None, None, *views)
        L.call("marl_qplex_bwd", M, C.byref(d), C.byref(p), q.data_ptr(), s.data_ptr(), actions.data_ptr(), max_q.data_ptr(),
               C.byref(wss), None, dout.data_ptr(), C.byref(dwss), dq.data_ptr(), C.byref(g), L.stream_ptr())
        return (None, None, None, None, None, None, *views)      # adv is detached: no gradient to agent_qs


class DMAQer(FlatPackedMixin, nn.Module):
    """Q_tot = sum_i (w_i Q_i + v_i) + sum_i (lambda_i - 1) A_i   (transformation + dueling mixing)."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.n_agents, self.n_actions = args.n_agents, args.n_actions
        self.state_dim = int(np.prod(args.state_shape))
        self.action_dim = args.n_agents * args.n_actions
        self.state_action_dim = self.state_dim + self.action_dim + 1
        self.embed_dim = args.mixing_embed_dim
        he = args.hypernet_embed
        if args.n_agents > 32:
            raise NotImplementedError("the QPLEX mixing kernel maps agents to warp lanes (n_agents <= 32)")
        self.hyper_w_final = nn.Sequential(nn.Linear(self.state_dim, he), nn.ReLU(), nn.Linear(he, self.n_agents))
        self.V = nn.Sequential(nn.Linear(self.state_dim, he), nn.ReLU(), nn.Linear(he, self.n_agents))
        self.si_weight = DMAQ_SI_Weight(args)
        self._flat = None
        self._pack(torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu"))

    # ---- flat storage ------------------------------------------------------------------------------
    def flat_names(self):
        return [n for names in qplex_flat_order(self.args.num_kernel, _layers(self.args)).values() for n in names]

    def flat_tight(self):
        return {n for names in qplex_flat_order(self.args.num_kernel, _layers(self.args)).values() for n in names[1:]}

    def flat_named_parameters(self):
        table = dict(self.named_parameters())
        tight = self.flat_tight()
        return [(n, table[n], "tight") if n in tight else (n, table[n]) for n in self.flat_names()]

    def _pack(self, device):
        self._flat = FlatBuffer(self.flat_named_parameters(), device=device, with_grad=False)

    def adopt(self, flat):
        self._flat = flat

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._pack(self.hyper_w_final[0].weight.device)
        return out

    # ---- reference surface ----------------------------------------------------------------------------
    def forward(self, agent_qs, states, actions=None, max_q_i=None, is_v=False):
        bs = agent_qs.size(0)
        self.ensure_packed()      # layer groups are read as one matrix each: re-pack a copied module first
        N = self.n_agents
        q = L.require_cuda(agent_qs, "agent_qs").reshape(-1, N).to(torch.float32).contiguous()
        s = L.require_cuda(states, "states").reshape(-1, self.state_dim).to(torch.float32).contiguous()
        table = dict(self.named_parameters())
        params = [table[n] for n in self.flat_names()]
        if is_v:
            y = _QplexFn.apply(q, s, None, None, True, self, *params)
        else:
            a = actions.reshape(-1, self.action_dim).to(torch.float32).contiguous()
            mq = max_q_i.reshape(-1, N).to(torch.float32).contiguous()
            y = _QplexFn.apply(q, s, a, mq, False, self, *params)
        return y.view(bs, -1, 1)
