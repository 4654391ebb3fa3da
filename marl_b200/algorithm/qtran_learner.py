"""QTRAN-base learner on sm_100a kernels, drop-in for ``algorithm/qtran_learner.py:10-272``.

Same surface as the reference's ``QTRANLearner`` (``train(batch, train_step) -> float`` with
L_td + lambda_opt L_opt + lambda_nopt L_nopt, ``mixer`` / ``target_mixer`` / ``v`` / ``q_sum_mixer``,
checkpoints with ``_v_net_params.pkl``); the step is the fixed launch sequence

    2 agent unrolls (eval/o with saved gates, target/o_next)     qtran_learner.py:95-101
    greedy actions + one-hots                                    :103-114
    Q(s, h, u), Q_target(s', h', a*_target), Q(s, h, a*_eval), V(s, h)        :165-200
    three losses and their gradients                             :116-152
    backward through Q and V into d(hidden), BPTT with d(q) and d(hidden)     :155
    clip + optimiser over agent + mixer + V (q_sum_mixer never receives a gradient, :37-38,131-132)

``qtran_alt`` is constructible in the reference but its forward is broken there; it
raises here.
"""
from __future__ import annotations

import copy
import ctypes as C
import os

import torch as th

from .. import _lib as L
from ..network.mixer import QMixMixer
from ..network.q_network import agent_param_struct, AGENT_FLAT_ORDER
from ..network.qtran import QtranQBase, QtranV, net_struct, net_workspace, net_ws_struct
from .q_learner import QLearner

H = 64


class QTRANLearner(QLearner):
    def _build_extra(self, args):
        self.v = QtranV(args)                       # qtran_learner.py:34-35
        self.params += list(self.v.parameters())
        self.q_sum_mixer = QMixMixer(args)          # :37-38 -- handed to the optimiser by the reference, never used
        self.params += list(self.q_sum_mixer.parameters())

    def _make_mixer(self, args):
        if args.alg == 'qtran_base':
            return QtranQBase(args)
        if args.alg == 'qtran_alt':
            raise NotImplementedError("qtran_alt: the reference's own forward is inconsistent (mixer.py:320,349)")
        raise ValueError("Mixer {} not recognised.".format(args.alg))

    def _extra_groups(self):
        return [("v.", self.v)]

    def _workspace(self, B, Lq):
        fresh = (B, Lq) not in self._ws
        ws = super()._workspace(B, Lq)
        if fresh:
            a, dev = self.args, self._dev
            N, A, M = a.n_agents, a.n_actions, B * Lq
            f = lambda *shape: th.empty(*shape, dtype=th.float32, device=dev)
            D, qh = H + A, a.qtran_hidden_dim
            ws.update(oh_e=f(B, Lq, N, A), oh_t=f(B, Lq, N, A), opt_e=th.empty(B, Lq, N, dtype=th.int64, device=dev),
                      q_max=f(B, Lq, N), q_taken=f(B, Lq, N), jq=f(M), jq_t=f(M), jq_hat=f(M), vv=f(M), d_jq=f(M),
                      d_v=f(M), dhid=f(B * Lq * N, H), parts=th.zeros(3, dtype=th.float32, device=dev),
                      nq=net_workspace(M, N, D, qh, dev), nqt=net_workspace(M, N, D, qh, dev),
                      nqh=net_workspace(M, N, D, qh, dev), nv=net_workspace(M, N, H, qh, dev),
                      dnq=net_workspace(M, N, D, qh, dev), dnv=net_workspace(M, N, H, qh, dev))
        return ws

    def _net_addrs(self, flat, prefix, module, source=None):
        return [flat.ptr(prefix + n, source) for n, _ in module.named_parameters()]

    def _launch_forward_backward(self, bt, ws, B, Lq):
        a = self.args
        d = self._dims(B, Lq)
        sp = L.stream_ptr()
        fl, tfl = self._flat, self._tflat
        fl.grad_full.zero_()
        ws["parts"].zero_()
        pe = agent_param_struct({n: fl.ptr("agent." + n) for n in AGENT_FLAT_ORDER})
        pt = agent_param_struct({n: tfl.ptr("agent." + n) for n in AGENT_FLAT_ORDER})
        arr = (L.UnrollStream * 2)()
        for i, (obs, shift, params, gates) in enumerate(((bt["o"], 1, pe, ws["gates"].data_ptr()), (bt["o_next"], 0, pt, None))):
            s = arr[i]
            s.obs, s.onehot, s.shift_onehot, s.full_input = obs.data_ptr(), bt["u_onehot"].data_ptr(), shift, 0
            s.h0_from, s.h0, s.params = -1, None, params
            s.q, s.hidden, s.h_last = ws["q"][i].data_ptr(), ws["hidden"][i].data_ptr(), ws["h_last"][i].data_ptr()
            s.x, s.gi, s.gates = ws["x"][i].data_ptr(), ws["gi"][i].data_ptr(), gates
            s.w_ih_t = ws["w_ih_t"].data_ptr() if gates else None       # W_ih^T for the backward's data gradient (same step)
        L.call("marl_agent_unroll_fwd", C.byref(d), arr, 2, sp)
        L.call("marl_qtran_select", C.byref(d), ws["q"][0].data_ptr(), ws["q"][1].data_ptr(), bt["avail_u"].data_ptr(),
               bt["avail_u_next"].data_ptr(), bt["u"].data_ptr(), ws["oh_e"].data_ptr(), ws["oh_t"].data_ptr(),
               ws["opt_e"].data_ptr(), ws["q_max"].data_ptr(), ws["q_taken"].data_ptr(), sp)
        M, N, S, A, qh = B * Lq, a.n_agents, a.state_shape, a.n_actions, a.qtran_hidden_dim
        pq = net_struct(self._net_addrs(fl, "mixer.", self.mixer))
        pqt = net_struct(self._net_addrs(tfl, "mixer.", self.target_mixer))
        pv = net_struct(self._net_addrs(fl, "v.", self.v))
        gq = net_struct(self._net_addrs(fl, "mixer.", self.mixer, fl.grad), L.QtranNetGrads)
        gv = net_struct(self._net_addrs(fl, "v.", self.v, fl.grad), L.QtranNetGrads)
        hid_e, hid_t = ws["hidden"][0].data_ptr(), ws["hidden"][1].data_ptr()
        wq, wqt, wqh, wv = (net_ws_struct(ws[k]) for k in ("nq", "nqt", "nqh", "nv"))
        dwq, dwv = net_ws_struct(ws["dnq"]), net_ws_struct(ws["dnv"])
        # get_qtran (qtran_learner.py:165-200): the four network evaluations (joint Q at the taken actions, target joint Q, joint Q at
        # the greedy actions, V) only share their inputs -- chains of small, latency-bound products: three of them run on side
        # streams beside the first
        cur = th.cuda.current_stream()
        if getattr(self, "_net_streams", None) is None:
            self._net_streams = [th.cuda.Stream() for _ in range(3)]
        calls = (
            (pq, bt["s"].data_ptr(), hid_e, bt["u_onehot"].data_ptr(), wq, ws["jq"], A),
            (pqt, bt["s_next"].data_ptr(), hid_t, ws["oh_t"].data_ptr(), wqt, ws["jq_t"], A),
            (pq, bt["s"].data_ptr(), hid_e, ws["oh_e"].data_ptr(), wqh, ws["jq_hat"], A),
            (pv, bt["s"].data_ptr(), hid_e, None, wv, ws["vv"], 0),
        )
        for k, (pp, st_, hid, act, w_, out, a_enc) in enumerate(calls):
            if k == 0:
                L.call("marl_qtran_net_fwd", M, N, S, a_enc, qh, C.byref(pp), st_, hid, act, C.byref(w_), out.data_ptr(), sp)
                continue
            side = self._net_streams[k - 1]
            side.wait_stream(cur)
            with th.cuda.stream(side):
                L.call("marl_qtran_net_fwd", M, N, S, a_enc, qh, C.byref(pp), st_, hid, act, C.byref(w_), out.data_ptr(), L.stream_ptr())
        for side in self._net_streams:
            cur.wait_stream(side)
        L.call("marl_qtran_losses_fwd_bwd", C.byref(d), ws["jq"].data_ptr(), ws["jq_t"].data_ptr(), ws["jq_hat"].data_ptr(),
               ws["vv"].data_ptr(), ws["q_max"].data_ptr(), ws["q_taken"].data_ptr(), ws["opt_e"].data_ptr(),
               bt["u"].data_ptr(), bt["avail_u"].data_ptr(), bt["r"].data_ptr(), bt["terminated"].data_ptr(),
               bt["padded"].data_ptr(), float(self.gamma), float(a.lambda_opt), float(a.lambda_nopt), ws["d_jq"].data_ptr(),
               ws["d_v"].data_ptr(), ws["dq"].data_ptr(), fl.tail.data_ptr(), ws["parts"].data_ptr(), sp)
        # (measured and dropped: V's backward on a side stream into its own dL/dhidden buffer, folded afterwards -- 2851 .. 2951 us
        # per step against 2883: the two chains' products already share the machine through the weight-gradient lane)
        L.call("marl_qtran_net_bwd", M, N, S, A, qh, C.byref(pq), bt["s"].data_ptr(), hid_e, bt["u_onehot"].data_ptr(),
               C.byref(wq), ws["d_jq"].data_ptr(), C.byref(dwq), ws["dhid"].data_ptr(), 0, C.byref(gq), sp)
        L.call("marl_qtran_net_bwd", M, N, S, 0, qh, C.byref(pv), bt["s"].data_ptr(), hid_e, None, C.byref(wv),
               ws["d_v"].data_ptr(), C.byref(dwv), ws["dhid"].data_ptr(), 1, C.byref(gv), sp)
        bw = L.UnrollBwd()
        bw.obs, bw.onehot, bw.shift_onehot, bw.full_input = bt["o"].data_ptr(), bt["u_onehot"].data_ptr(), 1, 0
        bw.params = pe
        bw.hidden, bw.x, bw.gates = ws["hidden"][0].data_ptr(), ws["x"][0].data_ptr(), ws["gates"].data_ptr()
        bw.h0, bw.dq, bw.dhidden = None, ws["dq"].data_ptr(), ws["dhid"].data_ptr()
        bw.dhext, bw.dgi, bw.dgh, bw.dx = (ws[k].data_ptr() for k in ("dhext", "dgi", "dgh", "dx"))
        bw.dh0 = None
        bw.w_ih_t = ws["w_ih_t"].data_ptr()
        bw.grads = agent_param_struct({n: fl.ptr("agent." + n, fl.grad) for n in AGENT_FLAT_ORDER}, L.AgentGrads)
        L.call("marl_agent_unroll_bwd", C.byref(d), C.byref(bw), sp)
        return 2 + 7 + 1 + 4 * 6 + 1 + 2 * 11 + 7

    def train(self, batch, train_step, episode_num=None):
        loss = super().train(batch, train_step, episode_num)
        ws = self.last["ws"]
        # hidden states as the reference leaves them (no double-Q unroll here)
        self.eval_net.hidden_states = ws["h_last"][0]
        self.target_net.hidden_states = ws["h_last"][1]
        return loss

    def save_models(self, train_step):
        num = str(train_step // self.args.save_cycle)
        if not os.path.exists(self.model_dir):
            os.makedirs(self.model_dir)
        self.eval_net.save_models(self.model_dir + '/' + num + '_rnn_net_params.pkl')
        th.save(self.mixer.state_dict(), self.model_dir + '/' + num + '_mixer_net_params.pkl')
        th.save(self.v.state_dict(), self.model_dir + '/' + num + '_v_net_params.pkl')

    def load_models(self):
        if os.path.exists(self.model_dir + '/rnn_net_params.pkl'):
            self.eval_net.load_models(self.model_dir + '/rnn_net_params.pkl')
            self.mixer.load_state_dict(th.load(self.model_dir + '/mixer_net_params.pkl', map_location=self._dev))
            self.v.load_state_dict(th.load(self.model_dir + '/v_net_params.pkl', map_location=self._dev))
        else:
            raise Exception("No model!")
