"""TEST INFRASTRUCTURE ONLY -- loader for the plain-C matrix-game oracle (oracle/matrix_game_oracle.c)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libmg_oracle.so")


def build(force=False):
    src = os.path.join(HERE, "matrix_game_oracle.c")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", SO, src])
    return SO


def step(payoff, actions, obs_value=0.0):
    """actions int64 [n,2] -> dict of numpy arrays in the device layout + 'r64'."""
    lib = C.CDLL(build())
    actions = np.ascontiguousarray(actions, dtype=np.int64)
    n = actions.shape[0]
    f = lambda *s: np.empty(s, dtype=np.float32)
    out = dict(o=f(n, 1, 2, 1), s=f(n, 1, 1), u=np.empty((n, 1, 2, 1), dtype=np.int64), r=f(n, 1, 1), o_next=f(n, 1, 2, 1),
               s_next=f(n, 1, 1), avail_u=f(n, 1, 2, 3), avail_u_next=f(n, 1, 2, 3), u_onehot=f(n, 1, 2, 3),
               padded=f(n, 1, 1), terminated=f(n, 1, 1))
    r64 = np.empty(n, dtype=np.float64)
    pay = np.ascontiguousarray(payoff, dtype=np.float64).reshape(-1)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.mg_oracle_step.restype = None
    lib.mg_oracle_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float] + [C.c_void_p] * 12
    lib.mg_oracle_step(p(pay), p(actions), n, obs_value, *[p(out[k]) for k in
                       ("o", "s", "u", "r", "o_next", "s_next", "avail_u", "avail_u_next", "u_onehot", "padded", "terminated")],
                       p(r64))
    out["r64"] = r64
    return out
