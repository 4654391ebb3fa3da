"""Synthetic episode batches in the reference's ReplayBuffer layout.

Layout contract: common/replaybuffer.py:19-30 (11 float64 keys, episode-major) with the
padding convention of rollout.py:122-133 (padded steps: every key zero except
``padded`` = ``terminated`` = 1).  Used by bench.py and the parity tests; SC2/SMAC
is not available offline, so SMAC-like shapes are filled with seeded random data.
"""
from __future__ import annotations

import numpy as np

KEYS = ("o", "u", "s", "r", "o_next", "s_next", "avail_u", "avail_u_next", "u_onehot", "padded", "terminated")

# BASELINE.json configs: name -> dict(B, T, N, A, O, S, alg)
CONFIGS = {
    "matrix_game": dict(B=9, T=1, N=2, A=3, O=1, S=1, alg="qmix"),
    "2s3z": dict(B=32, T=120, N=5, A=11, O=80, S=120, alg="qmix"),
    "3s5z": dict(B=128, T=150, N=8, A=14, O=128, S=216, alg="qplex"),
    "27m_vs_30m": dict(B=32, T=180, N=27, A=36, O=285, S=1170, alg="qtran_base"),
    "matrix_game_4096": dict(B=4096, T=1, N=2, A=3, O=1, S=1, alg="qmix"),
}


def synthetic_batch(seed, B, T, N, A, O, S, full_length_first=True, min_len=None):
    """Seeded random batch: lengths L_b ~ U{T/2..T} (L_0 = T), o/s ~ N(0,1) over T+1
    steps split into current/next (so o_next[t] == o[t+1] as in rollout.py:108-111),
    avail ~ Bernoulli(0.7) with action 0 always available, u uniform over available
    actions, r ~ N(0,1); terminated at L_b-1 and on padding."""
    rng = np.random.RandomState(seed)
    lo = max(1, T // 2) if min_len is None else min_len
    lens = rng.randint(lo, T + 1, size=B)
    if full_length_first:
        lens[0] = T
    o_all = rng.randn(B, T + 1, N, O)
    s_all = rng.randn(B, T + 1, S)
    av_all = (rng.rand(B, T + 1, N, A) < 0.7).astype(np.float64)
    av_all[..., 0] = 1.0
    pick = rng.rand(B, T, N)
    r = rng.randn(B, T, 1)
    cnt = av_all[:, :T].sum(-1)
    k = np.minimum((pick * cnt).astype(np.int64), cnt.astype(np.int64) - 1)     # k-th available action
    csum = np.cumsum(av_all[:, :T], axis=-1)
    u = ((csum <= k[..., None]) & True).sum(-1).astype(np.int64)                 # index of the (k+1)-th one
    u = np.minimum(u, A - 1)
    u_onehot = np.zeros((B, T, N, A))
    np.put_along_axis(u_onehot, u[..., None], 1.0, axis=-1)
    t_idx = np.arange(T)[None, :]
    pad = (t_idx >= lens[:, None]).astype(np.float64)                            # [B,T]
    term = ((t_idx >= lens[:, None] - 1)).astype(np.float64)
    live = 1.0 - pad
    batch = dict(
        o=o_all[:, :T] * live[..., None, None],
        u=(u * live[..., None]).astype(np.float64)[..., None],
        s=s_all[:, :T] * live[..., None],
        r=r * live[..., None],
        o_next=o_all[:, 1:] * live[..., None, None],
        s_next=s_all[:, 1:] * live[..., None],
        avail_u=av_all[:, :T] * live[..., None, None],
        avail_u_next=av_all[:, 1:] * live[..., None, None],
        u_onehot=u_onehot * live[..., None, None],
        padded=pad[..., None],
        terminated=term[..., None],
    )
    return {k_: np.ascontiguousarray(v, dtype=np.float64) for k_, v in batch.items()}


def config_batch(name, seed=0, **override):
    c = dict(CONFIGS[name])
    c.update(override)
    return synthetic_batch(seed, c["B"], c["T"], c["N"], c["A"], c["O"], c["S"])
