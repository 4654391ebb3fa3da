"""Device replay buffer -> learner: sampling from the HBM-resident ring and training on the
sampled view must give what the reference's numpy buffer + host batch gives."""
import numpy as np
import pytest
import torch

from marl_b200.common.replaybuffer import ReplayBuffer, KEYS
from marl_b200.synthetic import synthetic_batch
from oracle.replay_oracle import OracleReplayBuffer
from tests import parity_util as PU

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("alg", ["vdn", "qmix"])
def test_sampled_view_trains_like_the_host_batch(alg):
    N, A, O, S, T = 3, 4, 5, 6, 8
    args = PU.make_args(alg, N, A, O, S, T, buffer_size=10)
    la, st = PU.build_pair(args)
    lb, _ = PU.build_pair(args)
    buf = ReplayBuffer(args)
    orc = OracleReplayBuffer(10, T, N, A, O, S)
    for seed in range(4):                                    # 4 x 4 episodes into 10 slots: the ring wraps
        ep = synthetic_batch(seed, 4, T, N, A, O, S, full_length_first=False, min_len=2)
        buf.store_episode({k: v.copy() for k, v in ep.items()})
        orc.store(ep)
    for k in KEYS:                                           # ring contents bit-exact (fp32 cast of the float64 reference ring)
        want = orc.rings[k].astype(np.int64) if k == "u" else orc.rings[k].astype(np.float32)
        assert np.array_equal(buf.buffers[k].cpu().numpy(), want), k
    from oracle import marl_oracle as MO
    for step in range(4):
        np.random.seed(100 + step)
        host_batch, idx = orc.sample(6)
        np.random.seed(100 + step)
        view = buf.sample(6)
        assert list(view.idx_host) == list(idx)
        a = la.train({k: v.copy() for k, v in host_batch.items()}, step)   # reference-style host float64 dict
        b = lb.train(view, step)                                           # gathered from the ring in one launch
        assert lb.max_episode_len == la.max_episode_len
        assert abs(a - b) <= 1e-6 * abs(a)
        oloss, _ = MO.train_step(st, host_batch, step)
        assert abs(b - oloss) <= 1e-5 * abs(oloss)


def test_sampled_view_materialises_like_the_reference_dict():
    N, A, O, S, T = 2, 3, 4, 5, 6
    args = PU.make_args("vdn", N, A, O, S, T, buffer_size=5)
    buf = ReplayBuffer(args)
    ep = synthetic_batch(0, 5, T, N, A, O, S)
    buf.store_episode(ep)
    view = buf.sample_at([4, 0, 0, 2])
    for k in KEYS:
        want = ep[k][[4, 0, 0, 2]]
        want = want.astype(np.int64) if k == "u" else want.astype(np.float32)
        assert np.array_equal(view[k].cpu().numpy(), want), k


def test_pinned_episode_is_ingested_in_place():
    """store_episode on arrays that already live in page-locked host memory takes the zero-copy path (the cast kernel reads
    them over PCIe): same ring contents as the packed path and as the reference's numpy ring (common/replaybuffer.py:30-61),
    and the source may be overwritten as soon as the call returns."""
    from marl_b200 import _lib as L
    N, A, O, S, T = 5, 11, 80, 120, 30
    args = PU.make_args("qmix", N, A, O, S, T, buffer_size=6)
    pinned_buf, packed_buf = ReplayBuffer(args), ReplayBuffer(args)
    orc = OracleReplayBuffer(6, T, N, A, O, S)
    lib = L.load()
    for seed in range(8):                                     # one episode at a time, the ring wraps
        ep = synthetic_batch(seed, 1, T, N, A, O, S, full_length_first=False, min_len=2)
        keep = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).pin_memory() for k, v in ep.items()}
        src = {k: t.numpy() for k, t in keep.items()}
        assert all(lib.marl_host_registered(v.ctypes.data, v.nbytes) == 1 for v in src.values())
        assert lib.marl_host_registered(ep["o"].ctypes.data, ep["o"].nbytes) == 0          # pageable numpy memory
        pinned_buf.store_episode(src)
        for v in src.values():
            v.fill(-7.0)                                       # the caller reuses its arrays at once
        packed_buf.store_episode({k: v.copy() for k, v in ep.items()})
        orc.store(ep)
    assert pinned_buf.pinned_stores == 8 and packed_buf.pinned_stores == 0          # the path under test really ran
    for k in KEYS:
        want = orc.rings[k].astype(np.int64) if k == "u" else orc.rings[k].astype(np.float32)
        assert np.array_equal(pinned_buf.buffers[k].cpu().numpy(), want), k
        assert np.array_equal(packed_buf.buffers[k].cpu().numpy(), want), k
    assert np.array_equal(pinned_buf.first_terminated, packed_buf.first_terminated)
