"""2-GPU data-parallel learner (NCCL) against the single-GPU learner on the same global batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from marl_b200.synthetic import synthetic_batch
from tests import parity_util as PU

pytestmark = pytest.mark.gpu
SHAPE = dict(B=8, T=10, N=3, A=5, O=12, S=9)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make_learner(optimizer="RMS"):
    args = PU.make_args("qmix", SHAPE["N"], SHAPE["A"], SHAPE["O"], SHAPE["S"], SHAPE["T"])
    args.optimizer = optimizer
    learner, _ = PU.build_pair(args, seed=0)
    return learner


def _worker(rank, world, port, ret, peer, optimizer):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["MARL_B200_PEER_ALLREDUCE"] = "1" if peer else "0"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    learner = _make_learner(optimizer)
    learner.enable_data_parallel()
    assert (learner._peer is not None) == peer          # the NVLink peer-memory exchange is the one that runs
    batch = synthetic_batch(0, **SHAPE)
    losses = [learner.train({k: v.copy() for k, v in batch.items()}, i) for i in range(4)]   # eager, capture, 2 replays
    ret[rank] = (losses, learner._flat.data.cpu().numpy())
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("peer,optimizer", [(True, "RMS"), (True, "Adam"), (False, "RMS")])
def test_two_gpu_dp_matches_single_gpu(peer, optimizer):
    """peer=True: gradient sum inside the optimiser launch over NVLink peer memory (marl_clip_step_peer, one CUDA
    graph per step); peer=False: ncclAllReduce between two graphs.  Both: replicas bit-identical, single-GPU results."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret, peer, optimizer), nprocs=2, join=True)
    (l0, p0), (l1, p1) = ret[0], ret[1]
    assert l0 == l1 and np.array_equal(p0, p1)          # replicas identical without a broadcast
    single = _make_learner(optimizer)
    batch = synthetic_batch(0, **SHAPE)
    ls = [single.train({k: v.copy() for k, v in batch.items()}, i) for i in range(4)]
    assert np.allclose(l0, ls, rtol=1e-5)
    ps = single._flat.data.cpu().numpy()
    assert np.max(np.abs(p0 - ps)) <= 1e-5 * np.max(np.abs(ps)) + 0.1 * 5e-4 * 4
