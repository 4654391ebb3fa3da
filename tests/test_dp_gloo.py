"""World-size-2 data-parallel protocol on CPU (gloo): sharding + one all-reduce of
[unnormalised grad | loss_sum | mask_sum] reproduces the single-process step.  The per-rank compute
is the CPU oracle (the CUDA kernels need a GPU); the protocol code under test is marl_b200.parallel."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from marl_b200.parallel import allreduce_flat, shard_bounds
from marl_b200.synthetic import synthetic_batch
from oracle import marl_oracle as MO

SHAPE = dict(B=6, T=5, N=3, A=4, O=5, S=6)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _unnormalised(st, batch):
    """sum((mask*delta)^2), sum(mask) and the gradient of the former, on one shard."""
    out = MO.forward(st, batch)
    L = out["L"]
    mask_sum = float((1 - np.asarray(batch["padded"])[:, :L]).sum())
    loss_sum = out["loss"] * mask_sum
    plist = [p for _, _, p in st.flat_params()]
    grads = torch.autograd.grad(loss_sum, plist)
    return torch.cat([g.reshape(-1) for g in grads] + [loss_sum.detach().reshape(1), torch.tensor([mask_sum])])


def _worker(rank, world, port, init_params, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    cfg = MO.make_cfg(alg="qmix", n_agents=SHAPE["N"], n_actions=SHAPE["A"], obs_shape=SHAPE["O"],
                      state_shape=SHAPE["S"], episode_limit=SHAPE["T"])
    st = MO.LearnerState(cfg, init_params)
    batch = synthetic_batch(0, **SHAPE)
    # L must be the GLOBAL max episode length: truncate before sharding
    L = MO.max_episode_len(batch["terminated"], cfg.episode_limit)
    lo, hi = shard_bounds(SHAPE["B"], world, rank)
    shard = {k: v[lo:hi, :L] for k, v in batch.items()}
    cfg.episode_limit = L
    flat = _unnormalised(st, shard)
    allreduce_flat(flat, dist)
    n = flat.numel() - 2
    grads = flat[:n] / flat[-1]
    loss = float(flat[-2] / flat[-1])
    sizes = [p.numel() for _, _, p in st.flat_params()]
    glist = [g.view(p.shape) for g, (_, _, p) in zip(torch.split(grads, sizes), st.flat_params())]
    MO.clip_and_step(st, glist)
    ret[rank] = (loss, torch.cat([p.detach().reshape(-1) for _, _, p in st.flat_params()]).numpy())
    dist.destroy_process_group()


def test_two_rank_allreduce_matches_single_process():
    torch.manual_seed(0)
    cfg = MO.make_cfg(alg="qmix", n_agents=SHAPE["N"], n_actions=SHAPE["A"], obs_shape=SHAPE["O"],
                      state_shape=SHAPE["S"], episode_limit=SHAPE["T"])
    ref = MO.LearnerState(cfg)
    init = {g: {k: v.detach().clone() for k, v in sd.items()} for g, sd in ref.params.items()}
    ref_loss, _ = MO.train_step(ref, synthetic_batch(0, **SHAPE), 0)
    ref_flat = torch.cat([p.detach().reshape(-1) for _, _, p in ref.flat_params()]).numpy()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), init, ret), nprocs=2, join=True)
    assert set(ret.keys()) == {0, 1}
    (l0, p0), (l1, p1) = ret[0], ret[1]
    assert l0 == l1 and np.array_equal(p0, p1)                 # replicas stay identical without a broadcast
    assert abs(l0 - ref_loss) < 1e-5 * abs(ref_loss)
    assert np.max(np.abs(p0 - ref_flat)) < 1e-4 * np.max(np.abs(ref_flat))


def test_shard_bounds():
    assert [shard_bounds(32, 4, r) for r in range(4)] == [(0, 8), (8, 16), (16, 24), (24, 32)]
    try:
        shard_bounds(10, 4, 0)
    except ValueError:
        pass
    else:
        raise AssertionError("uneven shards must be rejected")
