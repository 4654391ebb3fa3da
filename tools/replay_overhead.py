"""Host-side cost of the pieces of the replay loop (store_episode / sample / train) at the 2s3z shape."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from bench import make_args, SHAPE, to_device_batch
from marl_b200.algorithm.q_learner import QLearner
from marl_b200.controller.share_params import SharedMAC
from marl_b200.common.replaybuffer import ReplayBuffer
from marl_b200.synthetic import synthetic_batch, KEYS

args = make_args(); args.buffer_size = 512
torch.manual_seed(0)
learner = QLearner(SharedMAC(args), args)
buf = ReplayBuffer(args)
full = "--full" in sys.argv
for i in range(0, 512, 32):
    hb = synthetic_batch(i, **SHAPE)
    if full:
        hb = synthetic_batch(i, **SHAPE, min_len=SHAPE["T"])
    buf.store_episode({k: to_device_batch(torch, hb, SHAPE["T"])[k] for k in KEYS})
eps = []
for i in range(8):
    hb = synthetic_batch(5000 + i, **dict(SHAPE, B=1))
    if "--pinned" in sys.argv:      # page-locked source arrays: the zero-copy ingest path
        keep = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).pin_memory() for k, v in hb.items()}
        eps.append({k: t.numpy() for k, t in keep.items()}); eps[-1]["_keep"] = keep
    else:
        eps.append({k: np.ascontiguousarray(v, dtype=np.float64) for k, v in hb.items()})
np.random.seed(0)
T = {"store": [], "sample": [], "train": [], "L": []}
for i in range(60):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); buf.store_episode(eps[i % 8]); t1 = time.perf_counter()
    b = buf.sample(32); t2 = time.perf_counter()
    learner.train(b, i); t3 = time.perf_counter()
    T["store"].append(t1 - t0); T["sample"].append(t2 - t1); T["train"].append(t3 - t2); T["L"].append(b.max_episode_len)
for k in ("store", "sample", "train"):
    v = np.array(T[k][20:]) * 1e6
    print(f"{k:7s} median {np.median(v):8.1f} us  mean {v.mean():8.1f}  max {v.max():9.1f}")
print("L values:", T["L"][20:])
if "--pinned" in sys.argv:
    from marl_b200 import _lib as L
    lib = L.load()
    e = eps[0]
    t0 = time.perf_counter()
    for _ in range(100):
        for k in KEYS:
            lib.marl_host_registered(e[k].ctypes.data, e[k].nbytes)
    t1 = time.perf_counter()
    print(f"11 x marl_host_registered: {(t1 - t0) / 100 * 1e6:.1f} us")
    t0 = time.perf_counter()
    for _ in range(100):
        ok = buf._store_pinned_in_place(e, 1, 3)
    t1 = time.perf_counter()
    print(f"_store_pinned_in_place (checks + launch + wait): {(t1 - t0) / 100 * 1e6:.1f} us  ok={ok}")
    t0 = time.perf_counter()
    for _ in range(100):
        term = e['terminated'].reshape(1, -1)[:, :buf.episode_limit] == 1
        first = np.where(term.any(axis=1), term.argmax(axis=1), -1)
    t1 = time.perf_counter()
    print(f"first-terminated bookkeeping: {(t1 - t0) / 100 * 1e6:.1f} us")
