#!/usr/bin/env python
"""bench.py -- QMIX learner episode-samples/s (2s3z shape) + matrix-game env-steps/s on B200.

    python bench.py --gpus N --steps K --warmup W          (N>1: launched under torch.distributed.run)
    python bench.py --impl reference ...                   (the CPU path of the reference algorithm)

One JSON line on stdout (rank 0).  A "step" is one ``QLearner.train()`` on one synthetic 2s3z-shaped
episode batch (BASELINE.json configs[1]: 5 agents, 11 actions, T=120, batch 32 per GPU).
  value : whole-job episode-samples/s with the batches already resident in HBM (CUDA events, max over ranks)
  e2e   : the same through the reference-facing call with pinned float64 HOST batches in the ReplayBuffer
          layout -- H2D copy, f64->f32 ingest and the loss read-back inside the timed region
  env   : matrix-game env-steps/s of the batched environment kernel (4096 envs and 2^24 envs)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SHAPE = dict(B=32, T=120, N=5, A=11, O=80, S=120)        # BASELINE.json configs[1]
FLOP_PER_STEP = 6.98e9                                   # SURVEY.md section 8(d), cfg 2 QMIX, B=32
BYTES_PER_STEP = 18.71e6                                 # each batch element read once as fp32 (u int64)
GRU_FWD_FLOP = 3 * 32 * 120 * 5 * 2 * (3 * 64 * 64)      # 3 unrolls x rows x (W_hh h): the sequential kernel
GRU_FWD_BYTES = 32 * 120 * 5 * 4 * (3 * 192 + 3 * 64 + 256)   # gi in (3 unrolls), hidden out (3), saved gates out (eval unroll)
ENV_BYTES = 140                                          # 16 B actions in + 124 B episode record out
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures (profiles/)
NCU_TRAFFIC = {"gru_unroll_fwd_kernel": 44405504 + 3192832, "gru_unroll_bwd_kernel": 29584128 + 183808,
               "linear_fwd_kernel": 7062784, "matrix_game_step_kernel": 268475136 + 2043086000}   # profiles/r1c_ncu_full_gru.txt, r1b_ncu_full_env.txt
PAYOFF1 = [[8, -12, -12], [-12, 0, 0], [-12, 0, 0]]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock / throttle reasons sampled every 5 ms DURING the timed regions (NVML; the same counters the
    nvidia-smi line of B200_PROFILING.md prints)."""

    def __init__(self, index=0):
        self.index, self.rows, self._stop, self.t = index, [], False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()
        except Exception as e:      # noqa: BLE001
            self.err = repr(e)

    def _loop(self):
        nv, h = self.nv, self.h
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:   # noqa: BLE001
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((sm, reasons))
            except Exception:       # noqa: BLE001
                pass
            time.sleep(0.005)

    def stop(self):
        if self.t is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")]}
        self._stop = True
        self.t.join(timeout=1)
        nv = self.nv
        sm = [r[0] for r in self.rows]
        mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        seen = 0
        for _, r in self.rows:
            seen |= r
        reasons = sorted(n for n, b in bits.items() if seen & b)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(mx), "reasons": reasons,
                "samples": len(sm)}


def make_args(alg="qmix"):
    from marl_b200.common.arguments import default_args
    return default_args(alg=alg, n_agents=SHAPE["N"], n_actions=SHAPE["A"], obs_shape=SHAPE["O"],
                        state_shape=SHAPE["S"], episode_limit=SHAPE["T"], map="synthetic_2s3z")


def oracle_state(args, seed=0):
    import torch
    from oracle import marl_oracle as MO
    torch.manual_seed(seed)
    cfg = MO.make_cfg(alg=args.alg, n_agents=args.n_agents, n_actions=args.n_actions, obs_shape=args.obs_shape,
                      state_shape=args.state_shape, episode_limit=args.episode_limit)
    return MO, MO.LearnerState(cfg)


WORKLOAD = ("QMIX learner step, synthetic 2s3z-shaped batch (5 agents, 11 actions, T=120, batch 32 per GPU, RMSprop, "
            "double-Q)")


def host_threads():
    """All the host cores this process may use, whatever OMP_NUM_THREADS torchrun exported."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def reference_learner(alg, shape, cuda=False, seed=0):
    """The UNMODIFIED reference (oracle/_ref, staged by oracle/fetch_ref.py) when it travelled with the snapshot,
    else the oracle port of its CPU path.  Returns (train_fn(batch, step) -> loss, kind)."""
    import torch
    from oracle import fetch_ref as FR
    if FR.available():
        ns = FR.load()
        a = FR.make_args(ns, alg, shape["N"], shape["A"], shape["O"], shape["S"], shape["T"], cuda=cuda)
        torch.manual_seed(seed)
        mac = ns.SharedMAC(a)
        learner = (ns.QTRANLearner if alg == "qtran_base" else ns.QLearner)(mac, a)
        return (lambda batch, step: learner.train({k: v.copy() for k, v in batch.items()}, step)), "reference"
    if cuda:
        return None, "port"
    from marl_b200.common.arguments import default_args
    args = default_args(alg=alg, n_agents=shape["N"], n_actions=shape["A"], obs_shape=shape["O"], state_shape=shape["S"],
                        episode_limit=shape["T"], map="synthetic")
    MO, st = oracle_state(args, seed)
    return (lambda batch, step: MO.train_step(st, batch, step)[0]), "port"


def time_reference(steps, warmup, alg="qmix", shape=None, cuda=False):
    """episode-samples/s of the reference's own train() on the host cores (or, cuda=True, on this GPU with stock ATen
    kernels: the same-box GPU baseline of SURVEY.md section 8(c))."""
    import torch
    from marl_b200.synthetic import synthetic_batch
    shape = shape or SHAPE
    torch.set_num_threads(host_threads())
    train, kind = reference_learner(alg, shape, cuda=cuda)
    if train is None:
        return None
    batch = synthetic_batch(0, **shape)
    for i in range(warmup):
        train(batch, i)
    if cuda:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        train(batch, warmup + i)          # returns loss.item(): synchronises every step, like the reference's caller
    if cuda:
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"value": shape["B"] * steps / dt, "unit": "episode-samples/s", "ms_per_step": dt / steps * 1e3, "kind": kind,
            "cores": torch.get_num_threads(), "steps": steps, "warmup": warmup}


def run_reference(opt):
    """--impl reference: the reference's own QLearner.train on this box's host cores (rank 0 only), plus the same
    code with args.cuda=True on GPU 0 as `gpu_baseline`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    steps, warm = max(1, min(opt.steps, 30)), max(1, min(opt.warmup, 3))
    r = time_reference(steps, warm)
    sample = (f"{steps} train steps (after {warm} warm-up) of one B=32 synthetic 2s3z-shaped batch through "
              + ("the UNMODIFIED reference QLearner.train (oracle/_ref), args.cuda=False" if r["kind"] == "reference"
                 else "the oracle port of the reference's CPU path (oracle/_ref was not staged)")
              + f", torch.set_num_threads({r['cores']})")
    gpu = None
    if torch.cuda.is_available() and r["kind"] == "reference":
        try:
            g = time_reference(10, 3, cuda=True)
            gpu = {"value": g["value"], "unit": g["unit"], "ms_per_step": g["ms_per_step"],
                   "what": "UNMODIFIED reference QLearner.train with args.cuda=True on this GPU (stock ATen / cuBLAS kernels; "
                           "o / o_next / avail stay on the host as in controller/share_params.py:132-134)", "steps": 10}
        except Exception as e:      # noqa: BLE001
            gpu = {"unavailable": repr(e)[:200]}
    line = {"impl": "reference", "metric": "QMIX learner episode-samples/sec (2s3z shape)", "value": r["value"],
            "unit": "episode-samples/s", "n_gpus": opt.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": r["value"], "unit": "episode-samples/s", "cores": r["cores"], "kind": r["kind"],
                             "sample": sample},
            "gpu_baseline": gpu,
            "e2e": {"value": r["value"], "unit": "episode-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def bench_env(torch, dist, world, n_envs, iters, hbm_peak):
    """n_envs independent matrix games PER GPU (env instances shard over ranks, no collective)."""
    from marl_b200.env.single_state_matrix_game import BatchedMatrixGame
    env = BatchedMatrixGame(PAYOFF1, n_envs)
    acts = torch.randint(0, 3, (n_envs, 2), device="cuda", dtype=torch.int64)
    for _ in range(5):
        env.step(acts)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        env.step(acts)
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / iters], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    rate = n_envs * world / (ms * 1e-3)
    gbs = rate / world * ENV_BYTES / 1e9
    return {"n_envs_per_gpu": n_envs, "value": rate, "unit": "env-steps/s", "us_per_launch": ms * 1e3,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                         "traffic": NCU_TRAFFIC["matrix_game_step_kernel"] if n_envs == (1 << 24) else None, "note": "per GPU"}}


def run_ours(opt):
    import torch
    import torch.distributed as dist
    from marl_b200 import _lib as L
    from marl_b200.algorithm.q_learner import QLearner
    from marl_b200.controller.share_params import SharedMAC
    from marl_b200.synthetic import synthetic_batch, KEYS

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hbm_peak, peak_src = peaks()
    K, W = opt.steps, max(opt.warmup, 3)
    B = SHAPE["B"]

    args = make_args()
    torch.manual_seed(0)
    learner = QLearner(SharedMAC(args), args)
    if world > 1:
        learner.enable_data_parallel()

    # Each rank holds its own shard: weak scaling, B=32 episodes per GPU, global batch 32*world.
    NB = 8                                            # 8 x 18.7 MB = 150 MB of inputs > 126 MB L2
    dev_batches, host_batches = [], []
    for i in range(NB):
        hb = synthetic_batch(1000 * rank + i, **SHAPE)
        db = {k: torch.as_tensor(v, device="cuda") for k, v in hb.items()}
        db = {k: (v.to(torch.int64) if k == "u" else v.to(torch.float32)).contiguous() for k, v in db.items()}
        db["max_episode_len"] = SHAPE["T"]
        dev_batches.append(db)
        if i < 2:                                     # pinned float64 host copies in the ReplayBuffer layout
            pinned = {k: torch.from_numpy(v).pin_memory() for k, v in hb.items()}
            host_batches.append({k: t.numpy() for k, t in pinned.items()})
            host_batches[-1]["_keep"] = pinned
    shard = learner._shard
    if world > 1:
        learner._shard = lambda B_glob: (0, B_glob)  # every rank already holds exactly its shard

    def host_view(i):
        return {k: host_batches[i % 2][k] for k in KEYS}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step = 0
    # resident batches are read in place, each through its own captured graph: run every batch through the eager
    # call and the capture call before anything is timed (W warm-up replays follow)
    for _ in range(2):
        for b in dev_batches:
            learner.train(b, step); step += 1
    for i in range(W):
        learner.train(dev_batches[i % NB], step); step += 1
    for i in range(max(W, 3)):
        learner.train(host_view(i), step); step += 1

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- value: device-resident batches -------------------------------------------------------------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(K):
        learner.train(dev_batches[i % NB], step); step += 1
    ev1.record()
    barrier()
    ms_dev = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    # ---- e2e: pinned host float64 batches through the reference-facing call ---------------------------
    barrier()
    ev0.record()
    for i in range(K):
        learner.train(host_view(i), step); step += 1
    ev1.record()
    barrier()
    ms_e2e_sync = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    h2d = learner.h2d_bytes_last
    # same, with the next batch's H2D copy started (learner.prefetch) before the current train() call: the copy
    # of every step is still inside the timed region, it just overlaps the previous step's compute
    hv = [host_view(0), host_view(1)]
    learner.prefetch(hv[0])                           # untimed: first use allocates the two prefetch staging slots
    for i in range(4):
        learner.prefetch(hv[(i + 1) & 1])
        learner.train(hv[i & 1], step); step += 1
    learner.train(hv[0], step); step += 1
    barrier()
    ev0.record()
    learner.prefetch(hv[0])
    for i in range(K):
        if i + 1 < K:
            learner.prefetch(hv[(i + 1) & 1])
        learner.train(hv[i & 1], step); step += 1
    ev1.record()
    barrier()
    ms_e2e = torch.tensor([ev0.elapsed_time(ev1)], device="cuda")
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        dist.all_reduce(ms_dev, op=dist.ReduceOp.MAX)
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
        dist.all_reduce(ms_e2e_sync, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, ms_e2e_sync = float(ms_dev) / K, float(ms_e2e) / K, float(ms_e2e_sync) / K
    launches = learner.launches_per_step + 1          # + the ingest launch (host-batch path; device batches in the
                                                      # working-set layout are read in place: no ingest launch)

    # ---- per-kernel device time (eager pass, CUDA events around every launch of the library) ---------
    learner._use_graph = False
    L.profile(rank == 0)
    P = 10
    for i in range(P):
        L.call("marl_spin_us", 4000, L.stream_ptr())      # let the host run ahead: events then bracket device time only
        learner.train(dev_batches[i % NB], step); step += 1
    prof = L.profile_collect() if rank == 0 else {}
    L.profile(False)
    learner._use_graph = True
    env_small = bench_env(torch, dist, world, 4096, 200, hbm_peak)
    env_big = bench_env(torch, dist, world, 1 << 24, 20, hbm_peak)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    tot_ms = sum(ms for _, ms in prof.values()) or 1.0
    kernels = {k: {"launches_per_step": c / P, "us_per_step": ms * 1e3 / P, "share": ms / tot_ms}
               for k, (c, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])}
    # FP32 FMA peak of this GPU (probe kernel), the compute-roofline denominator
    scratch = torch.zeros(4, device="cuda")
    import ctypes as C
    flops = C.c_double()
    for _ in range(2):
        L.call("marl_fma_probe", scratch.data_ptr(), 1 << 14, 148 * 8, C.byref(flops), L.stream_ptr())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    L.call("marl_fma_probe", scratch.data_ptr(), 1 << 14, 148 * 8, C.byref(flops), L.stream_ptr())
    b.record()
    torch.cuda.synchronize()
    fp32_peak = flops.value / (a.elapsed_time(b) * 1e-3) / 1e12

    # ---- BASELINE config 5: 4096 matrix-game envs stepped on the GPU feeding the QMIX learner directly ----
    cfg5 = None
    if world == 1:
        from marl_b200.common.arguments import default_args
        from marl_b200.env.single_state_matrix_game import BatchedMatrixGame
        a5 = default_args(alg="qmix", n_agents=2, n_actions=3, obs_shape=1, state_shape=1, episode_limit=1, map="matrix")
        l5 = QLearner(SharedMAC(a5), a5)
        env5 = BatchedMatrixGame(PAYOFF1, 4096)
        acts5 = torch.randint(0, 3, (4096, 2), device="cuda")
        for i in range(5):
            ep = dict(env5.step(acts5)); ep["max_episode_len"] = 1
            l5.train(ep, i)
        torch.cuda.synchronize()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        for i in range(50):
            ep = dict(env5.step(acts5)); ep["max_episode_len"] = 1
            l5.train(ep, 5 + i)
        b_.record(); torch.cuda.synchronize()
        ms5 = a_.elapsed_time(b_) / 50
        cfg5 = {"workload": "4096 matrix-game envs (one kernel launch) -> QMIX train step on the emitted device batch",
                "ms_per_iteration": ms5, "env_steps_per_s": 4096 / (ms5 * 1e-3), "episode_samples_per_s": 4096 / (ms5 * 1e-3)}

    peaks_json = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    bf16_peak = peaks_json.get("bf16_tflops_sustained", 1400.0)      # kernels timed inside a long step: the sustained figure
    # ---- larger batches of the same shape (device-resident): where the step leaves the latency-bound regime ----
    sweep = None
    if world == 1 and not opt.no_sweep:
        sweep = []
        for Bs in (256, 1024):
            hb = synthetic_batch(7, **dict(SHAPE, B=Bs))
            db = {k: torch.as_tensor(v, device="cuda") for k, v in hb.items()}
            db = {k: (v.to(torch.int64) if k == "u" else v.to(torch.float32)).contiguous() for k, v in db.items()}
            db["max_episode_len"] = SHAPE["T"]
            for i in range(4):
                learner.train(db, step); step += 1
            torch.cuda.synchronize()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            for i in range(10):
                learner.train(db, step); step += 1
            b_.record(); torch.cuda.synchronize()
            msb = a_.elapsed_time(b_) / 10
            tf = FLOP_PER_STEP * Bs / SHAPE["B"] / (msb * 1e-3) / 1e12
            sweep.append({"batch": Bs, "ms_per_step": msb, "episode_samples_per_s": Bs / (msb * 1e-3), "fp32_tflops": tf,
                          "frac_of_fp32_peak": tf / fp32_peak,
                          "hbm_frac": BYTES_PER_STEP * Bs / SHAPE["B"] / (msb * 1e-3) / 1e9 / hbm_peak})
            del db
        learner._ws = {k: v for k, v in learner._ws.items() if k[0] == "stage" or k == (B, SHAPE["T"])}
        torch.cuda.empty_cache()

    roofline = None
    if kernels:
        # dominant kernel = the longest single launch of the step (it sits on the critical path): the recurrent part of
        # the three agent unrolls.  Neither HBM nor the tensor pipe binds it: it is a dependent chain of 2L small
        # (rows x 64 x 192) fp32 mat-vecs, so it is reported against the measured FP32-FMA peak, with the HBM and
        # tensor-pipe fractions beside it for context.
        dom = max(kernels, key=lambda k: kernels[k]["us_per_step"] / max(kernels[k]["launches_per_step"], 1))
        dom_us = kernels[dom]["us_per_step"] / max(kernels[dom]["launches_per_step"], 1)
        is_gru = dom == "gru_unroll_fwd_kernel"
        flop = GRU_FWD_FLOP if is_gru else FLOP_PER_STEP * kernels[dom]["share"]
        ach = flop / (dom_us * 1e-6) / 1e12
        gru_bytes = GRU_FWD_BYTES if is_gru else None
        roofline = {"kernel": dom, "bound": "fp32",
                    "bound_note": "dependent chain of 2L = 240 GRU steps on B*N = 160 rows (~1.1 rows per SM): latency-bound at "
                                  "B=32; FP32 FMA is the nearest roof (HBM and tensor-pipe fractions given for context; see "
                                  "batch_sweep for the fraction at larger batches)",
                    "achieved": ach, "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach / fp32_peak,
                    "traffic": NCU_TRAFFIC.get(dom), "traffic_source": "profiles/r1c_ncu_full_gru.txt (dram_read + dram_write, one --set full capture)",
                    "us_per_launch": dom_us, "algorithmic_flop_per_launch": flop, "algorithmic_bytes_per_launch": gru_bytes,
                    "hbm_frac": (gru_bytes / (dom_us * 1e-6) / 1e9 / hbm_peak) if gru_bytes else None,
                    "peak_source": "marl_fma_probe on this GPU (MEASURED_PEAKS.json has no fp32 figure); hbm: " + peak_src}
    gemm_us = sum(v["us_per_step"] for k, v in kernels.items() if k.startswith("linear_"))
    gemm_flop = FLOP_PER_STEP - GRU_FWD_FLOP - GRU_FWD_FLOP / 3        # everything but the two recurrent kernels
    roofline_gemm = None
    if gemm_us:
        ach = gemm_flop / (gemm_us * 1e-6) / 1e12
        roofline_gemm = {"kernel": "linear_{fwd,dgrad,wgrad}_kernel (tcgen05 kind::tf32, 3xTF32)", "bound": "tensor",
                         "achieved": ach, "peak": bf16_peak, "unit": "TFLOP/s", "frac": ach / bf16_peak,
                         "traffic": NCU_TRAFFIC.get("linear_fwd_kernel"), "us_per_step": gemm_us,
                         "note": "algorithmic fp32 FLOPs of all dense layers / summed launch time (launches on parallel streams "
                                 "overlap, so the sum overstates the wall time); peak = measured sustained bf16; the TF32 pipe peaks "
                                 "at half of it and every product is issued 3 times (3xTF32), so 1/6 of `peak` is the ceiling of "
                                 "this scheme; at cfg-2 sizes each launch is a single wave of 150 CTAs x 3-6 k-tiles and is "
                                 "bound by fixed per-launch latency and shared-memory bandwidth (profiles/README.md)"}
    step_tflops = FLOP_PER_STEP / (ms_dev * 1e-3) / 1e12

    cpu = None
    if world == 1:
        val, per, cores = time_cpu_reference(20, 3)
        cpu = {"value": val, "unit": "episode-samples/s", "cores": cores, "kind": "port",
               "sample": "20 train steps (after 3 warm-up) of the same B=32 2s3z-shaped batch, oracle port of the "
                         "reference CPU path, all host threads", "ms_per_step": per * 1e3}

    line = {
        "metric": "QMIX learner episode-samples/sec (2s3z shape)", "value": B * world / (ms_dev * 1e-3),
        "unit": "episode-samples/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_dev,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "QMIX learner step, synthetic 2s3z-shaped batch (5 agents, 11 actions, T=120, "
                               "batch 32 per GPU, RMSprop, double-Q)",
                   "global_batch": B * world, "parallelism": f"dp{world}",
                   "gradient_exchange": (None if world == 1 else
                                         "fused into the optimiser launch over NVLink peer memory (marl_clip_step_peer)"
                                         if getattr(learner, "_peer", None) is not None else "ncclAllReduce between two graphs"),
                   "l2": f"inputs rotate over {NB} resident batches (150 MB > 126 MB L2)"},
        "e2e": {"value": B * world / (ms_e2e * 1e-3), "unit": "episode-samples/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 8,
                "how": "QLearner.train(host float64 dict) with learner.prefetch(next batch) issued before it: H2D of step "
                       "k+1 overlaps the compute of step k; every copy and every loss read-back is inside the timed region",
                "without_prefetch": {"value": B * world / (ms_e2e_sync * 1e-3), "ms_per_step": ms_e2e_sync}},
        "gpu_launches": int((launches - 1) * K + launches * K),   # value leg (in place) + e2e leg (with ingest)
        "launches_per_step": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "roofline_gemm": roofline_gemm,
        "step_fp32": {"tflops": step_tflops, "frac_of_fp32_peak": step_tflops / fp32_peak, "fp32_peak_tflops": fp32_peak,
                      "hbm_frac": BYTES_PER_STEP / (ms_dev * 1e-3) / 1e9 / hbm_peak, "hbm_peak_gbs": hbm_peak,
                      "peak_source": peak_src},
        "kernels": kernels,
        "batch_sweep": sweep,
        "env": {"metric": "matrix-game env-steps/sec", "value": env_big["value"], "unit": "env-steps/s",
                "bytes_per_env_step": ENV_BYTES, "cfg5_4096_envs": env_small, "bandwidth_regime_2^24_envs": env_big},
        "cfg5_env_plus_learner": cfg5,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-sweep", dest="no_sweep", action="store_true", help="skip the B=256/1024 legs")
    opt = ap.parse_args()
    if opt.impl == "reference":
        run_reference(opt)
    else:
        run_ours(opt)


if __name__ == "__main__":
    main()
