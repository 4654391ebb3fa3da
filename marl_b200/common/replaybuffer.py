"""Device-resident episode replay buffer, drop-in for ``common/replaybuffer.py:5-80``.

Same surface and semantics as the reference's ``ReplayBuffer`` -- 11 episode-major keys of shape
``[buffer_size, episode_limit, ...]`` (``:19-30``), ``store_episode`` writing at the ring indices
``_get_storage_idx`` hands out (``:35-51, :63-80``: contiguous, wrap-around, restart at 0 once full),
``sample`` drawing ``batch_size`` episodes WITH replacement by ``np.random.randint(0, current_size,
batch_size)`` (``:54-60``; the same host RNG call, so a seeded run samples the same episodes) -- but the
ring lives in HBM as fp32 (``u`` int64), i.e. exactly what ``QLearner.train`` consumes:

* ``store_episode`` pays the float64 -> fp32 cast and the host -> device copy ONCE per collected
  episode (the reference pays both for every sampled batch inside ``train``, q_learner.py:74-91);
* ``sample`` returns a :class:`DeviceEpisodeBatch`: the ring plus the drawn indices.  The learner gathers
  (and truncates to the batch's ``max_episode_len``) straight into its working set with ONE
  ``marl_replay_gather_f32`` launch; nothing is materialised unless somebody indexes the dict, in which
  case the key is gathered on demand and behaves like the reference's array.

The first-terminated index of every stored episode is kept on the host, so the episode-length cut of
``get_max_episode_len`` (q_learner.py:49-66) needs no device synchronisation.

One difference from the reference, which COPIES the sampled rows inside ``sample``: the batch returned here refers to
the ring, so it must be trained on (or materialised) before a ``store_episode`` overwrites one of its rows.  The
reference's loop (runner.py:92-97: store, then sample, then train) satisfies that; a collector thread that stores
between ``sample`` and ``train`` does not, and is DETECTED: every ring row carries a version counter, and
``DeviceEpisodeBatch.check_fresh`` (called by the learner) raises if a sampled row was rewritten meanwhile.
"""
from __future__ import annotations

import threading

import numpy as np
import torch as th

KEYS = ("o", "u", "s", "r", "o_next", "s_next", "avail_u", "avail_u_next", "u_onehot", "padded", "terminated")


class DeviceEpisodeBatch(dict):
    """What ``ReplayBuffer.sample`` returns: lazily materialised view of the sampled episodes.

    Behaves like the reference's dict of ``[batch, episode_limit, ...]`` arrays (keys are gathered on
    first access as CUDA tensors); ``QLearner.train`` recognises it and skips the materialisation."""

    def __init__(self, ring, idx_host, idx_dev, max_episode_len, owner=None, versions=None):
        super().__init__()
        self.ring, self.idx_host, self.idx = ring, idx_host, idx_dev
        self.max_episode_len = int(max_episode_len)
        self.owner, self.versions = owner, versions

    def check_fresh(self):
        """Raises if an episode of this batch was overwritten in the ring after ``sample`` drew it."""
        if self.owner is not None and not np.array_equal(self.owner.version[self.idx_host], self.versions):
            raise RuntimeError("ReplayBuffer: a sampled episode was overwritten by store_episode() before the batch was "
                               "consumed; train on (or index) a sampled batch before storing new episodes")

    def __missing__(self, key):
        if key == "max_episode_len":
            return self.max_episode_len
        if key not in self.ring:
            raise KeyError(key)
        self.check_fresh()
        val = self.ring[key].index_select(0, self.idx)
        self[key] = val
        return val

    def keys(self):
        return list(KEYS)

    def __contains__(self, key):
        return key in KEYS or dict.__contains__(self, key)

    def get(self, key, default=None):
        try:
            return self[key]
        except KeyError:
            return default

    def __len__(self):
        return len(KEYS)


class ReplayBuffer:
    def __init__(self, args, device=None):
        self.args = args
        self.n_actions = self.args.n_actions
        self.n_agents = self.args.n_agents
        self.state_shape = self.args.state_shape
        self.obs_shape = self.args.obs_shape
        self.size = self.args.buffer_size
        self.episode_limit = self.args.episode_limit
        # memory management
        self.current_idx = 0
        self.current_size = 0
        if device is None:
            if not th.cuda.is_available():
                raise RuntimeError("marl_b200.ReplayBuffer keeps the episodes in HBM: no CUDA device")
            device = th.device("cuda")
        self.device = th.device(device)
        S, T, N, A, O, St = self.size, self.episode_limit, self.n_agents, self.n_actions, self.obs_shape, self.state_shape
        f = lambda *shape: th.zeros(*shape, dtype=th.float32, device=self.device)
        self.buffers = {'o': f(S, T, N, O),
                        'u': th.zeros(S, T, N, 1, dtype=th.int64, device=self.device),
                        's': f(S, T, St),
                        'r': f(S, T, 1),
                        'o_next': f(S, T, N, O),
                        's_next': f(S, T, St),
                        'avail_u': f(S, T, N, A),
                        'avail_u_next': f(S, T, N, A),
                        'u_onehot': f(S, T, N, A),
                        'padded': f(S, T, 1),
                        'terminated': f(S, T, 1)}
        # index of the first terminated == 1 per stored episode (-1: none), for the episode-length cut
        self.first_terminated = np.full(self.size, -1, dtype=np.int64)
        self.version = np.zeros(self.size, dtype=np.int64)      # bumped whenever a ring row is rewritten
        self._stage = {}                                          # pinned / device staging of the host fast path
        self._idx_stage = {}
        self._pending_read = None
        self.pinned_stores = 0                                    # episodes ingested straight from page-locked arrays
        self.lock = threading.Lock()

    # ---- store ------------------------------------------------------------------------------------
    def _store_host_f64(self, episode_batch, batch_size, start):
        """Fast path of ``store_episode`` for what the rollout produces (rollout.py:135-149): float64 numpy arrays going
        to CONTIGUOUS ring rows.  The 11 keys are packed into one pinned buffer, cross PCIe in ONE copy and are cast to
        the ring's fp32 / int64 layout by ONE ``marl_ingest_f64`` launch writing the ring rows in place."""
        import ctypes as C
        from .. import _lib as L
        if self._store_pinned_in_place(episode_batch, batch_size, start):
            return
        st = self._stage.get(batch_size)
        if st is None:
            sizes = [batch_size * int(np.prod(self.buffers[k].shape[1:])) for k in KEYS]
            offs = np.concatenate(([0], np.cumsum(sizes)))
            total = int(offs[-1])
            # two slots: the pinned buffer of call k may still be read by the DMA engine while call k+1 packs its episode
            st = {"host": [th.empty(total, dtype=th.float64).pin_memory() for _ in range(2)],
                  "dev": [th.empty(total, dtype=th.float64, device=self.device) for _ in range(2)],
                  "done": [None, None], "seq": 0, "e64": [], "views": [],
                  "row_bytes": [self.buffers[k][0].numel() * self.buffers[k].element_size() for k in KEYS],
                  "base": [self.buffers[k].data_ptr() for k in KEYS],
                  "dims": L.Dims(batch_size, self.episode_limit, self.n_agents, self.n_actions, self.obs_shape, self.state_shape)}
            for slot in range(2):
                e64 = L.EpisodeF64()
                hnp = st["host"][slot].numpy()
                views = []
                for i, k in enumerate(KEYS):
                    setattr(e64, k, st["dev"][slot].data_ptr() + 8 * int(offs[i]))
                    views.append(hnp[int(offs[i]):int(offs[i + 1])].reshape((batch_size,) + tuple(self.buffers[k].shape[1:])))
                e64.u_is_int64 = 0
                st["e64"].append(e64)
                st["views"].append(views)
            self._stage[batch_size] = st
        slot = st["seq"] & 1
        st["seq"] += 1
        if st["done"][slot] is not None:
            st["done"][slot].synchronize()
        for view, k in zip(st["views"][slot], KEYS):
            np.copyto(view, episode_batch[k], casting="unsafe")
        st["dev"][slot].copy_(st["host"][slot], non_blocking=True)
        e32 = L.EpisodeF32()
        for i, k in enumerate(KEYS):
            setattr(e32, k, st["base"][i] + start * st["row_bytes"][i])
        L.call("marl_ingest_f64", C.byref(st["e64"][slot]), self.episode_limit, C.byref(st["dims"]), C.byref(e32), L.stream_ptr())
        ev = th.cuda.Event()
        ev.record()
        st["done"][slot] = ev

    def _store_pinned_in_place(self, episode_batch, batch_size, start):
        """Zero-copy variant of the fast path: when every array of the episode already lives in page-locked host memory
        (``torch.Tensor.pin_memory().numpy()``, ``cudaHostRegister``), the cast kernel reads the float64 values over PCIe
        itself and writes the fp32 / int64 ring rows -- no packing pass on the host, no staging copy.  ``store_episode``
        returns only when the kernel has read the arrays, so the caller may reuse them at once, exactly as after the
        reference's synchronous numpy copy (common/replaybuffer.py:30-61)."""
        import ctypes as C
        from .. import _lib as L
        meta = self._stage.get(("meta", batch_size))
        if meta is None:
            n = len(KEYS)
            meta = {"row_bytes": [self.buffers[k][0].numel() * self.buffers[k].element_size() for k in KEYS],
                    "base": [self.buffers[k].data_ptr() for k in KEYS],
                    "dims": L.Dims(batch_size, self.episode_limit, self.n_agents, self.n_actions, self.obs_shape, self.state_shape),
                    "e64": L.EpisodeF64(), "e32": L.EpisodeF32(), "done": th.cuda.Event(),
                    "ptrs": (C.c_void_p * n)(), "bytes": (C.c_size_t * n)(), "check": L.load().marl_host_registered_all}
            meta["e64"].u_is_int64 = 0
            self._stage[("meta", batch_size)] = meta
        ptrs, nbytes = meta["ptrs"], meta["bytes"]
        for i, k in enumerate(KEYS):
            a = episode_batch[k]
            p = a.__array_interface__["data"][0]          # (a.ctypes.data builds a ctypes object per call: ~1 us each)
            if not a.flags.c_contiguous or a.dtype != np.float64 or p % 16:
                return False
            ptrs[i], nbytes[i] = p, a.nbytes
        if meta["check"](ptrs, nbytes, len(KEYS)) != 1:     # every array page-locked and device-readable (one call for the 11)
            return False
        e64, e32 = meta["e64"], meta["e32"]
        for i, k in enumerate(KEYS):
            setattr(e64, k, ptrs[i])
            setattr(e32, k, meta["base"][i] + start * meta["row_bytes"][i])
        L.call("marl_ingest_f64", C.byref(e64), self.episode_limit, C.byref(meta["dims"]), C.byref(e32), L.stream_ptr())
        meta["done"].record()
        self._pending_read = meta["done"]        # store_episode waits on it after its host-side bookkeeping
        self.pinned_stores += 1
        return True

    def store_episode(self, episode_batch):
        batch_size = episode_batch['o'].shape[0]
        with self.lock:
            idxs = self._get_storage_idx(inc=batch_size)
            idx_np = np.atleast_1d(np.asarray(idxs, dtype=np.int64))
            contiguous = batch_size == 1 or bool(np.all(np.diff(idx_np) == 1))
            host_f64 = all(isinstance(episode_batch[k], np.ndarray) and episode_batch[k].dtype == np.float64 for k in KEYS
                           if k != 'u') and isinstance(episode_batch['u'], np.ndarray) and episode_batch['u'].dtype.kind == 'f' \
                and episode_batch['o'].shape[1] == self.episode_limit and self.device.type == "cuda"
            if contiguous and host_f64:
                self._store_host_f64(episode_batch, batch_size, int(idx_np[0]))
            else:
                self._store_generic(episode_batch, batch_size, idx_np)
            term = episode_batch['terminated']
            term = term.detach().cpu().numpy() if th.is_tensor(term) else np.asarray(term)
            if batch_size == 1:                     # (the rollout's case: scalar bookkeeping instead of five numpy calls)
                hit = np.flatnonzero(term.reshape(-1)[:self.episode_limit] == 1)
                row = int(idx_np[0])
                self.first_terminated[row] = int(hit[0]) if hit.size else -1
                self.version[row] += 1
            else:
                term = term.reshape(batch_size, -1)[:, :self.episode_limit] == 1
                first = np.where(term.any(axis=1), term.argmax(axis=1), -1)
                self.first_terminated[idx_np] = first
                self.version[idx_np] += 1
            if self._pending_read is not None:       # zero-copy path: the kernel has finished reading the caller's arrays
                self._pending_read.synchronize()
                self._pending_read = None

    def _store_generic(self, episode_batch, batch_size, idx_np):
        """Any other input (CUDA tensors from the batched environment, int64 ``u``, wrapped ring positions): per-key copies."""
        idx_dev = th.from_numpy(idx_np).to(self.device)
        for key in KEYS:
            src = episode_batch[key]
            want = th.int64 if key == 'u' else th.float32
            if th.is_tensor(src):
                t = src.to(device=self.device)
                t = t.to(want) if t.dtype != want else t        # float -> int64 truncates toward zero like th.tensor(..., long)
            else:
                a = np.ascontiguousarray(src)
                t = th.from_numpy(a).to(self.device, non_blocking=True).to(want)
            self.buffers[key].index_copy_(0, idx_dev, t.reshape((batch_size,) + tuple(self.buffers[key].shape[1:])))

    # ---- sample -----------------------------------------------------------------------------------
    def sample(self, batch_size):
        idx = np.random.randint(0, self.current_size, batch_size)
        return self.sample_at(idx)

    def sample_at(self, idx):
        """The batch made of the given ring rows (what ``sample`` does after drawing the indices)."""
        idx = np.asarray(idx, dtype=np.int64)
        with self.lock:
            ft = self.first_terminated[idx]
            versions = self.version[idx].copy()
        m = int(ft.max())                         # episodes that never terminate (-1) do not count (q_learner.py:49-61)
        max_len = m + 1 if m >= 0 else int(self.episode_limit)
        return DeviceEpisodeBatch(self.buffers, idx, self._indices_to_device(idx), max_len, owner=self, versions=versions)

    def _indices_to_device(self, idx):
        """Sampled ring rows -> device int64, through a small ring of pinned staging buffers (an asynchronous copy instead
        of the blocking pageable one)."""
        if self.device.type != "cuda":
            return th.from_numpy(idx).to(self.device)
        n = idx.shape[0]
        ring = self._idx_stage.get(n)
        if ring is None:
            host = [th.empty(n, dtype=th.int64).pin_memory() for _ in range(16)]
            ring = {"host": host, "view": [h.numpy() for h in host], "seq": 0}
            self._idx_stage[n] = ring
        # (a pinned slot is only needed until its asynchronous copy has run; sixteen of them: a caller would have to queue
        # sixteen samples without ever touching their data for a slot to be rewritten early)
        slot = ring["seq"] & 15
        ring["seq"] += 1
        ring["view"][slot][:] = idx
        return ring["host"][slot].to(self.device, non_blocking=True)

    # ---- ring indices (common/replaybuffer.py:63-80) -----------------------------------------------
    def _get_storage_idx(self, inc=None):
        """Ring positions of the next `inc` episodes.  Semantics of the reference, including its quirk that a
        write ending exactly at the end of the ring leaves the cursor AT `size` (the next write restarts at 0):
        contiguous while it fits, otherwise the tail of the ring followed by its head."""
        n = inc or 1
        start = self.current_idx if self.current_idx < self.size else 0
        end = start + n
        positions = (start + np.arange(n)) % self.size if end > self.size else np.arange(start, end)
        self.current_idx = end if end <= self.size else end - self.size
        self.current_size = min(self.size, self.current_size + n)
        return positions[0] if n == 1 else positions
