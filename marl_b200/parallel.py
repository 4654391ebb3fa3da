"""Data-parallel protocol of the learner step.

Episodes never interact before the loss reduction, so the sampled batch shards along the episode axis
with NO data-path collective; the only exchange is ONE all-reduce (sum) per train step over the flat
buffer ``[d/dtheta of sum((mask*delta)^2)  |  sum((mask*delta)^2)  |  sum(mask)]``.  Every rank then
divides by the global sum(mask), clips by the same global norm and applies the same optimiser update to
its replica, so parameters stay bit-identical across ranks without a broadcast.
"""
from __future__ import annotations


def shard_bounds(n_episodes, world_size, rank):
    """Contiguous episode shard of rank `rank`: [lo, hi).  The batch size must divide evenly."""
    per = n_episodes // world_size
    if per * world_size != n_episodes:
        raise ValueError("data-parallel training needs the batch size to be a multiple of the world size")
    return rank * per, (rank + 1) * per


def allreduce_flat(grad_full, dist, group=None):
    """Sum the flat [grad | loss_sum | mask_sum] buffer over the ranks, in place (NCCL on GPUs, gloo in tests)."""
    dist.all_reduce(grad_full, op=dist.ReduceOp.SUM, group=group)
    return grad_full


class PeerGradients:
    """The ranks' flat [grad | loss_sum | mask_sum] buffers in cudaIpc-shared memory (one node, one GPU per process).

    With it the data-parallel exchange needs no library collective: ``marl_clip_step_peer`` reads the W buffers over
    NVLink inside the clip + optimiser launch (two flag barriers around the reads), sums them in rank order and
    updates the local replica, so the whole step stays ONE CUDA graph.  Construction is collective over `group`
    (handles are exchanged with one all_gather); `ok` is False on every rank if any rank could not map its peers,
    and the caller then keeps the all-reduce path."""

    FLAG_WORDS = 16          # [2 phases][MARL_PEER_MAX_WORLD]

    def __init__(self, numel, dist, group, device):
        import ctypes as C
        import torch
        from . import _lib as L
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.numel, self._opened, self._base, self.ok = numel, [], None, False
        ok = self.world <= 8 and numel % 4 == 0 and numel <= (1 << 18)
        grad_bytes = (4 * (numel + 2) + 255) // 256 * 256
        # Every rank takes part in BOTH collectives below whatever happened locally (a rank that skipped the all_gather
        # while its peers sat in it would dead-lock NCCL): a failed rank contributes an all-zero handle, and the
        # decision is taken only after the MIN all-reduce.
        mine = torch.zeros(64, dtype=torch.uint8)
        if ok:
            try:
                base = C.c_void_p()
                L.call("marl_peer_alloc", grad_bytes + 4 * self.FLAG_WORDS, C.byref(base))
                self._base = base.value
                hbuf = C.create_string_buffer(64)
                L.call("marl_peer_export", self._base, hbuf)
                mine = torch.frombuffer(bytearray(hbuf.raw), dtype=torch.uint8).clone()
            except Exception:
                ok = False
        handles = [torch.empty(64, dtype=torch.uint8, device=device) for _ in range(self.world)]
        dist.all_gather(handles, mine.to(device), group=group)
        if ok and any(int(h.sum().item()) == 0 for h in handles):
            ok = False
        bases = [None] * self.world
        if ok:
            try:
                for r in range(self.world):
                    if r == self.rank:
                        bases[r] = self._base
                        continue
                    p = C.c_void_p()
                    L.call("marl_peer_open", bytes(handles[r].cpu().numpy().tobytes()), C.byref(p))
                    bases[r] = p.value
                    self._opened.append(p.value)
            except Exception:
                ok = False
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            self.close()
            return
        self.ok = True
        self.grad_full = _device_view(self._base, numel + 2, device)
        self.state = torch.zeros(2, dtype=torch.int32, device=device)        # [epoch, error]
        pg = L.PeerGroup()
        pg.world, pg.rank = self.world, self.rank
        for r in range(self.world):
            pg.grads[r], pg.flags[r] = bases[r], bases[r] + grad_bytes
        pg.epoch, pg.error = self.state.data_ptr(), self.state.data_ptr() + 4
        self.struct = pg

    def __del__(self):
        try:
            self.close()
        except Exception:      # interpreter shutdown: the driver reclaims the mappings
            pass

    def close(self):
        from . import _lib as L
        for p in self._opened:
            L.load().marl_peer_close(p)
        self._opened = []
        if self._base is not None:
            L.load().marl_peer_free(self._base)
            self._base = None


class _RawDevice:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def _device_view(ptr, n, device):
    import torch
    return torch.as_tensor(_RawDevice(ptr, n), device=device)
