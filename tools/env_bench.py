"""env-steps/s of the batched matrix game kernel at several sizes."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from marl_b200.env.single_state_matrix_game import BatchedMatrixGame
for n in (4096, 1 << 16, 1 << 20, 1 << 24, (1 << 24) + 12345):
    for dt in (torch.int64, torch.int32):
        env = BatchedMatrixGame([[8, -12, -12], [-12, 0, 0], [-12, 0, 0]], n)
        a = torch.randint(0, 3, (n, 2), device="cuda", dtype=dt)
        for _ in range(5): env.step(a)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        it = 50
        e0.record()
        for _ in range(it): env.step(a)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / it * 1e3
        byt = (124 + (16 if dt == torch.int64 else 8)) * n
        print(f"n={n:9d} {str(dt):12s} {us:9.1f} us/launch  {n/us/1e3:8.2f} G env-steps/s  {byt/us/1e3:8.1f} GB/s")
