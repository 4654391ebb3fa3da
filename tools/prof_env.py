"""A few env launches at 2^24 envs (the command profiled by ncu for the HBM roofline)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from marl_b200.env.single_state_matrix_game import BatchedMatrixGame
n = 1 << 24
env = BatchedMatrixGame([[8, -12, -12], [-12, 0, 0], [-12, 0, 0]], n)
a = torch.randint(0, 3, (n, 2), device="cuda")
for _ in range(4):
    env.step(a)
torch.cuda.synchronize()
