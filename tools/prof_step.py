"""One soft_wpmi call at the bench shape (c4) for ncu: launch list and per-kernel captures.
env: PK = neurons (32768), PITERS = calls (3), PPIPE = pipe_chunks tunable (-1: one stream), PSTAGES = filter ring depth."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mammo_clip_dissect_b200 import _lib, similarity as sim
dev = torch.device("cuda:0")
N, K, C = 100000, int(os.environ.get("PK", 32768)), 763
_lib.set_tunable("pipe_chunks", int(os.environ.get("PPIPE", -1)))
_lib.set_tunable("filter_stages", int(os.environ.get("PSTAGES", 0)))
g = torch.Generator(device=dev).manual_seed(0)
P = torch.randn(N, C, generator=g, device=dev) * 0.044
A = torch.randn(N, K, generator=g, device=dev)
for _ in range(int(os.environ.get("PITERS", 3))):
    out = sim.soft_wpmi(P, A, device=dev)
torch.cuda.synchronize()
print("ok", tuple(out.shape), float(out[0, 0]))
