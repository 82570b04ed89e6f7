"""rank_reorder at c5 for an ncu launch list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mammo_clip_dissect_b200 import similarity as sim
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
P = torch.randn(5000, 763, generator=g, device=dev) * 0.044
A = torch.randn(5000, 512, generator=g, device=dev)
for _ in range(3):
    out = sim.rank_reorder(P, A, device=dev)
    o2 = sim.cos_similarity_cubed(P, A, device=dev)
torch.cuda.synchronize()
print("ok")
