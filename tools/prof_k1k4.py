"""K1 (image x text matrix on tcgen05) and K4 (pooling, both memory orders) once each at their headline shapes, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mammo_clip_dissect_b200 import features, hooks, similarity as sim
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
I = torch.randn(100000, 512, generator=g, device=dev)
T = torch.randn(763, 512, generator=g, device=dev)
for _ in range(2):
    P, S = features.similarity_matrix(I, T, device=dev, softmax_scale=10)
x = torch.randn(32, 24, 760, 456, generator=g, device=dev)           # EfficientNet-B5 stem-stage block, batch 32 (2.66 GB)
xl = torch.randn(32, 128, 95, 57, generator=g, device=dev).contiguous(memory_format=torch.channels_last)
for _ in range(2):
    a = hooks.pool_nchw(x, "avg")
    b = hooks.pool_nchw(xl, "avg")
A = torch.randn(10000, 9216, generator=g, device=dev)
Pc = torch.randn(10000, 763, generator=g, device=dev) * 0.05 + 0.2
for _ in range(2):
    c = sim.cos_similarity_cubed(Pc, A, device=dev)
torch.cuda.synchronize()
print("ok", features.last_gemm_path(), sim.last_cos_path())
