"""Small driver for ncu: one column-top-k call on a one-wave problem (148 column blocks)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mammo_clip_dissect_b200 import _lib, similarity as sim
dev = torch.device("cuda:0")
N, K = int(os.environ.get("PN", 100000)), int(os.environ.get("PK", 148 * 128))
A = torch.randn(N, K, device=dev)
_lib.set_tunable("topk_splits", int(os.environ.get("PSPLITS", 1)))
for _ in range(3):
    idx = sim._topk_int32(A, 100, dev)
torch.cuda.synchronize()
print("ok", idx.shape)
