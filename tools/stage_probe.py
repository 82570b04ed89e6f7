"""Experiment: per-warp time of the column scan vs rows per warp and co-resident warps (one wave each)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mammo_clip_dissect_b200 import _lib, similarity as sim
dev = torch.device("cuda:0")
def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
N = 100000
for occ, splits in ((7, 1), (7, 2), (6, 2), (4, 2), (6, 3), (7, 4)):
    K = (148 * occ // splits) * 32
    A = torch.randn(N, K, device=dev)
    gb = 4.0 * N * K / 1e9
    _lib.set_tunable("topk_splits", splits)
    _lib.set_tunable("topk_pre", 3)
    ms = timeit(lambda: sim._topk_int32(A, 100, dev))
    print("warps/SM %d  splits %d  K %d: %.3f ms  %.0f GB/s" % (occ, splits, K, ms, gb / ms * 1e3), flush=True)
    del A
