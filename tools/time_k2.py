"""Quick timing of the K2 stage and the whole call at c4 for a few ring / item settings.  usage: python tools/time_k2.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mammo_clip_dissect_b200 import _lib, similarity as sim
from tools.tune_filter import timeit

dev = torch.device("cuda:0")
N, K, C = 100000, int(os.environ.get("PK", 32768)), 763
g = torch.Generator(device=dev).manual_seed(0)
A = torch.randn(N, K, generator=g, device=dev)
P = torch.randn(N, C, generator=g, device=dev) * 0.044
gb = 4.0 * N * K / 1e9
for ns in (2,):
    for ct in (0, 12, 16, 24, 31, 48):
        _lib.set_tunable("filter_stages", ns)
        _lib.set_tunable("filter_chunk_tiles", ct)
        ms = timeit(lambda: sim._topk_int32(A, 100, dev))
        print("K2 stage: ring %d, %3d tiles per item (0 = default): %7.3f ms  %6.0f GB/s" % (ns, ct, ms, gb / ms * 1e3), flush=True)
_lib.set_tunable("filter_stages", 0)
_lib.set_tunable("filter_chunk_tiles", 0)
ms = timeit(lambda: sim.soft_wpmi(P, A, device=dev))
print("soft_wpmi: %.3f ms  (%.4f of 6542.7 GB/s)" % (ms, (4.0 * N * K + 4.0 * N * C + 4.0 * K * C) / 1e9 / ms * 1e3 / 6542.7), flush=True)
