"""K2 filter scan with nothing to append after the first 200 rows (the feed + compare floor of the real kernel) against
the same call on random data, full width.  usage: python tools/time_k2_floor.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mammo_clip_dissect_b200 import _lib, similarity as sim
from tools.tune_filter import timeit

dev = torch.device("cuda:0")
N, K = 100000, 32768
gb = 4.0 * N * K / 1e9
A = torch.zeros(N, K, device=dev)
A[:200] = torch.arange(200, 0, -1, device=dev, dtype=torch.float32)[:, None]
for ns in (2, 3):
    _lib.set_tunable("filter_stages", ns)
    ms = timeit(lambda: sim._topk_int32(A, 100, dev))
    print("nothing to append, ring %d: stage %7.3f ms  %6.0f GB/s" % (ns, ms, gb / ms * 1e3), flush=True)
g = torch.Generator(device=dev).manual_seed(0)
A.normal_(generator=g)
for ns in (2, 3):
    _lib.set_tunable("filter_stages", ns)
    ms = timeit(lambda: sim._topk_int32(A, 100, dev))
    print("randn,             ring %d: stage %7.3f ms  %6.0f GB/s" % (ns, ms, gb / ms * 1e3), flush=True)
_lib.set_tunable("filter_stages", 0)
