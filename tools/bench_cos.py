"""cos / cos^3 timing at the c5 shape (N = 5000, K = 512) and at a long-contraction shape: tensor-core path against the
exact CUDA-core kernel.  usage: python tools/bench_cos.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mammo_clip_dissect_b200 import _lib, similarity as sim


def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


dev = torch.device("cuda:0")
for N, K in ((5000, 512), (10000, 9216), (100000, 4096)):
    g = torch.Generator(device=dev).manual_seed(0)
    P = torch.randn(N, 763, generator=g, device=dev) * 0.05 + 0.2
    A = torch.randn(N, K, generator=g, device=dev)
    fl = 2.0 * N * K * 763
    for name, fn in (("cos_similarity_cubed", sim.cos_similarity_cubed), ("cos_similarity", sim.cos_similarity)):
        for var in (0, 1):
            if var == 1 and N * K > 1e8:
                continue
            _lib.set_tunable("gemm_variant", var)
            ms = timeit(lambda: fn(P, A, device=dev))
            print("%-22s N=%6d K=%5d  %-9s %9.3f ms  %7.1f TFLOP/s (2NKC)" % (name, N, K, sim.last_cos_path(), ms, fl / ms / 1e9), flush=True)
    _lib.set_tunable("gemm_variant", 0)
