"""Kept-set scan against the (forced) filter form below the automatic crossover.  usage: python tools/time_k2_crossover.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mammo_clip_dissect_b200 import _lib, similarity as sim
from tools.tune_filter import timeit
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
for n, k in ((8192, 9216), (8500, 9216), (9000, 9216), (10000, 9216), (12000, 9216), (16000, 8192), (20000, 8192), (24000, 8192), (40000, 8192), (10000, 768), (20000, 512)):
    A = torch.randn(n, k, generator=g, device=dev)
    ref = None
    for flt, name in ((0, "automatic"), (1, "kept-set scan"), (2, "filter form")):
        _lib.set_tunable("topk_filter", flt)
        ms = timeit(lambda: sim._topk_int32(A, 100, dev))
        idx = sim._topk_int32(A, 100, dev)
        same = "" if ref is None else ("  same indices: %s" % bool(torch.equal(idx, ref)))
        ref = idx if ref is None else ref
        print("N=%6d K=%5d  %-14s %7.3f ms%s" % (n, k, name, ms, same), flush=True)
    del A
_lib.set_tunable("topk_filter", 0)
