"""K1 timing at the headline probe count: I [100000, 512] x T [763, 512]^T (+ softmax), CUDA events, per path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mammo_clip_dissect_b200 import _lib, features

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
I = torch.randn(100000, 512, generator=g, device=dev)
T = torch.randn(763, 512, generator=g, device=dev)


def timeit(fn, iters=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


ref = (torch.nn.functional.normalize(I.double(), dim=1) @ torch.nn.functional.normalize(T.double(), dim=1).T)
for var, name in ((0, "streaming kernel (default)"), (2, "one CTA per tile"), (4, "streaming, row pairs fused"),
                  (3, "band kernel, rescale fused"), (1, "fp32 CUDA cores")):
    _lib.set_tunable("gemm_variant", var)
    for scale in (None, 10.0):
        ms = timeit(lambda: features.similarity_matrix(I, T, device=dev, softmax_scale=scale))
        out = features.similarity_matrix(I, T, device=dev, softmax_scale=scale)
        P = out[0] if scale else out
        err = (P.double() - ref).abs().max().item()
        line = "%-28s softmax=%-5s %-8s %8.3f ms   max|P - fp64| %.2e" % (name, scale, features.last_gemm_path(), ms, err)
        if scale:
            S64 = torch.softmax(scale * ref, dim=1)
            line += "   max rel |S - fp64| %.2e" % ((out[1].double() - S64).abs() / S64).max().item()
        print(line, flush=True)
_lib.set_tunable("gemm_variant", 0)
# measurement aid: the same pipeline with one MMA term instead of three (plain TF32: wrong digits, same operand traffic
# through TMA, a third of the tensor-core work and of its shared-memory operand reads)
_lib.set_tunable("gemm_debug_terms", 1)
print("streaming kernel, 1 of 3 MMA terms: %.3f ms" % timeit(lambda: features.similarity_matrix(I, T, device=dev)), flush=True)
_lib.set_tunable("gemm_debug_terms", 0)
for tiles in (1, 2, 3):
    _lib.set_tunable("gemm_tiles_per_cta", tiles)
    print("streaming kernel, %d column tiles per CTA: %.3f ms" % (tiles, timeit(lambda: features.similarity_matrix(I, T, device=dev))), flush=True)
_lib.set_tunable("gemm_tiles_per_cta", 0)
