"""Kernel-level timing sweeps on the c4 shape (not the bench; used while tuning).
usage: python tools/tune.py [topk|accum|all] [--n N] [--k K]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mammo_clip_dissect_b200 import _lib, similarity as sim  # noqa: E402


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="?", default="all")
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--k", type=int, default=32768)
    ap.add_argument("--topk", type=int, default=100)
    ap.add_argument("--splits", default="0,1,2,3,4,6,8")
    ap.add_argument("--variants", default="0,128,112,64")
    ap.add_argument("--tiles", default="0,384,256,192,128")
    ap.add_argument("--data", default="randn")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    N, K, C = args.n, args.k, 763
    g = torch.Generator(device=dev).manual_seed(0)
    if args.what == "pool":
        N, K = 8, 8
    A = torch.randn(N, K, generator=g, device=dev)
    if args.data == "zeros":
        A.zero_()
    elif args.data == "desc":
        A = -torch.arange(N, device=dev, dtype=torch.float32)[:, None].expand(N, K).contiguous()
    P = torch.randn(N, C, generator=g, device=dev) * 0.044
    gb = 4.0 * N * K / 1e9
    if args.what in ("topk", "all"):
        for pre in (0, 2):
            _lib.set_tunable("topk_pre", pre)
            ms = timeit(lambda: sim._topk_int32(A, args.topk, dev))
            print("topk pre-threshold %s: %.3f ms  %.0f GB/s" % ("on" if pre == 0 else "off", ms, gb / ms * 1e3), flush=True)
        _lib.set_tunable("topk_pre", 0)
        for var in [int(x) for x in args.variants.split(",")]:
            _lib.set_tunable("topk_cols", var)
            for s in [int(x) for x in args.splits.split(",")]:
                _lib.set_tunable("topk_splits", s)
                ms = timeit(lambda: sim._topk_int32(A, args.topk, dev))
                print("topk cols=%d splits=%d: %.3f ms  %.0f GB/s" % (var, s, ms, gb / ms * 1e3), flush=True)
        _lib.set_tunable("topk_cols", 0)
        _lib.set_tunable("topk_splits", 0)
    if args.what in ("accum", "all"):
        S = sim.concept_probabilities(P, 10, dev)
        idx = sim._topk_int32(A, args.topk, dev)
        w = sim._reference_ramp(args.topk, 0.998, 0.97).to(dev)
        out = torch.empty((K, C), device=dev)
        for hint in (0,):
            for t in [int(x) for x in args.tiles.split(",")]:
                _lib.set_tunable("accum_tile", t)
                ms = timeit(lambda: sim.log_sums(S, idx, w, 1e-7, out=out))
                print("accum tile=%d hint=%d: %.3f ms  gather %.0f GB/s" % (t, hint, ms, args.topk * K * 768 * 4 / 1e9 / ms * 1e3), flush=True)
        _lib.set_tunable("accum_unroll", 0)
        _lib.set_tunable("accum_tile", 0)
        print("softmax: %.3f ms" % timeit(lambda: sim.concept_probabilities(P, 10, dev)))
    if args.what in ("pool",):
        from mammo_clip_dissect_b200.hooks import pool_nchw
        # EfficientNet-B5 @1520x912 hooked block shapes, batch 4 (reference loader) and 32
        for B in (4, 32):
            tot_ms = tot_b = 0.0
            for (Cc, H, W, reps) in ((24, 760, 456, 3), (40, 380, 228, 5), (64, 190, 114, 5), (128, 95, 57, 7),
                                     (176, 95, 57, 7), (304, 48, 29, 9), (512, 48, 29, 3)):
                x = torch.randn(B, Cc, H, W, device=dev)
                ms = timeit(lambda: pool_nchw(x, "avg"), iters=10)
                ref = timeit(lambda: x.mean(dim=[2, 3]), iters=10)
                gbytes = x.numel() * 4 / 1e9
                print("pool B=%d C=%d %dx%d: %.4f ms  %.0f GB/s   (torch mean: %.4f ms)" % (B, Cc, H, W, ms, gbytes / ms * 1e3, ref), flush=True)
                tot_ms += ms * reps
                tot_b += gbytes * reps
                del x
            print("pool all 39 blocks, B=%d: %.3f ms for %.2f GB -> %.0f GB/s" % (B, tot_ms, tot_b, tot_b / tot_ms * 1e3), flush=True)
        return
    if args.what in ("gemm", "all"):
        from mammo_clip_dissect_b200 import features
        I = torch.randn(N, 512, generator=g, device=dev)
        T = torch.randn(C, 512, generator=g, device=dev)
        fl = 2.0 * N * C * 512
        for var in (0, 3, 1):
            _lib.set_tunable("gemm_variant", var)
            ms = timeit(lambda: features.similarity_matrix(I, T, device=dev))
            print("K1 I.T^T variant=%d (%s): %.3f ms  %.1f TFLOP/s (algorithmic 2NCD)" % (var, {0: "tcgen05 3xTF32 + stand-alone softmax", 3: "tcgen05 3xTF32 band kernel, softmax fused in the epilogue"}.get(var, "fp32 FFMA"), ms, fl / ms / 1e9), flush=True)
            ms = timeit(lambda: features.similarity_matrix(I, T, device=dev, softmax_scale=10))
            print("   + softmax: %.3f ms" % ms, flush=True)
        _lib.set_tunable("gemm_variant", 0)
    if args.what in ("accum", "all"):
        L = out.clone()
        print("lse+finalize: %.3f ms" % timeit(lambda: sim.pmi_finalize(L, sim.lse_partials(L), K, 1.0)))


if __name__ == "__main__":
    main()
