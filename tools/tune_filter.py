"""Round-2 timing sweeps on the c4 shape: the filter form of K2 (ring depth, item size) against the kept-set scan, and the
column-chunk pipeline of the whole call.  usage: python tools/tune_filter.py [--k K] [--quick]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mammo_clip_dissect_b200 import _lib, similarity as sim  # noqa: E402


def timeit(fn, iters=8, warm=2):
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
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--k", type=int, default=32768)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    N, K, C = args.n, args.k, 763
    g = torch.Generator(device=dev).manual_seed(0)
    A = torch.randn(N, K, generator=g, device=dev)
    P = torch.randn(N, C, generator=g, device=dev) * 0.044
    gb = 4.0 * N * K / 1e9

    def topk():
        return sim._topk_int32(A, 100, dev)

    def say(label, ms, bytes_gb=gb):
        print("%-64s %8.3f ms  %7.0f GB/s" % (label, ms, bytes_gb / ms * 1e3), flush=True)

    Z = torch.zeros_like(A[:, : min(K, 8192)])
    Z[:200] = torch.arange(200, 0, -1, device=dev, dtype=torch.float32)[:, None]      # nothing passes after the first rows
    zb = 4.0 * N * Z.shape[1] / 1e9
    say("K2 filter form, nothing to append (streaming floor, %d cols)" % Z.shape[1], timeit(lambda: sim._topk_int32(Z, 100, dev)), zb)
    del Z
    _lib.set_tunable("topk_filter", 1)
    say("K2 kept-set scan (round 1)", timeit(topk))
    _lib.set_tunable("topk_filter", 0)
    for ns in ([4] if args.quick else [2, 3, 4, 5, 6]):
        _lib.set_tunable("filter_stages", ns)
        say("K2 filter form, per-warp ring of %d x 4 KB" % ns, timeit(topk))
    _lib.set_tunable("filter_stages", 0)
    for rows, stages in ((16, 2), (16, 3), (8, 2), (8, 3)):
        _lib.set_tunable("filter_order", rows)
        _lib.set_tunable("filter_stages", stages)
        say("K2 filter form, %d-row tiles, ring of %d" % (rows, stages), timeit(topk))
    _lib.set_tunable("filter_order", 0)
    _lib.set_tunable("filter_stages", 0)
    if not args.quick:
        for ct in (43, 85, 170, 340):
            _lib.set_tunable("filter_chunk_tiles", ct)
            say("K2 filter form, %d tiles per work item" % ct, timeit(topk))
        _lib.set_tunable("filter_chunk_tiles", 0)
    ramp = sim._device_ramp(sim._reference_ramp(100, 0.998, 0.97), 100, 0.998, 0.97, dev)
    S = sim.concept_probabilities(P, 10, dev)
    idx = topk()
    L = torch.empty((K, C), device=dev)
    say("K1b softmax", timeit(lambda: sim.concept_probabilities(P, 10, dev)), 2 * 4.0 * N * 768 / 1e9)
    say("K3 gather + log-sum", timeit(lambda: sim.log_sums(S, idx, ramp, 1e-7, out=L)), 100 * K * 768 * 4 / 1e9)
    say("K3b partials + finalize", timeit(lambda: sim.pmi_finalize(L, sim.lse_partials(L), K, 1.0)), 3 * 4.0 * K * C / 1e9)
    alg = (4.0 * N * K + 4.0 * N * C + 4.0 * K * C) / 1e9
    for q in ([-1, 1, 4] if args.quick else [-1, 1, 2, 3, 4, 5, 6, 8]):
        _lib.set_tunable("pipe_chunks", q)
        for ns in ([0] if args.quick else [0, 3]):
            _lib.set_tunable("filter_stages", ns)
            say("soft_wpmi, %d column chunk(s), ring %s" % (q, ns or "default"),
                timeit(lambda: sim.soft_wpmi(P, A, device=dev)), alg)
    _lib.set_tunable("pipe_chunks", 0)
    _lib.set_tunable("filter_stages", 0)
    for q in (2, 3, 4, 6):
        for pad in (19, 21, 24, 30, 40):
            _lib.set_tunable("pipe_chunks", q)
            _lib.set_tunable("accum_pad_kb", pad)
            say("soft_wpmi, %d column chunks, K3 padded by %d KB beside the scan" % (q, pad),
                timeit(lambda: sim.soft_wpmi(P, A, device=dev)), alg)
    _lib.set_tunable("pipe_chunks", 0)
    _lib.set_tunable("accum_pad_kb", 0)
    if not args.quick:
        for n2, k2 in ((10000, 9216), (20000, 8192), (40000, 8192)):
            A2 = torch.randn(n2, k2, generator=g, device=dev)
            for flt in (0, 1):
                _lib.set_tunable("topk_filter", flt)
                say("K2 at N=%d K=%d, %s" % (n2, k2, "kept-set scan" if flt else "filter form"),
                    timeit(lambda: sim._topk_int32(A2, 100, dev)), 4.0 * n2 * k2 / 1e9)
            del A2
    _lib.set_tunable("topk_filter", 1)
    _lib.set_tunable("pipe_chunks", -1)
    say("soft_wpmi, round-1 kernels (kept-set scan, one stream)", timeit(lambda: sim.soft_wpmi(P, A, device=dev)), alg)
    _lib.set_tunable("topk_filter", 0)
    _lib.set_tunable("pipe_chunks", 0)


if __name__ == "__main__":
    main()
