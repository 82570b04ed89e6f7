"""Per-call times of the path on the reference's real shapes (BASELINE.json configs c1, c2, c3, c5): synthetic inputs,
device-resident, CUDA events, median of 20 calls.  Not the bench line (that is c4, bench.py) -- a table for DESIGN.md."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mammo_clip_dissect_b200 import similarity as sim
dev = torch.device("cuda:0")
C = 763


def med(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def inputs(N, K, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    return torch.randn(N, C, generator=g, device=dev) * 0.044, torch.randn(N, K, generator=g, device=dev)


print("| config | call | ms / call | neurons / s |")
print("|---|---|---|---|")
P, A = inputs(2000, 2048)
t = med(lambda: sim.soft_wpmi(P, A, device=dev))
print("| c1: N=2000, K=2048 | soft_wpmi(top_k=100) | %.3f | %.2f M |" % (t, 2048 / t / 1e3))
# c2: EfficientNet-B5 block widths, N = 5000, 39 calls
widths = [24] * 3 + [40] * 5 + [64] * 5 + [128] * 7 + [176] * 7 + [304] * 9 + [512] * 3
P, _ = inputs(5000, 8)
As = {w: inputs(5000, w, seed=w)[1] for w in set(widths)}
def c2():
    for w in widths:
        sim.soft_wpmi(P, As[w], device=dev)
t = med(c2, iters=5)
print("| c2: N=5000, 39 layers, sum K = %d | 39 x soft_wpmi | %.3f (all layers) | %.2f M |" % (sum(widths), t, sum(widths) / t / 1e3))
layers = [As[w] for w in widths]
t = med(lambda: sim.soft_wpmi_layers(P, layers, device=dev), iters=10)
print("| c2, same 39 layers | 1 x soft_wpmi_layers | %.3f (all layers) | %.2f M |" % (t, sum(widths) / t / 1e3))
P, A = inputs(10000, 768)
blocks = [inputs(10000, 768, seed=100 + i)[1] for i in range(12)]
t = med(lambda: [sim.soft_wpmi(P, b, device=dev) for b in blocks], iters=5)
print("| c3: N=10000, 12 ViT blocks x 768 | 12 x soft_wpmi | %.3f (all blocks) | %.2f M |" % (t, 12 * 768 / t / 1e3))
t = med(lambda: sim.soft_wpmi_layers(P, blocks, device=dev), iters=5)
print("| c3, same 12 blocks | 1 x soft_wpmi_layers | %.3f (all blocks) | %.2f M |" % (t, 12 * 768 / t / 1e3))
t = med(lambda: sim.soft_wpmi(P, A, device=dev))
print("| c3: N=10000, K=768 (one of 12 ViT blocks) | soft_wpmi | %.3f | %.2f M |" % (t, 768 / t / 1e3))
P, A = inputs(5000, 512)
for name, fn in (("wpmi(top_k=28)", lambda: sim.wpmi(P, A, device=dev)),
                 ("soft_wpmi(top_k=10)", lambda: sim.soft_wpmi(P, A, top_k=10, device=dev)),
                 ("soft_wpmi(top_k=200)", lambda: sim.soft_wpmi(P, A, top_k=200, device=dev)),
                 ("cos_similarity_cubed", lambda: sim.cos_similarity_cubed(P, A, device=dev)),
                 ("cos_similarity", lambda: sim.cos_similarity(P, A, device=dev)),
                 ("rank_reorder", lambda: sim.rank_reorder(P, A, device=dev))):
    t = med(fn, iters=10)
    print("| c5: N=5000, K=512 | %s | %.3f | %.2f M |" % (name, t, 512 / t / 1e3))
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    sim.rank_reorder(P, A, device=dev)
torch.cuda.synchronize()
print("| c5: N=5000, K=512 | rank_reorder, wall clock incl. the host's RNG replay | %.3f | |" % ((time.perf_counter() - t0) / 5 * 1e3))
P, A = inputs(100000, 64)
t = med(lambda: sim.rank_reorder(P, A, device=dev), iters=3)
print("| N=100000, K=64 (top_n = 5000) | rank_reorder | %.3f | %.4f M |" % (t, 64 / t / 1e3))
