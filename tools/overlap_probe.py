"""Experiment: does K3 (gather + log-sum) run behind K2's scan when launched on another stream?
K2's scan is a single wave of persistent warps that takes (almost) all shared memory of every SM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mammo_clip_dissect_b200 import _lib, similarity as sim
dev = torch.device("cuda:0")
N, K, C = 100000, 32768, 763
g = torch.Generator(device=dev).manual_seed(0)
P = torch.randn(N, C, generator=g, device=dev) * 0.044
A = torch.randn(N, K, generator=g, device=dev)
S = sim.concept_probabilities(P, 10, dev)
idx = sim._topk_int32(A, 100, dev)
w = sim._reference_ramp(100, 0.998, 0.97).to(dev)
out = torch.empty((K, C), device=dev)
hi = torch.cuda.Stream(device=dev, priority=-1)
lo = torch.cuda.Stream(device=dev, priority=0)

def timed(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def topk():
    sim._topk_int32(A, 100, dev)
def accum():
    sim.log_sums(S, idx, w, 1e-7, out=out)
def both(order):
    cur = torch.cuda.current_stream(dev)
    ev = torch.cuda.Event(); ev.record(cur)
    hi.wait_event(ev); lo.wait_event(ev)
    def a():
        with torch.cuda.stream(hi): topk()
    def b():
        with torch.cuda.stream(lo): accum()
    (a(), b()) if order == 0 else (b(), a())
    cur.wait_stream(hi); cur.wait_stream(lo)

print("topk alone   %.3f ms" % timed(topk))
print("accum alone  %.3f ms" % timed(accum))
print("serial       %.3f ms" % timed(lambda: (topk(), accum())))
print("concurrent (scan stream first)  %.3f ms" % timed(lambda: both(0)))
print("concurrent (accum first)        %.3f ms" % timed(lambda: both(1)))
