"""MCD_DEBUG_SYNC=1 python tools/debug_filter.py N K k : one column top-k call with a synchronise after every launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mammo_clip_dissect_b200 import _lib, similarity as sim
N, K, k = (int(x) for x in sys.argv[1:4])
for kv in sys.argv[4:]:
    name, v = kv.split("=")
    _lib.set_tunable(name, int(v))
A = torch.randn(N, K, generator=torch.Generator().manual_seed(1)).cuda()
n0 = _lib.launch_count()
try:
    idx = sim.topk_cols(A, k, device="cuda:0")
    torch.cuda.synchronize()
    ref = torch.topk(A, k, dim=0)[1]
    print("ok launches", _lib.launch_count() - n0, "match", bool(torch.equal(idx, ref)))
except Exception as exc:
    print("FAILED after", _lib.launch_count() - n0, "launches:", str(exc)[:300])
