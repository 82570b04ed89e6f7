"""torchrun --nproc-per-node G tools/check_multi_gpu.py : the neuron-sharded NCCL path must reproduce the
single-GPU soft_wpmi bit for bit (fixed 256-neuron LSE blocks combined in global block order)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from mammo_clip_dissect_b200 import distributed as mdist, similarity

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g = torch.Generator().manual_seed(0)
N, K, C = 4000, 2304, 763
P = (torch.randn(N, C, generator=g) * 0.05).to(dev)
A = torch.randn(N, K, generator=g).to(dev)
b = mdist.shard_bounds(K, world)
sizes = [b[i + 1] - b[i] for i in range(world)]
full = mdist.soft_wpmi_sharded(P, A[:, b[rank]:b[rank + 1]].contiguous(), sizes, device=dev)
single = similarity.soft_wpmi(P, A, device=dev)
same = torch.equal(full, single)
w = mdist.wpmi_sharded(P, A[:, b[rank]:b[rank + 1]].contiguous(), sizes, device=dev)
same_w = torch.equal(w, similarity.wpmi(P, A, device=dev))
same_x = True
for mode in ("copy", "fused"):
    ex = mdist.PeerScoreExchange(sizes, C, dev, mode=mode)
    for it in range(3):                      # more calls than buffers: the round-robin reuse is exercised
        got = mdist.soft_wpmi_sharded(P, A[:, b[rank]:b[rank + 1]].contiguous(), sizes, device=dev, exchange=ex)
        same_x = same_x and torch.equal(got, single)
    h = mdist.soft_wpmi_sharded(P, A[:, b[rank]:b[rank + 1]].contiguous(), sizes, device=dev, exchange=ex, wait=False)
    same_x = same_x and torch.equal(h.wait(), single)
    if rank == 0:
        print("exchange mode %s: %s" % (mode, "bit-identical" if same_x else "MISMATCH"), flush=True)
flag = torch.tensor([int(same and same_w and same_x)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("multi-gpu parity (world=%d): %s" % (world, "bit-identical" if flag.item() else "MISMATCH"), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
