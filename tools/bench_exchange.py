"""Multi-GPU box: time the three score-exchange variants at the bench shape (K_g = 32768 neurons per rank, C = 763).
launch: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/bench_exchange.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from mammo_clip_dissect_b200 import distributed as mdist, similarity as sim

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
Kg, C = 32768, 763
sizes = [Kg] * world
L0 = torch.randn(Kg, C, device=dev) - 400.0
part = sim.lse_partials(L0)
part_all = mdist._all_gather_var(part, [Kg // 256] * world, None)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def nccl():
    L = L0.clone()
    loc = sim.pmi_finalize(L, part_all, Kg * world, 1.0)[0]
    return mdist._all_gather_var(loc, sizes, None)


res = {"clone only": timeit(lambda: L0.clone()), "nccl finalize + all_gather": timeit(nccl)}
for mode in ("copy", "fused"):
    ex = mdist.PeerScoreExchange(sizes, C, dev, mode=mode)
    res["exchange " + mode] = timeit(lambda: ex.exchange(L0.clone(), part_all, 1.0, sim))
if rank == 0:
    gb = (world - 1) * Kg * C * 4 / 1e9
    for k, v in res.items():
        print("%-28s %.3f ms   (%.0f GB/s into each GPU)" % (k, v, gb / v * 1e3), flush=True)
dist.destroy_process_group()
