"""Probe (multi-GPU box): torch symmetric memory rendezvous, peer views, copy-engine push bandwidth, barrier cost.
launch: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/symm_probe.py"""
import os
import time
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    rows, C = 32768, 763
    full = symm_mem.empty((world * rows, C), dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(full, dist.group.WORLD)
    if rank == 0:
        print("rendezvous ok: world", hdl.world_size, "multicast", hdl.has_multicast_support, "ptrs", [hex(p) for p in hdl.buffer_ptrs], flush=True)
    local = torch.full((rows, C), float(rank + 1), device=dev)
    peers = [hdl.get_buffer(p, (world * rows, C), torch.float32) for p in range(world)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]

    def push():
        ev = torch.cuda.Event()
        ev.record()
        for i in range(world):
            p = (rank + i) % world
            st = streams[i]
            st.wait_event(ev)
            with torch.cuda.stream(st):
                peers[p][rank * rows:(rank + 1) * rows].copy_(local, non_blocking=True)
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)

    hdl.barrier(channel=0)
    push()
    hdl.barrier(channel=0)
    torch.cuda.synchronize()
    ok = all(bool((full[p * rows:(p + 1) * rows] == float(p + 1)).all()) for p in range(world))
    print("rank", rank, "content ok", ok, flush=True)
    for name, fn in (("push (copy engines, %d peers)" % world, push), ("barrier", lambda: hdl.barrier(channel=0)),
                     ("nccl all_gather_into_tensor", None)):
        if fn is None:
            out = torch.empty((world * rows, C), device=dev)
            fn = lambda: dist.all_gather_into_tensor(out, local)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        if rank == 0:
            gb = (world - 1) * rows * C * 4 / 1e9
            print("%s: %.3f ms  (%.0f GB/s in per GPU)" % (name, ms, gb / ms * 1e3), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
