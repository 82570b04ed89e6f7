#!/usr/bin/env python
"""Headline benchmark: neurons dissected per second with soft-WPMI over the 763-concept set.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[3], SURVEY.md section 8 "c4"): clip_feats P [100000, 763],
target_feats A [100000, 32768] per GPU, top_k = 100 -- synthetic (seeded randn), resident in HBM
when the timed region starts.  One step = one full soft_wpmi call over the rank's 32768 neurons.
N > 1: neuron-sharded (each rank owns 32768 columns of a [100000, 32768*N] activation matrix,
P replicated), LSE partials and score shards exchanged over NCCL -> weak scaling.

The JSON line carries `value` (device-resident), `e2e` (same call with pinned HOST inputs, H2D
and D2H inside the timed region), `roofline` of the dominant kernel stage (column top-k scan over
A), `roofline_path` for the whole call, `cpu_baseline` (the oracle port of the reference loop on
the host cores, bounded sample), `clocks` and `gpu_launches`.

--impl reference times the reference algorithm's CPU port (oracle/similarity_oracle.py loop form,
torch CPU ops) on a bounded sample of the same workload; /root/reference itself is not on the box.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

N_IMG, K_NEURONS, C_CONCEPTS, TOP_K = 100_000, 32_768, 763, 100
METRIC = "neurons dissected/sec (soft-WPMI, 763 concepts)"
UNIT = "neurons/s"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_note():
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {}


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples taken DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference loop on a bounded sample of the workload
# ------------------------------------------------------------------------------------------------
def cpu_sample_inputs(n_neurons, seed=2):
    g = torch.Generator().manual_seed(0)
    P = torch.randn(N_IMG, C_CONCEPTS, generator=g) * 0.044        # cosines of unit-norm random rows: sigma ~ 1/sqrt(512)
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(N_IMG, n_neurons, generator=g)
    return P, A


def time_cpu_port(P, A, threads):
    from oracle import similarity_oracle as orc
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    inds = torch.topk(A, dim=0, k=TOP_K)[1]              # the reference's own call (similarity.py:55); tie-free data
    orc.soft_wpmi(P, A, top_k=TOP_K, inds=inds)
    return time.perf_counter() - t0


def cpu_baseline(sample_one=2048, sample_all=256):
    """~10-20 s of CPU work: the loop port at 1 thread and at all host threads (the small torch ops of the
    per-neuron loop do not scale with threads -- SURVEY.md section 6 -- so the better of the two is reported)."""
    cores = os.cpu_count() or 1
    P, A = cpu_sample_inputs(sample_one)
    t_one = time_cpu_port(P, A, 1)
    t_all = time_cpu_port(P, A[:, :sample_all].contiguous(), cores)
    v_one, v_all = sample_one / t_one, sample_all / t_all
    torch.set_num_threads(cores)
    best_all = v_all >= v_one
    return {"value": round(max(v_all, v_one), 1), "unit": UNIT, "cores": cores if best_all else 1, "kind": "port",
            "sample": "oracle loop port of similarity.py:49-73 (torch CPU ops) on P[100000,763], top_k=100: "
                      "%d of the 32768 neuron columns with 1 thread -> %.1f neurons/s; %d columns with %d threads -> "
                      "%.1f neurons/s" % (sample_one, v_one, sample_all, cores, v_all)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    all_cores = os.cpu_count() or 1
    sample = 512
    P, A = cpu_sample_inputs(sample)
    # warm-up doubles as the thread-count probe: the per-neuron loop of tiny torch ops is usually
    # FASTER on one thread than on all of them; time the steps with whichever wins here
    small = A[:, :128].contiguous()
    probe = {}
    for w in range(max(args.warmup, 2)):
        th = all_cores if w % 2 == 0 else 1
        probe[th] = min(probe.get(th, 1e30), time_cpu_port(P, small, th))
    cores = min(probe, key=probe.get)
    times = [time_cpu_port(P, A, cores) for _ in range(args.steps)]
    ms = 1e3 * sum(times) / len(times)
    value = sample / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args.gpus), score_exchange="n/a (CPU reference arm)"),
            "cpu_baseline": {"value": round(value, 1), "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "each step: oracle loop port (torch CPU ops, %d threads) on P[100000,763] and "
                                       "%d of the 32768 neuron columns, top_k=100" % (cores, sample)},
            "e2e": {"value": round(value, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {"workload": "c4: soft_wpmi, clip_feats 100000x763, target_feats 100000x32768 per GPU, top_k=100, a=10, lam=1",
            "N_img": N_IMG, "K_per_gpu": K_NEURONS, "K_total": K_NEURONS * n_gpus, "C": C_CONCEPTS, "top_k": TOP_K,
            "parallelism": "neuron-sharded x%d" % n_gpus,
            "score_exchange": ("none (one GPU)" if n_gpus == 1 else
                               {"copy": "DMA pushes into peer-mapped [K_total,C] buffers, overlapped with the next call's scan",
                                "fused": "finalize kernel stores into all peers' [K_total,C] buffers",
                                "nccl": "NCCL all_gather"}.get(os.environ.get("MCD_EXCHANGE", "copy"), "?")),
            "l2_policy": "inputs (13.4 GB per GPU) exceed the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    from mammo_clip_dissect_b200 import _lib, similarity
    from mammo_clip_dissect_b200 import distributed as mdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.lib().mcd_device_check(), "mcd_device_check")

    # ---- synthetic inputs, resident in HBM --------------------------------------------------
    g = torch.Generator(device=dev).manual_seed(0)
    P = torch.randn(N_IMG, C_CONCEPTS, generator=g, device=dev) * 0.044
    g = torch.Generator(device=dev).manual_seed(2 + 1000 * rank)
    A = torch.randn(N_IMG, K_NEURONS, generator=g, device=dev)
    shard_sizes = [K_NEURONS] * world
    backend = mdist.CudaBackend(dev) if world > 1 else None
    # N > 1: the [K_total, C] scores are exchanged through symmetric (peer-mapped) memory.  "copy": DMA pushes on side
    # streams, the exchange of call i runs behind the column scan of call i+1 (double-buffered, at most one exchange in
    # flight, the last one is waited for inside the timed region); "fused": the finalize kernel stores into all peers;
    # "nccl": torch.distributed all_gather.
    xmode = os.environ.get("MCD_EXCHANGE", "copy")
    exchange = None
    if world > 1 and xmode != "nccl":
        # every rank must take the same path: agree on whether the symmetric-memory rendezvous worked everywhere
        import torch.distributed as dist
        try:
            exchange = mdist.PeerScoreExchange(shard_sizes, C_CONCEPTS, dev, mode=xmode)
            ok = 1
        except Exception as exc:                                   # no peer access / no symmetric memory on this box
            sys.stderr.write("rank %d: peer score exchange unavailable (%s); using NCCL all_gather\n" % (rank, exc))
            ok = 0
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            exchange, xmode = None, "nccl"
        os.environ["MCD_EXCHANGE"] = xmode                         # what workload_config() reports
    in_flight = []

    def step(P_in, A_in, overlap=False):
        if world == 1:
            return similarity.soft_wpmi(P_in, A_in, top_k=TOP_K, device=dev)
        h = mdist.soft_wpmi_sharded(P_in, A_in, shard_sizes, top_k=TOP_K, device=dev, backend=backend,
                                    exchange=exchange, wait=not (overlap and exchange is not None))
        if overlap and exchange is not None:
            drain()
            in_flight.append(h)
        return h

    def drain():
        while in_flight:
            in_flight.pop().wait()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(max(args.warmup, 3)):
        out = step(P, A, overlap=True)
    drain()
    barrier()

    # ---- device-resident timing: K steps between two events ----------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    similarity.PROFILE = []
    n0 = _lib.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step(P, A, overlap=True)
    drain()
    e1.record()
    barrier()
    launches = _lib.launch_count() - n0
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    stage_ms = similarity.profile_summary()
    similarity.PROFILE = None
    clocks = sampler.stop() if rank == 0 else None
    value = K_NEURONS * world / (ms_step / 1e3)

    # ---- end to end: pinned host inputs -> public API -> host result -----------------------
    e2e = None
    try:
        P_h = torch.empty(P.shape, dtype=P.dtype, pin_memory=True).copy_(P)
        A_h = torch.empty(A.shape, dtype=A.dtype, pin_memory=True).copy_(A)
        del A
        torch.cuda.empty_cache()
        res_h = torch.empty((K_NEURONS * world if world > 1 else K_NEURONS, C_CONCEPTS), dtype=torch.float32,
                            pin_memory=True)
        e2e_steps = max(2, min(args.steps, 5))
        for _ in range(2):
            res_h.copy_(step(P_h, A_h), non_blocking=True)
        barrier()
        e0.record()
        for _ in range(e2e_steps):
            res_h.copy_(step(P_h, A_h), non_blocking=True)
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1) / e2e_steps)
        e2e = {"value": round(K_NEURONS * world / (ms_e2e / 1e3), 1), "unit": UNIT,
               "h2d_bytes_per_step": int(P_h.numel() * 4 + A_h.numel() * 4), "d2h_bytes_per_step": int(res_h.numel() * 4),
               "ms_per_step": round(ms_e2e, 3), "steps": e2e_steps,
               "api": ("similarity.soft_wpmi" if world == 1 else "distributed.soft_wpmi_sharded") +
                      "(P_host_pinned, A_host_pinned, device='cuda') -> score matrix copied to pinned host memory"}
    except RuntimeError as exc:                      # e.g. the box cannot pin 13.4 GB
        e2e = {"value": None, "unit": UNIT, "error": str(exc)[:200]}

    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant stage + whole path ----------------------------------------
    peak, peak_src = peaks()
    topk_ms = stage_ms.get("topk_cols")
    alg_topk = 4.0 * N_IMG * K_NEURONS                                  # read A once
    alg_path = 4.0 * N_IMG * K_NEURONS + 4.0 * N_IMG * C_CONCEPTS + 4.0 * K_NEURONS * C_CONCEPTS
    tr = traffic_note()
    roof = None
    if topk_ms:
        ach = alg_topk / (topk_ms / 1e3) / 1e9
        roof = {"bound": "hbm", "kernel": "topk_scan_kernel; timed as the whole topk_cols stage (sample_tilemax + sample_select + scan + redo + finish) with CUDA events, so the fraction is a lower bound for the scan kernel itself",
                "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                "traffic": tr.get("topk_scan_kernel"), "peak_source": peak_src, "ms_per_launch": round(topk_ms, 4),
                "algorithmic_bytes_per_launch": alg_topk}
    ach_path = alg_path / (ms_step / 1e3) / 1e9
    line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(world),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
            "roofline_path": {"bound": "hbm", "achieved": round(ach_path, 1), "peak": peak, "unit": "GB/s",
                              "frac": round(ach_path / peak, 4), "algorithmic_bytes_per_step": alg_path,
                              "note": "B_alg = 4NK + 4NC + 4KC per GPU (SURVEY.md 8d); gather re-reads not counted"},
            "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()}}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
