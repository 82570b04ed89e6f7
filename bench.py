#!/usr/bin/env python
"""Headline benchmark: neurons dissected per second with soft-WPMI over the 763-concept set.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[3], SURVEY.md section 8 "c4"): clip_feats P [100000, 763], target_feats
A [100000, 32768], top_k = 100 -- synthetic (seeded randn), resident in HBM when the timed region starts.
One step = one full soft_wpmi call over all 32768 neurons.

N > 1 (one rank per GPU, NCCL): the SAME 32768 neurons are sharded by column over the ranks (strong scaling, the
configuration BASELINE names: rank r owns columns [r*32768/N, (r+1)*32768/N); the matrix is generated as eight
4096-column pieces with seeds 2 + 1000*piece, so every N scores identical data).  P is replicated and every rank runs the
softmax itself; the 256-neuron LSE partials are all-gathered and every rank ends up with its own finalized shard
(`value`), and -- separately timed, `with_score_exchange` -- with the full [32768, 763] matrix.  The weak-scaling curve
(32768 neurons PER rank, round 1's headline) is measured in the same run and reported under `weak`.

The JSON line carries `value` (device-resident, the public call), `e2e` (same call with pinned HOST inputs, H2D and
D2H inside the timed region), `stage_ms` + `roofline` (a separate staged pass: per-stage CUDA events; the dominant
stage is the column top-k scan over A), `roofline_path` for the whole call, `parity` (sampled columns of the benched
problem against the CPU oracle, outside the timed region; N > 1: sharded bits == single-GPU bits for every exchange
mode on a small problem), `cpu_baseline`, `reference_gpu` (the reference's own device='cuda' loop on a neuron slice,
when the reference is staged under baseline/_ref), `clocks` and `gpu_launches`.

--impl reference times the reference's CPU implementation of the path on a bounded sample of the same workload: the
unmodified concept_vit/similarity.py when tools/stage_reference.py has staged it under baseline/_ref (kind
"reference"), else the oracle's loop port (kind "port").
"""
import argparse
import json
import math
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
N_PIECES = 8                                 # A is generated as 8 column pieces of 4096 neurons
METRIC = "neurons dissected/sec (soft-WPMI, 763 concepts)"
UNIT = "neurons/s"
REF_DIR = os.path.join(ROOT, "baseline", "_ref", "concept_vit")


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
                                          "--format=csv,noheader,nounits", "-lms", "50"],
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
# CPU arm: the reference loop on a bounded sample of the workload
# ------------------------------------------------------------------------------------------------
def cpu_sample_inputs(n_neurons, seed=2):
    g = torch.Generator().manual_seed(0)
    P = torch.randn(N_IMG, C_CONCEPTS, generator=g) * 0.044        # cosines of unit-norm random rows: sigma ~ 1/sqrt(512)
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(N_IMG, n_neurons, generator=g)
    return P, A


def staged_reference():
    """The unmodified reference similarity.py, if tools/stage_reference.py staged it (it travels to the GPU box with the
    snapshot; /root/reference itself does not exist there)."""
    path = os.path.join(REF_DIR, "similarity.py")
    if not os.path.exists(path):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("_reference_similarity", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def time_cpu_loop(P, A, threads, ref=None):
    """One pass of the reference's per-neuron loop (similarity.py:49-73) on CPU tensors."""
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    if ref is not None:
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            ref.soft_wpmi(P, A, top_k=TOP_K, device="cpu")
    else:
        from oracle import similarity_oracle as orc
        inds = torch.topk(A, dim=0, k=TOP_K)[1]          # the reference's own call (similarity.py:55); tie-free data
        orc.soft_wpmi(P, A, top_k=TOP_K, inds=inds)
    return time.perf_counter() - t0


def cpu_baseline(sample_one=1024, sample_all=256):
    """~10-20 s of CPU work: the loop at 1 thread and at all host threads (the small torch ops of the per-neuron loop do
    not scale with threads -- SURVEY.md section 6 -- so the better of the two is reported)."""
    cores = os.cpu_count() or 1
    ref = staged_reference()
    P, A = cpu_sample_inputs(sample_one)
    t_one = time_cpu_loop(P, A, 1, ref)
    t_all = time_cpu_loop(P, A[:, :sample_all].contiguous(), cores, ref)
    v_one, v_all = sample_one / t_one, sample_all / t_all
    torch.set_num_threads(cores)
    best_all = v_all >= v_one
    what = ("unmodified reference concept_vit/similarity.py soft_wpmi(device='cpu') (baseline/_ref)" if ref is not None
            else "oracle loop port of similarity.py:49-73 (torch CPU ops)")
    return {"value": round(max(v_all, v_one), 1), "unit": UNIT, "cores": cores if best_all else 1,
            "kind": "reference" if ref is not None else "port",
            "sample": "%s on P[100000,763], top_k=100: %d of the 32768 neuron columns with 1 thread -> %.1f neurons/s; "
                      "%d columns with %d threads -> %.1f neurons/s" % (what, sample_one, v_one, sample_all, cores, v_all)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    all_cores = os.cpu_count() or 1
    ref = staged_reference()
    sample = 512
    P, A = cpu_sample_inputs(sample)
    # warm-up doubles as the thread-count probe: the per-neuron loop of tiny torch ops is usually FASTER on one thread
    # than on all of them; time the steps with whichever wins here
    small = A[:, :128].contiguous()
    probe = {}
    for w in range(max(args.warmup, 2)):
        th = all_cores if w % 2 == 0 else 1
        probe[th] = min(probe.get(th, 1e30), time_cpu_loop(P, small, th, ref))
    cores = min(probe, key=probe.get)
    times = [time_cpu_loop(P, A, cores, ref) for _ in range(args.steps)]
    ms = 1e3 * sum(times) / len(times)
    value = sample / (ms / 1e3)
    kind = "reference" if ref is not None else "port"
    what = ("unmodified reference similarity.soft_wpmi(device='cpu') from baseline/_ref" if ref is not None
            else "oracle loop port (torch CPU ops)")
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args.gpus), score_exchange="n/a (CPU reference arm)"),
            "cpu_baseline": {"value": round(value, 1), "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": "each step: %s, %d threads, on P[100000,763] and %d of the 32768 neuron columns, "
                                       "top_k=100" % (what, cores, sample)},
            "e2e": {"value": round(value, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {"workload": "c4: soft_wpmi, clip_feats 100000x763, target_feats 100000x32768 (all GPUs together), top_k=100, "
                        "a=10, lam=1",
            "N_img": N_IMG, "K_total": K_NEURONS, "K_per_gpu": K_NEURONS // n_gpus, "C": C_CONCEPTS, "top_k": TOP_K,
            "parallelism": "neuron-sharded x%d (strong scaling: the 32768 neurons are split over the ranks)" % n_gpus,
            "l2_policy": "inputs (13.4 GB / N per GPU) exceed the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def gen_columns(pieces, dev):
    """Columns of the c4 activation matrix: the concatenation of 4096-column pieces, piece s seeded 2 + 1000 s."""
    parts = []
    for s in pieces:
        g = torch.Generator(device=dev).manual_seed(2 + 1000 * s)
        parts.append(torch.randn(N_IMG, K_NEURONS // N_PIECES, generator=g, device=dev))
    return parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)


def sampled_parity(similarity, P, A, dev, n_cols=96, n_tie_cols=32):
    """Outside the timed region: the benched problem against the CPU oracle on sampled neuron columns (those with an fp32
    tie inside their top k first: the stated tie rule is live at this size), indices bit-exact, log-sums L within 1e-5
    relative, and log p(d) over all neurons against an fp64 logsumexp."""
    from oracle import similarity_oracle as orc
    K = A.shape[1]
    vals, idx = similarity.topk_cols(A, TOP_K, device=dev, want_values=True)
    tie_cols = ((vals[1:] == vals[:-1]).any(dim=0)).nonzero().flatten()
    g = torch.Generator().manual_seed(11)
    rnd = torch.randperm(K, generator=g)[:n_cols].to(dev)
    cols = torch.unique(torch.cat([tie_cols[:n_tie_cols], rnd]))[: n_cols + n_tie_cols]
    sub = A[:, cols].cpu()
    ref_i = orc.topk_cols(sub, TOP_K)[1]
    exact = bool(torch.equal(idx[:, cols].cpu(), ref_i))
    ramp = similarity._reference_ramp(TOP_K, 0.998, 0.97).to(dev)
    L, part = similarity.pmi_logsums(P, A, TOP_K, 10, dev, 1e-7, ramp)
    ref_L = orc.soft_wpmi_fast(P.cpu(), sub, top_k=TOP_K, lam=0, inds=ref_i)
    got = L[cols].cpu()
    rel = ((got - ref_L).abs() / ref_L.abs().clamp_min(1e-30)).max().item()
    agree = (got.argmax(1) == ref_L.argmax(1)).float().mean().item()
    out, prob_d = similarity.pmi_finalize(L.clone(), part, K, 1.0)
    truth = torch.logsumexp(L.double(), dim=0) - math.log(K)
    d_err = (prob_d.double() - truth).abs().max().item()
    del vals, idx, L, out
    return {"checked_against": "oracle/similarity_oracle.py on %d sampled neuron columns of the benched matrix" % len(cols),
            "cols_checked": int(len(cols)), "cols_with_fp32_tie_in_topk": int(len(tie_cols)),
            "tie_cols_checked": int(min(len(tie_cols), n_tie_cols)), "topk_indices_bit_exact": exact,
            "L_max_rel_err": rel, "L_tol": 1e-5, "top_concept_agreement": agree,
            "log_pd_max_abs_err_vs_fp64": d_err, "ok": bool(exact and rel <= 1e-5 and d_err <= 1e-3)}


def sharded_parity(similarity, mdist, dev, world, rank):
    """N > 1, before timing: the neuron-sharded call on a small problem must reproduce the single-GPU bits (every rank
    also runs the unsharded call) for every score-exchange mode."""
    import torch.distributed as dist
    N, K, C = 40000, 256 * 3 * world + 100, C_CONCEPTS        # long columns: the bench's kernels (filter scan)
    g = torch.Generator(device=dev).manual_seed(5)
    P = torch.randn(N, C, generator=g, device=dev) * 0.05
    A = torch.randn(N, K, generator=g, device=dev)
    want = similarity.soft_wpmi(P, A, top_k=TOP_K, device=dev)
    b = mdist.shard_bounds(K, world)
    sizes = [b[i + 1] - b[i] for i in range(world)]
    shard = A[:, b[rank]:b[rank + 1]].contiguous()
    res = {}
    backend = mdist.CudaBackend(dev)
    modes = ["nccl", "copy", "fused", "peer_partials+nccl", "peer_partials+copy"]
    for mode in modes:
        try:
            score_mode = mode.split("+")[-1]
            ex = None if score_mode == "nccl" else mdist.PeerScoreExchange(sizes, C, dev, mode=score_mode)
            pex = mdist.PeerPartialsExchange(sizes, C, dev) if mode.startswith("peer_partials") else None
            ok = 1
            for _ in range(3 if pex is not None else 1):      # the partials tables alternate: go round more than once
                got = mdist.soft_wpmi_sharded(P, shard, sizes, top_k=TOP_K, device=dev, backend=backend, exchange=ex,
                                              partials_exchange=pex)
                ok = min(ok, int(torch.equal(got, want)))
        except Exception as exc:                              # no symmetric memory on this box: reported, not hidden
            sys.stderr.write("rank %d: exchange mode %s unavailable: %s\n" % (rank, mode, str(exc)[:200]))
            ok = -1
        t = torch.tensor([ok], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        res[mode] = {1: "bit-identical to the single-GPU call on every rank", 0: "MISMATCH", -1: "unavailable"}[int(t.item())]
        torch.cuda.synchronize(dev)
    pex = None
    if res.get("peer_partials+nccl", "").startswith("bit-identical"):
        pex = mdist.PeerPartialsExchange(sizes, C, dev)
    local = mdist.soft_wpmi_sharded(P, shard, sizes, top_k=TOP_K, device=dev, backend=backend, gather_scores=False,
                                    partials_exchange=pex)
    t = torch.tensor([int(torch.equal(local, want[b[rank]:b[rank + 1]]))], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    res["own_shard_only"] = "bit-identical" if int(t.item()) == 1 else "MISMATCH"
    res["problem"] = "N=%d, K=%d over %d ranks (shards %s), C=%d, top_k=%d" % (N, K, world, sizes, C, TOP_K)
    res["ok"] = all(v != "MISMATCH" for v in res.values())
    return res


def bind_to_gpu_cpus(local_rank):
    """Several ranks stream gigabytes from pinned host memory at once (the e2e leg): run this rank on the CPUs next to its
    GPU, so that the pages it pins are first touched -- and therefore placed -- on that NUMA node.  Returns a short
    description, or None when NVML does not say (nothing is changed then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return "%d CPUs next to GPU %d" % (len(cpus), local_rank)
    except Exception:
        return None


def run_ours(args):
    from mammo_clip_dissect_b200 import _lib, similarity
    from mammo_clip_dissect_b200 import distributed as mdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    if N_PIECES % world != 0:
        raise SystemExit("--gpus must divide %d" % N_PIECES)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_cpus(local_rank) if world > 1 else None      # pinned host buffers on the GPU's own NUMA node
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.lib().mcd_device_check(), "mcd_device_check")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def per_rank(x):
        if world == 1:
            return [round(x, 4)]
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [round(float(o.item()), 4) for o in out]

    parity = {}
    if world > 1:
        parity["sharded_vs_single_gpu"] = sharded_parity(similarity, mdist, dev, world, rank)
        torch.cuda.empty_cache()

    # ---- synthetic inputs, resident in HBM --------------------------------------------------
    g = torch.Generator(device=dev).manual_seed(0)
    P = torch.randn(N_IMG, C_CONCEPTS, generator=g, device=dev) * 0.044
    per = N_PIECES // world
    A = gen_columns(range(rank * per, (rank + 1) * per), dev)              # this rank's shard of the 32768 neurons
    K_local = A.shape[1]
    sizes = [K_local] * world
    backend = mdist.CudaBackend(dev) if world > 1 else None
    xmode = os.environ.get("MCD_EXCHANGE", "copy")
    exchange = None
    if world > 1 and xmode != "nccl":
        try:
            exchange = mdist.PeerScoreExchange(sizes, C_CONCEPTS, dev, mode=xmode)
            ok = 1
        except Exception as exc:                                   # no peer access / no symmetric memory on this box
            sys.stderr.write("rank %d: peer score exchange unavailable (%s); using NCCL all_gather\n" % (rank, exc))
            ok = 0
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            exchange, xmode = None, "nccl"

    # the LSE partials go through symmetric memory too (one small kernel + the signal-pad barrier instead of an NCCL
    # all_gather); MCD_PARTIALS=nccl keeps the collective
    pexchange, pmode = None, os.environ.get("MCD_PARTIALS", "peer")
    if world > 1 and pmode != "nccl":
        try:
            pexchange = mdist.PeerPartialsExchange(sizes, C_CONCEPTS, dev)
            ok = 1
        except Exception as exc:
            sys.stderr.write("rank %d: peer partials exchange unavailable (%s); using NCCL all_gather\n" % (rank, exc))
            ok = 0
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            pexchange, pmode = None, "nccl"

    def step(P_in, A_in, gather=False):
        """One full soft_wpmi over the 32768 neurons: every rank ends with its finalized [K/N, C] shard (gather=False) or
        with the whole [K, C] matrix."""
        if world == 1:
            return similarity.soft_wpmi(P_in, A_in, top_k=TOP_K, device=dev)
        return mdist.soft_wpmi_sharded(P_in, A_in, sizes, top_k=TOP_K, device=dev, backend=backend, gather_scores=gather,
                                       exchange=exchange if gather else None, partials_exchange=pexchange)

    def timed(fn, steps, warm):
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1) / steps)

    warm = max(args.warmup, 3)
    # ---- device-resident timing: K steps between two events, the public call ---------------------
    for _ in range(warm):
        out = step(P, A)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    n0 = _lib.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step(P, A)
    e1.record()
    barrier()
    launches = _lib.launch_count() - n0
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    if rank == 0:
        time.sleep(0.1)
    clocks = sampler.stop() if rank == 0 else None
    value = K_NEURONS / (ms_step / 1e3)
    del out

    # ---- N > 1: the same step plus the exchange of the score shards (every rank gets [32768, 763]) -----------------
    with_exchange = None
    if world > 1:
        ms_x = timed(lambda: step(P, A, gather=True), max(3, args.steps // 2), 2)
        with_exchange = {"ms_per_step": round(ms_x, 4), "value": round(K_NEURONS / (ms_x / 1e3), 1), "unit": UNIT,
                         "mode": {"copy": "finalize in place + one DMA push per peer into peer-mapped [K,C] buffers",
                                  "fused": "finalize kernel stores its slice into every peer's [K,C] buffer (NVLink stores)",
                                  "nccl": "NCCL all_gather"}[xmode]}

    # ---- staged pass: per-stage CUDA events (the kernels one after the other on one stream) --------------------------
    def staged_call():
        if world == 1:
            similarity.pmi_scores(P, A, TOP_K, 10, 1, dev, 1e-7,
                                  similarity._device_ramp(similarity._reference_ramp(TOP_K, 0.998, 0.97), TOP_K, 0.998, 0.97, dev))
        else:
            step(P, A)

    # (the staged kernels keep their own cached workspaces: two untimed calls first, so that no allocation of the 0.6 GB
    # K2 workspace lands between a stage's events)
    similarity.PROFILE = []
    for _ in range(2):
        staged_call()
    torch.cuda.synchronize(dev)
    similarity.PROFILE = []
    for _ in range(5):
        staged_call()
    stage_ms = similarity.profile_summary()
    similarity.PROFILE = None
    stage_ms = {k: max_over_ranks(v) for k, v in sorted(stage_ms.items())}

    # ---- parity on the benched problem (rank 0's shard), outside every timed region ---------------------------------
    if rank == 0 and not args.no_parity:
        parity["benched_problem_vs_oracle"] = sampled_parity(similarity, P, A, dev)
    barrier()

    # ---- weak scaling (N > 1): 32768 neurons PER rank -----------------------------------------------------------------
    weak = None
    if world > 1 and not args.no_weak:
        del A
        torch.cuda.empty_cache()
        A = torch.randn(N_IMG, K_NEURONS, generator=torch.Generator(device=dev).manual_seed(2 + 1000 * rank), device=dev)
        wsizes = [K_NEURONS] * world
        wpex = None
        if pexchange is not None:
            wpex = mdist.PeerPartialsExchange(wsizes, C_CONCEPTS, dev)
        wfn = lambda: mdist.soft_wpmi_sharded(P, A, wsizes, top_k=TOP_K, device=dev, backend=backend, gather_scores=False,  # noqa: E731
                                              partials_exchange=wpex)
        wsampler = ClockSampler(local_rank)
        if rank == 0:
            wsampler.start()
            time.sleep(0.25)
        ms_w = timed(wfn, max(10, 2 * args.steps), 3)
        if rank == 0:
            time.sleep(0.1)
        wclocks = wsampler.stop() if rank == 0 else None
        similarity.PROFILE = []
        wfn()                                        # untimed: first use of the staged kernels' workspaces at this width
        torch.cuda.synchronize(dev)
        similarity.PROFILE = []
        for _ in range(3):
            wfn()
        wstage = {k: round(max_over_ranks(v), 4) for k, v in sorted(similarity.profile_summary().items())}
        similarity.PROFILE = None
        # the same call without the exchange of the LSE partials: what a rank does on its own
        solo_fn = lambda: similarity.pmi_logsums(P, A, TOP_K, 10, dev, 1e-7, similarity._device_ramp(  # noqa: E731
            similarity._reference_ramp(TOP_K, 0.998, 0.97), TOP_K, 0.998, 0.97, dev))
        solo = timed(solo_fn, max(5, args.steps), 2)
        # ... and every rank's own time for it (no collective inside): tells a slow GPU from a scaling effect
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            solo_fn()
        e1.record()
        torch.cuda.synchronize(dev)
        solo_ranks = per_rank(e0.elapsed_time(e1) / 5)
        weak = {"ms_per_step": round(ms_w, 4), "value": round(K_NEURONS * world / (ms_w / 1e3), 1), "unit": UNIT,
                "K_per_gpu": K_NEURONS, "K_total": K_NEURONS * world, "stage_ms": wstage, "clocks": wclocks,
                "ms_per_step_without_the_partials_exchange": round(solo, 4), "the_same_per_rank_ms": solo_ranks,
                "note": "every rank scores 32768 neurons (layer width grows with N); LSE partials all-gathered, shards stay local"}
        del A
        torch.cuda.empty_cache()
        A = gen_columns(range(rank * per, (rank + 1) * per), dev)

    # ---- end to end: pinned host inputs -> public API -> host result ------------------------------------------------
    e2e = None
    try:
        P_h = torch.empty(P.shape, dtype=P.dtype, pin_memory=True).copy_(P)
        A_h = torch.empty(A.shape, dtype=A.dtype, pin_memory=True).copy_(A)
        del A
        torch.cuda.empty_cache()
        res_h = torch.empty((K_local, C_CONCEPTS), dtype=torch.float32, pin_memory=True)
        # the link's own ceiling, next to the number: one plain pinned H2D copy of this rank's A
        stage_buf = torch.empty(A_h.shape, dtype=A_h.dtype, device=dev)
        pcie_ms = timed(lambda: stage_buf.copy_(A_h, non_blocking=True), 2, 1)
        del stage_buf
        torch.cuda.empty_cache()
        e2e_steps = max(2, min(args.steps, 5))
        ms_e2e = timed(lambda: res_h.copy_(step(P_h, A_h), non_blocking=True), e2e_steps, 2)
        e2e = {"value": round(K_NEURONS / (ms_e2e / 1e3), 1), "unit": UNIT,
               "h2d_bytes_per_step": int(P_h.numel() * 4 + A_h.numel() * 4), "d2h_bytes_per_step": int(res_h.numel() * 4),
               "ms_per_step": round(ms_e2e, 3), "steps": e2e_steps,
               "pcie_h2d_gbs_plain_copy": round(A_h.numel() * 4 / pcie_ms / 1e6, 1), "cpu_binding": numa,
               "note": "PCIe-bound: per rank %.2f GB in, %.3f GB out per step; the scoring itself is %.1f ms of it"
                       % ((P_h.numel() + A_h.numel()) * 4 / 1e9, res_h.numel() * 4 / 1e9, ms_step),
               "api": ("similarity.soft_wpmi" if world == 1 else "distributed.soft_wpmi_sharded(gather_scores=False)") +
                      "(P_host_pinned, A_host_pinned, device='cuda') -> this rank's score rows copied to pinned host memory"}
    except RuntimeError as exc:                      # e.g. the box cannot pin 13.4 GB
        e2e = {"value": None, "unit": UNIT, "error": str(exc)[:200]}

    # ---- the reference's own GPU path on a neuron slice (rank 0, N = 1) ----------------------------------------------
    ref_gpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref = staged_reference()
        if ref is not None:
            try:
                import contextlib
                import io
                n_slice = 256
                A_s = gen_columns([0], dev)[:, :n_slice].contiguous()
                with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                    ref.soft_wpmi(P, A_s[:, :32], top_k=TOP_K, device=dev)
                    torch.cuda.synchronize(dev)
                    t0 = time.perf_counter()
                    ref.soft_wpmi(P, A_s, top_k=TOP_K, device=dev)
                    torch.cuda.synchronize(dev)
                    dt = time.perf_counter() - t0
                ref_gpu = {"value": round(n_slice / dt, 1), "unit": UNIT,
                           "sample": "unmodified reference similarity.soft_wpmi(device='cuda') (Python loop over neurons, "
                                     "empty_cache per neuron) on P[100000,763] and %d neuron columns, inputs resident on the "
                                     "GPU; per-neuron rate, extrapolates linearly to 32768" % n_slice}
                del A_s
            except Exception as exc:
                ref_gpu = {"value": None, "error": str(exc)[:200]}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant stage + whole path ----------------------------------------
    peak, peak_src = peaks()
    topk_ms = stage_ms.get("topk_cols")
    alg_topk = 4.0 * N_IMG * K_local                                    # read this rank's A once
    alg_path = (4.0 * N_IMG * K_local + 4.0 * N_IMG * C_CONCEPTS + 4.0 * K_local * C_CONCEPTS) * world
    tr = traffic_note()
    roof = None
    if topk_ms:
        ach = alg_topk / (topk_ms / 1e3) / 1e9
        roof = {"bound": "hbm",
                "kernel": "filter_scan_kernel; timed as the whole topk_cols stage of the staged pass (sample_tilemax + "
                          "sample_select + filter scan + select + redo) with CUDA events, so the fraction is a lower bound "
                          "for the scan kernel itself",
                "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                "traffic": (tr.get("filter_scan_kernel") * K_local / 32768.0 if tr.get("filter_scan_kernel") else None), "peak_source": peak_src, "ms_per_launch": round(topk_ms, 4),
                "algorithmic_bytes_per_launch": alg_topk}
    ach_path = alg_path / (ms_step / 1e3) / 1e9
    cfg = workload_config(world)
    cfg["score_exchange"] = "none (one GPU)" if world == 1 else "shards stay on their rank in `value`; see with_score_exchange"
    if world > 1:
        cfg["partials_exchange"] = ("symmetric-memory stores + signal-pad barrier" if pexchange is not None
                                    else "NCCL all_gather")
    line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
            "roofline_path": {"bound": "hbm", "achieved": round(ach_path, 1), "peak": peak * world, "unit": "GB/s",
                              "frac": round(ach_path / (peak * world), 4), "algorithmic_bytes_per_step": alg_path,
                              "note": "B_alg = 4NK + 4NC + 4KC (SURVEY.md 8d; P and its softmax are replicated, so 4NC "
                                      "counts once per GPU); gather re-reads not counted; peak = N x one GPU's measured HBM peak"},
            "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()},
            "stage_note": "stage_ms is a separate staged pass with CUDA events between the stages (max over ranks); "
                          "ms_per_step is the public call (one C entry point; the softmax of a narrow shard runs on a "
                          "side stream beside the sample pass and the scan)",
            "with_score_exchange": with_exchange, "weak": weak, "parity": parity, "reference_gpu": ref_gpu}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-weak", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
