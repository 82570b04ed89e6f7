"""Neuron-sharded soft-WPMI / WPMI across the GPUs of one node (one process per GPU,
torch.distributed over NCCL / NVLink).

The reference path is single-GPU; the shard structure follows from its maths:
top-k, gather and log-sum are independent per neuron (column of target_feats), and the only
coupling is log p(d)[c] = logsumexp_j L[j,c] - log K over ALL neurons of the call
(reference similarity.py:70-71).  So

  rank g holds target_feats[:, shard_g] (contiguous neuron ranges, boundaries on multiples of the
  256-neuron LSE block), every rank holds clip_feats (replicated; each rank runs the softmax itself
  rather than receiving a 305 MB broadcast),
  exchange 1: all_gather of the per-256-neuron-block (max, sum-exp) partials  [nb_g, 2, C]   (tiny)
  local     : combine ALL partials in global block order (fp64)  -> bit-identical for any G
  exchange 2: all_gather of the [K_g, C] score shards into the full [K, C] matrix (optional).  Three ways:
              * NCCL all_gather (default; a fresh tensor per call);
              * PeerScoreExchange(mode="fused"): K3b's finalize kernel stores its slice straight into every GPU's
                [K, C] buffer through peer-mapped (symmetric) memory -- one kernel, no separate collective;
              * PeerScoreExchange(mode="copy"): the slice is pushed to the peers by the copy engines on side
                streams, so the exchange of one call can overlap the column scan of the next call (the scan is a
                single wave of persistent warps that fills every SM: any SM-based collective running beside it
                would delay it by its own duration; DMA pushes do not).

Compute is injected through a small backend object so that the exchange logic can be tested on
CPU with gloo (tests/test_dist_gloo.py drives it with the oracle); the product backend below
calls the CUDA kernels and nothing else.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

LSE_BLOCK = 256


def shard_bounds(K: int, world: int, block: int = LSE_BLOCK) -> List[int]:
    """Boundaries b[0..world] of contiguous neuron shards; every interior boundary is a multiple of
    `block`, so the global list of LSE blocks is the concatenation of the ranks' local blocks."""
    nb = (K + block - 1) // block
    base, extra = divmod(nb, world)
    bounds = [0]
    for g in range(world):
        blocks = base + (1 if g < extra else 0)
        bounds.append(min(K, bounds[-1] + blocks * block))
    bounds[-1] = K
    return bounds


class CudaBackend:
    """The product backend: sm_100a kernels through the C ABI (similarity.py building blocks)."""

    def __init__(self, device):
        from . import similarity
        self.sim = similarity
        self.device = similarity._cuda_device(device)
        self._ramps = {}            # top_k -> (host ramp, device copy): no pageable H2D copy per call

    def _ramp(self, ramp, top_k):
        if ramp is None:
            return None
        if ramp.is_cuda:
            return ramp if ramp.device == self.device else ramp.to(self.device)
        hit = self._ramps.get(int(top_k))
        if hit is None or not torch.equal(hit[0], ramp):
            hit = self._ramps[int(top_k)] = (ramp.clone(), ramp.to(self.device))
        return hit[1]

    def log_sums(self, clip_feats, target_shard, top_k, a, min_prob, ramp):
        sim = self.sim
        with torch.no_grad(), torch.cuda.device(self.device):
            A = sim._as_f32_matrix(target_shard, self.device, "target_feats")
            with sim._Stage("softmax_rows"):
                S = sim.concept_probabilities(clip_feats, a, self.device)
            with sim._Stage("topk_cols"):
                idx32 = sim._topk_int32(A, int(top_k), self.device)
            weights = self._ramp(ramp, top_k)
            with sim._Stage("wpmi_accum"):
                return sim.log_sums(S, idx32, weights, min_prob)

    def lse_partials(self, L):
        with torch.no_grad(), torch.cuda.device(self.device), self.sim._Stage("lse_partials"):
            return self.sim.lse_partials(L)

    def log_sums_and_partials(self, clip_feats, target_shard, top_k, a, min_prob, ramp):
        """Both of the above behind one entry point (mcd_pmi_logsums_f32: column chunks pipelined, K3 of chunk q under
        the scan of chunk q + 1).  Per-stage profiling needs the staged kernels, so it takes the two-call route."""
        sim = self.sim
        if sim.PROFILE is not None:
            L = self.log_sums(clip_feats, target_shard, top_k, a, min_prob, ramp)
            return L, self.lse_partials(L)
        with torch.no_grad(), torch.cuda.device(self.device):
            return sim.pmi_logsums(clip_feats, target_shard, top_k, a, self.device, min_prob, self._ramp(ramp, top_k))

    def finalize(self, L, partials_all, K_total, lam):
        with torch.no_grad(), torch.cuda.device(self.device), self.sim._Stage("lse_finalize"):
            return self.sim.pmi_finalize(L, partials_all, K_total, lam)[0]


class GatheredScores:
    """Result of an asynchronous score exchange: `.wait()` makes the current stream wait for the exchange and
    returns the [K_total, C] matrix (a view of a double-buffered symmetric buffer: it is overwritten by the
    exchange `depth` calls later)."""

    def __init__(self, tensor, event, keep=None):
        self.tensor, self.event, self._keep = tensor, event, keep

    def wait(self):
        if self.event is not None:
            torch.cuda.current_stream(self.tensor.device).wait_event(self.event)
            self.event = None
            self._keep = None
        return self.tensor


class PeerScoreExchange:
    """[K_total, C] score buffers in symmetric memory (torch.distributed._symmetric_memory: every rank's buffer is
    mapped into every process of the node), `depth` of them used round-robin.

    mode "fused": mcd_pmi_finalize_bcast_f32 writes the finalized slice into all ranks' buffers (NVLink stores);
    mode "copy" : mcd_pmi_finalize_f32 in place, then one DMA push per peer on side streams.
    Both are ordered by the symmetric-memory barrier (signal pads, a few microseconds): one before the first remote
    write (every rank has finished the calls that could still be reading this buffer) and one after the last."""

    def __init__(self, shard_sizes: Sequence[int], C: int, device, group=None, mode: str = "copy", depth: int = 2):
        import torch.distributed._symmetric_memory as symm_mem
        if mode not in ("copy", "fused"):
            raise ValueError("mode must be 'copy' or 'fused'")
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.mode, self.depth, self.C = mode, int(depth), int(C)
        self.sizes = [int(x) for x in shard_sizes]
        self.K_total = sum(self.sizes)
        self.row0 = sum(self.sizes[: self.rank])
        self.device = torch.device(device)
        self.bufs, self.hdls, self.views = [], [], []
        for _ in range(self.depth):
            t = symm_mem.empty((self.K_total, self.C), dtype=torch.float32, device=self.device)
            h = symm_mem.rendezvous(t, self.group)
            self.bufs.append(t)
            self.hdls.append(h)
            self.views.append([h.get_buffer(p, (self.K_total, self.C), torch.float32) for p in range(self.world)])
        self.comm = torch.cuda.Stream(device=self.device)
        self.push_streams = [torch.cuda.Stream(device=self.device) for _ in range(self.world)]
        self.turn = 0

    def exchange(self, L, partials_all, lam, sim, wait: bool = True):
        """L: this rank's [K_g, C] log-sums (finalized in place in mode "copy").  Returns the gathered matrix
        (wait=True) or a GatheredScores handle (wait=False; only mode "copy" really runs behind the caller)."""
        with torch.no_grad(), torch.cuda.device(self.device):
            return self._exchange(L, partials_all, lam, sim, wait)

    def _exchange(self, L, partials_all, lam, sim, wait):
        b = self.turn
        self.turn = (self.turn + 1) % self.depth
        hdl, buf, views = self.hdls[b], self.bufs[b], self.views[b]
        main = torch.cuda.current_stream(self.device)
        rows = slice(self.row0, self.row0 + self.sizes[self.rank])
        if self.mode == "fused":
            hdl.barrier(channel=0)
            if L.shape[0] > 0:
                with sim._Stage("finalize_bcast"):
                    sim.pmi_finalize_bcast(L, partials_all, self.K_total, lam, list(hdl.buffer_ptrs), self.row0)
            hdl.barrier(channel=1)
            return buf if wait else GatheredScores(buf, None)
        local = L
        if L.shape[0] > 0:
            with sim._Stage("lse_finalize"):
                local = sim.pmi_finalize(L, partials_all, self.K_total, lam)[0]
        ready = torch.cuda.Event()
        ready.record(main)
        self.comm.wait_event(ready)
        with torch.cuda.stream(self.comm):
            hdl.barrier(channel=0)
            start = torch.cuda.Event()
            start.record(self.comm)
            for i in range(self.world):
                p = (self.rank + i) % self.world            # own copy first, then the peers in ring order
                st = self.push_streams[i]
                st.wait_event(start)
                with torch.cuda.stream(st):
                    if local.shape[0] > 0:
                        views[p][rows].copy_(local, non_blocking=True)
                self.comm.wait_stream(st)
            hdl.barrier(channel=1)
            done = torch.cuda.Event()
            done.record(self.comm)
        for st in self.push_streams + [self.comm]:
            local.record_stream(st)
        out = GatheredScores(buf, done, keep=local)
        return out.wait() if wait else out


class PeerPartialsExchange:
    """The all-gather of the LSE partials through symmetric memory: every rank stores its [blocks_g, 2, C] slice into all
    ranks' [blocks_total, 2, C] tables with one small kernel (NVLink stores) and the ranks meet at the signal-pad
    barrier.  Two tables used in turn: a rank may still be combining table b of call i while a faster rank already writes
    table b ^ 1 of call i + 1; nobody can write table b again before every rank has passed the barrier of call i + 1, i.e.
    has finished call i.  Same values as the NCCL all_gather (a copy), so the result stays bit-identical."""

    def __init__(self, shard_sizes: Sequence[int], C: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.sizes = [int(x) for x in shard_sizes]
        self.nblocks = [(s + LSE_BLOCK - 1) // LSE_BLOCK for s in self.sizes]
        self.C = int(C)
        self.device = torch.device(device)
        self.blk0 = sum(self.nblocks[: self.rank])
        self.total = sum(self.nblocks)
        self.bufs, self.hdls = [], []
        for _ in range(2):
            t = symm_mem.empty((self.total, 2, self.C), dtype=torch.float32, device=self.device)
            self.hdls.append(symm_mem.rendezvous(t, self.group))
            self.bufs.append(t)
        self.turn = 0

    def gather(self, part, sim):
        b = self.turn
        self.turn ^= 1
        hdl, buf = self.hdls[b], self.bufs[b]
        with torch.no_grad(), torch.cuda.device(self.device):
            if part.shape[0] != self.nblocks[self.rank] or (part.shape[0] and part.shape[2] != self.C):
                raise RuntimeError("PeerPartialsExchange was built for other shard sizes")
            if part.shape[0] > 0:
                sim.bcast_rows(part.contiguous(), list(hdl.buffer_ptrs), self.blk0 * 2 * self.C)
            hdl.barrier(channel=0)
        return buf


def _all_gather_var(t: torch.Tensor, sizes: Sequence[int], group) -> torch.Tensor:
    """all_gather of tensors whose leading dimension differs per rank (sizes known to all ranks)."""
    world = dist.get_world_size(group)
    if len(set(sizes)) == 1:
        out = torch.empty((world * sizes[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
        return out
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    buf[: t.shape[0]] = t
    out = torch.empty((world * pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return torch.cat([out[g * pad: g * pad + sizes[g]] for g in range(world)], dim=0)


def pmi_scores_sharded(clip_feats, target_shard, shard_sizes: Sequence[int], top_k, a, lam, min_prob, ramp,
                       backend, group=None, gather_scores: bool = True, exchange=None, wait: bool = True,
                       partials_exchange=None):
    """Scores for this rank's neurons (and, with gather_scores, for all neurons [K, C]).

    target_shard : [N, K_g] activations of this rank's neurons
    shard_sizes  : K_g of every rank, in rank order (interior boundaries multiples of 256)
    exchange     : a PeerScoreExchange for the score all-gather (default: NCCL all_gather); with wait=False the
                   result is a GatheredScores handle
    partials_exchange : a PeerPartialsExchange for the all-gather of the LSE partials (default: NCCL all_gather)
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert len(shard_sizes) == world and shard_sizes[rank] == target_shard.shape[1]
    # a boundary needs 256-alignment only if neurons follow it (small layers: K < 256 * (world - 1) leaves the last
    # ranks with empty shards, which skip the compute and contribute no partial blocks)
    for g in range(world - 1):
        if shard_sizes[g] % LSE_BLOCK != 0 and any(int(x) > 0 for x in shard_sizes[g + 1:]):
            raise RuntimeError("interior shard boundaries must be multiples of %d neurons" % LSE_BLOCK)
    K_total = int(sum(shard_sizes))
    C = clip_feats.shape[1]
    empty = int(shard_sizes[rank]) == 0
    if empty:
        dev = getattr(backend, "device", target_shard.device)
        L = torch.empty((0, C), dtype=torch.float32, device=dev)
        part = torch.empty((0, 2, C), dtype=torch.float32, device=dev)
    elif hasattr(backend, "log_sums_and_partials"):
        L, part = backend.log_sums_and_partials(clip_feats, target_shard, top_k, a, min_prob, ramp)
    else:
        L = backend.log_sums(clip_feats, target_shard, top_k, a, min_prob, ramp)
        part = backend.lse_partials(L)
    nblocks = [(s + LSE_BLOCK - 1) // LSE_BLOCK for s in shard_sizes]
    if world > 1 and partials_exchange is not None:
        if list(partials_exchange.sizes) != [int(x) for x in shard_sizes] or partials_exchange.C != C:
            raise RuntimeError("PeerPartialsExchange was built for other shard sizes")
        part_all = partials_exchange.gather(part, backend.sim)
    else:
        part_all = _all_gather_var(part, nblocks, group) if world > 1 else part
    if exchange is not None and gather_scores and world > 1:
        if list(exchange.sizes) != [int(x) for x in shard_sizes] or exchange.C != L.shape[1]:
            raise RuntimeError("PeerScoreExchange was built for other shard sizes")
        return exchange.exchange(L, part_all, lam, backend.sim, wait=wait)
    local = L if empty else backend.finalize(L, part_all, K_total, lam)
    if not gather_scores or world == 1:
        return local
    stage = getattr(getattr(backend, "sim", None), "_Stage", None)
    if stage is None:
        return _all_gather_var(local, list(shard_sizes), group)
    with stage("allgather_scores"):
        return _all_gather_var(local, list(shard_sizes), group)


def soft_wpmi_sharded(clip_feats, target_shard, shard_sizes, top_k=100, a=10, lam=1, device='cuda',
                      min_prob=1e-7, p_start=0.998, p_end=0.97, group=None, gather_scores=True, backend=None,
                      exchange=None, wait=True, partials_exchange=None):
    """Neuron-sharded soft_wpmi (reference similarity.py:49-73 semantics over the union of shards)."""
    from .similarity import _reference_ramp
    backend = backend or CudaBackend(device)
    return pmi_scores_sharded(clip_feats, target_shard, shard_sizes, top_k, a, lam, min_prob,
                              _reference_ramp(int(top_k), p_start, p_end), backend, group, gather_scores,
                              exchange, wait, partials_exchange)


def wpmi_sharded(clip_feats, target_shard, shard_sizes, top_k=28, a=2, lam=0.6, device='cuda', min_prob=1e-7,
                 group=None, gather_scores=True, backend=None, exchange=None, wait=True, partials_exchange=None):
    backend = backend or CudaBackend(device)
    return pmi_scores_sharded(clip_feats, target_shard, shard_sizes, top_k, a, lam, min_prob, None, backend, group,
                              gather_scores, exchange, wait, partials_exchange)
