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
  exchange 2: all_gather of the [K_g, C] score shards into the full [K, C] matrix (optional)

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

    def log_sums(self, clip_feats, target_shard, top_k, a, min_prob, ramp):
        sim = self.sim
        A = sim._as_f32_matrix(target_shard, self.device, "target_feats")
        with sim._Stage("softmax_rows"):
            S = sim.concept_probabilities(clip_feats, a, self.device)
        with sim._Stage("topk_cols"):
            idx32 = sim._topk_int32(A, int(top_k), self.device)
        weights = ramp.to(self.device) if ramp is not None else None
        with sim._Stage("wpmi_accum"):
            return sim.log_sums(S, idx32, weights, min_prob)

    def lse_partials(self, L):
        with self.sim._Stage("lse_partials"):
            return self.sim.lse_partials(L)

    def finalize(self, L, partials_all, K_total, lam):
        with self.sim._Stage("lse_finalize"):
            return self.sim.pmi_finalize(L, partials_all, K_total, lam)[0]


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
                       backend, group=None, gather_scores: bool = True):
    """Scores for this rank's neurons (and, with gather_scores, for all neurons [K, C]).

    target_shard : [N, K_g] activations of this rank's neurons
    shard_sizes  : K_g of every rank, in rank order (interior boundaries multiples of 256)
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert len(shard_sizes) == world and shard_sizes[rank] == target_shard.shape[1]
    for g in range(world - 1):
        if shard_sizes[g] % LSE_BLOCK != 0:
            raise RuntimeError("interior shard boundaries must be multiples of %d neurons" % LSE_BLOCK)
    K_total = int(sum(shard_sizes))
    L = backend.log_sums(clip_feats, target_shard, top_k, a, min_prob, ramp)
    part = backend.lse_partials(L)
    nblocks = [(s + LSE_BLOCK - 1) // LSE_BLOCK for s in shard_sizes]
    part_all = _all_gather_var(part, nblocks, group) if world > 1 else part
    local = backend.finalize(L, part_all, K_total, lam)
    if not gather_scores or world == 1:
        return local
    stage = getattr(getattr(backend, "sim", None), "_Stage", None)
    if stage is None:
        return _all_gather_var(local, list(shard_sizes), group)
    with stage("allgather_scores"):
        return _all_gather_var(local, list(shard_sizes), group)


def soft_wpmi_sharded(clip_feats, target_shard, shard_sizes, top_k=100, a=10, lam=1, device='cuda',
                      min_prob=1e-7, p_start=0.998, p_end=0.97, group=None, gather_scores=True, backend=None):
    """Neuron-sharded soft_wpmi (reference similarity.py:49-73 semantics over the union of shards)."""
    from .similarity import _reference_ramp
    backend = backend or CudaBackend(device)
    return pmi_scores_sharded(clip_feats, target_shard, shard_sizes, top_k, a, lam, min_prob,
                              _reference_ramp(int(top_k), p_start, p_end), backend, group, gather_scores)


def wpmi_sharded(clip_feats, target_shard, shard_sizes, top_k=28, a=2, lam=0.6, device='cuda', min_prob=1e-7,
                 group=None, gather_scores=True, backend=None):
    backend = backend or CudaBackend(device)
    return pmi_scores_sharded(clip_feats, target_shard, shard_sizes, top_k, a, lam, min_prob, None, backend, group,
                              gather_scores)
