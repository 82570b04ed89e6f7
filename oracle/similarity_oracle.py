"""CPU oracle for the neuron->concept scoring path.  TEST INFRASTRUCTURE ONLY.

This module restates, on the CPU with torch tensor arithmetic, the algorithm of the
reference's scoring path.  It is the *checker* for the CUDA product in
``mammo_clip_dissect_b200`` and the ``cpu_baseline`` leg of ``bench.py``; nothing in
the product package may import it (the product fails loudly without its CUDA
library instead of falling back to this file).

Reference sites restated (paths relative to /root/reference):
  * concept_vit/similarity.py:49-73   soft_wpmi
  * concept_vit/similarity.py:75-97   wpmi
  * concept_vit/similarity.py:7-31    cos_similarity_cubed
  * concept_vit/similarity.py:33-47   cos_similarity
  * concept_vit/similarity.py:99-132  rank_reorder
  * concept_vit/CLIP_og_utils.py:155-160 (= utils.py:570-594) row-normalise + I @ T.T
  * concept_vit/utils.py:27-52        get_activation forward hook

Where the arithmetic lives: the reference calls PyTorch (pinned torch==2.2.2 in
requirements.txt:2; this image has torch 2.11).  The reference ships no tests and no
golden vectors, so the pin is "reference source x installed torch": the fixtures in
tests/golden/ were produced by importing /root/reference/concept_vit/similarity.py in
the authoring container (tests/golden/make_golden.py) and this oracle is checked
against them bit-for-bit on tie-free inputs (tests/test_oracle_golden.py).

One deliberate difference from the reference: ``torch.topk`` leaves both the order of
equal values and WHICH equal values make the cut unspecified.  The oracle (and the
CUDA path) use a stated total order instead -- value descending, then probe-image
index ascending, NaN greater than +inf, -0.0 == +0.0 -- implemented as a stable
descending sort.  On tie-free columns this equals ``torch.topk`` exactly.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional

import torch

LSE_BLOCK = 256  # neurons per log-sum-exp partial (mirrors the product's fixed block)


# --------------------------------------------------------------------------------------
# top-k along the probe-image axis under the stated tie-break rule
# --------------------------------------------------------------------------------------
def topk_cols(target_feats: torch.Tensor, k: int):
    """(values[k,K], indices[k,K]) of the k largest entries of every column.

    Order: value desc, image index asc; NaN is the largest value; -0.0 == +0.0.
    Replaces ``torch.topk(target_feats, dim=0, k=k)`` (similarity.py:55,82,107).
    """
    n = target_feats.shape[0]
    if k > n or k < 0:
        raise RuntimeError("selected index k out of range")
    vals, idx = torch.sort(target_feats, dim=0, descending=True, stable=True)
    return vals[:k].contiguous(), idx[:k].contiguous()


def p_ramp(top_k: int, p_start: float, p_end: float) -> torch.Tensor:
    """The rank weights of similarity.py:58 with the same fp32 rounding sequence."""
    steps = torch.arange(start=0, end=top_k) / top_k * (p_start - p_end)
    return p_start - steps.unsqueeze(1)


def _lme_over_neurons(scores: torch.Tensor) -> torch.Tensor:
    """log-mean-exp over the neuron axis, similarity.py:70-71 / :93-94."""
    count = scores.shape[0] * torch.ones([1])      # fp32 even for fp64 scores, as in the reference
    return torch.logsumexp(scores, dim=0, keepdim=True) - torch.log(count)


# --------------------------------------------------------------------------------------
# soft-WPMI / WPMI  -- loop form (faithful op sequence; this is what bench.py times)
# --------------------------------------------------------------------------------------
def soft_wpmi(clip_feats, target_feats, top_k=100, a=10, lam=1, min_prob=1e-7,
              p_start=0.998, p_end=0.97, inds: Optional[torch.Tensor] = None,
              return_parts: bool = False):
    """similarity.py:49-73, one neuron at a time (same tensor ops per neuron)."""
    with torch.no_grad():
        probs = torch.softmax(a * clip_feats, dim=1)
        if inds is None:
            inds = topk_cols(target_feats, top_k)[1]
        weights = p_ramp(top_k, p_start, p_end).to(probs.dtype)
        rows = []
        for j in range(target_feats.shape[1]):
            picked = probs.index_select(0, inds[:, j])          # == gather of whole rows
            mixed = 1 + weights * (picked - 1)
            rows.append(torch.sum(torch.log(mixed + min_prob), dim=0, keepdim=True))
        log_sums = torch.cat(rows, dim=0)
        out = log_sums - lam * _lme_over_neurons(log_sums)
    if return_parts:
        return out, log_sums, inds
    return out


def wpmi(clip_feats, target_feats, top_k=28, a=2, lam=0.6, min_prob=1e-7,
         inds: Optional[torch.Tensor] = None, return_parts: bool = False):
    """similarity.py:75-97."""
    with torch.no_grad():
        probs = torch.softmax(a * clip_feats, dim=1)
        if inds is None:
            inds = topk_cols(target_feats, top_k)[1]
        rows = []
        for j in range(target_feats.shape[1]):
            picked = probs.index_select(0, inds[:, j])
            rows.append(torch.sum(torch.log(picked + min_prob), dim=0, keepdim=True))
        log_sums = torch.cat(rows, dim=0)
        out = log_sums - lam * _lme_over_neurons(log_sums)
    if return_parts:
        return out, log_sums, inds
    return out


# --------------------------------------------------------------------------------------
# chunked forms (same per-element arithmetic, many neurons per torch call) for mid sizes
# --------------------------------------------------------------------------------------
def log_sums_chunked(probs, inds, weights=None, min_prob=1e-7, chunk=64):
    """L[j,c] = sum_r log(1 + w_r (S[inds[r,j],c] - 1) + eps)   (w=None: log(S+eps))."""
    k, K = inds.shape
    out = torch.empty(K, probs.shape[1], dtype=probs.dtype)
    for j0 in range(0, K, chunk):
        sel = inds[:, j0:j0 + chunk]                                   # [k, m]
        picked = probs[sel.reshape(-1)].reshape(k, sel.shape[1], -1)   # [k, m, C]
        if weights is not None:
            picked = 1 + weights.reshape(k, 1, 1) * (picked - 1)
        out[j0:j0 + sel.shape[1]] = torch.log(picked + min_prob).sum(dim=0)
    return out


def soft_wpmi_fast(clip_feats, target_feats, top_k=100, a=10, lam=1, min_prob=1e-7,
                   p_start=0.998, p_end=0.97, inds=None, dtype=None, return_parts=False):
    """Chunked soft-WPMI; pass dtype=torch.float64 for the fp64 'truth' run."""
    with torch.no_grad():
        if dtype is not None:
            clip_feats = clip_feats.to(dtype)
        probs = torch.softmax(a * clip_feats, dim=1)
        if inds is None:
            inds = topk_cols(target_feats, top_k)[1]
        weights = p_ramp(top_k, p_start, p_end).to(probs.dtype)
        log_sums = log_sums_chunked(probs, inds, weights, min_prob)
        out = log_sums - lam * _lme_over_neurons(log_sums)
    if return_parts:
        return out, log_sums, inds
    return out


def wpmi_fast(clip_feats, target_feats, top_k=28, a=2, lam=0.6, min_prob=1e-7,
              inds=None, dtype=None, return_parts=False):
    with torch.no_grad():
        if dtype is not None:
            clip_feats = clip_feats.to(dtype)
        probs = torch.softmax(a * clip_feats, dim=1)
        if inds is None:
            inds = topk_cols(target_feats, top_k)[1]
        log_sums = log_sums_chunked(probs, inds, None, min_prob)
        out = log_sums - lam * _lme_over_neurons(log_sums)
    if return_parts:
        return out, log_sums, inds
    return out


# --------------------------------------------------------------------------------------
# log-sum-exp in fixed neuron blocks (the product's G-invariant combine, restated)
# --------------------------------------------------------------------------------------
def lse_block_partials(log_sums: torch.Tensor, block: int = LSE_BLOCK):
    """Per block of `block` neurons: (max[c], sum_j exp(L[j,c]-max[c])) -> [nb,2,C]."""
    K, C = log_sums.shape
    nb = (K + block - 1) // block
    out = torch.empty(nb, 2, C, dtype=log_sums.dtype)
    for b in range(nb):
        seg = log_sums[b * block:(b + 1) * block]
        m = seg.max(dim=0).values
        out[b, 0] = m
        out[b, 1] = torch.exp(seg - m).sum(dim=0)
    return out


def lse_combine(partials: torch.Tensor) -> torch.Tensor:
    """Combine [nb,2,C] partials in block order (fp64 accumulation) -> lse[C] fp64."""
    m = partials[:, 0].double()
    s = partials[:, 1].double()
    big = m.max(dim=0).values
    big_safe = torch.where(torch.isinf(big), torch.zeros_like(big), big)
    total = torch.zeros_like(big)
    for b in range(partials.shape[0]):
        total = total + s[b] * torch.exp(m[b] - big_safe)
    return big_safe + torch.log(total)


# --------------------------------------------------------------------------------------
# cosine similarities
# --------------------------------------------------------------------------------------
def cos_similarity_cubed(clip_feats, target_feats, batch_size=10000, min_norm=1e-3):
    """similarity.py:7-31: centre columns, cube, column-normalise (clipped), A^T P."""
    with torch.no_grad():
        c = clip_feats - clip_feats.mean(dim=0, keepdim=True)
        t = target_feats - target_feats.mean(dim=0, keepdim=True)
        c = c ** 3
        t = t ** 3
        c = c / torch.clip(torch.norm(c, p=2, dim=0, keepdim=True), min_norm)
        t = t / torch.clip(torch.norm(t, p=2, dim=0, keepdim=True), min_norm)
        return _blocked_at_b(t, c, batch_size)


def cos_similarity(clip_feats, target_feats):
    """similarity.py:33-47 (no clip on the norm: zero columns give NaN, as there)."""
    with torch.no_grad():
        c = clip_feats / torch.norm(clip_feats, p=2, dim=0, keepdim=True)
        t = target_feats / torch.norm(target_feats, p=2, dim=0, keepdim=True)
        return _blocked_at_b(t, c, 10000)


def _blocked_at_b(t, c, bs):
    blocks = []
    for i in range(math.ceil(t.shape[1] / bs)):
        left = t[:, i * bs:(i + 1) * bs].T
        blocks.append(torch.cat([left @ c[:, j * bs:(j + 1) * bs]
                                 for j in range(math.ceil(c.shape[1] / bs))], dim=1))
    return torch.cat(blocks, dim=0)


# --------------------------------------------------------------------------------------
# rank_reorder (random in the reference: 5 x torch.randperm per neuron, global CPU RNG)
# --------------------------------------------------------------------------------------
def rank_reorder(clip_feats, target_feats, p=3, top_fraction=0.05, scale_p=0.5,
                 perms: Optional[Callable[[int, int], torch.Tensor]] = None):
    """similarity.py:99-132.  `perms(neuron, n)` -> LongTensor[5, n]; default draws
    torch.randperm(n) five times per neuron from the global RNG in reference order."""
    with torch.no_grad():
        top_n = int(target_feats.shape[0] * top_fraction)
        vals, inds = topk_cols(target_feats, top_n)
        rows = []
        for j in range(target_feats.shape[1]):
            picked = clip_feats.index_select(0, inds[:, j])            # [top_n, C] raw P
            avg = picked.mean(dim=0, keepdim=True)
            # rank 0 = smallest.  The reference's torch.argsort is not stable, so ties among the gathered cosines
            # are ordered arbitrarily there; the stated rule (here and in the CUDA path) is: ties by row position.
            ranks = torch.argsort(torch.argsort(picked, dim=0, stable=True), dim=0, stable=True)
            tgt = vals[:, j:j + 1]                                      # descending
            asc = torch.flip(tgt, dims=[0])
            if perms is None:
                shuffled = torch.cat([asc[torch.randperm(len(asc))] for _ in range(5)], dim=1)
            else:
                shuffled = torch.cat([asc[q] for q in perms(j, len(asc))], dim=1)
            baseline = torch.mean(torch.abs(asc - shuffled) ** p)
            reorg = asc.expand(-1, ranks.shape[1]).gather(0, ranks)
            err = torch.mean(torch.abs(tgt - reorg) ** p, dim=0, keepdim=True) / baseline
            rows.append(err / avg ** scale_p)
        return -torch.cat(rows, dim=0)


# --------------------------------------------------------------------------------------
# image x text similarity matrix and the pooling hook
# --------------------------------------------------------------------------------------
def similarity_matrix(image_features, text_features):
    """CLIP_og_utils.py:155-160: float(), row-normalise, I @ T.T (inputs untouched)."""
    with torch.no_grad():
        i = image_features.float()
        t = text_features.float()
        i = i / i.norm(dim=-1, keepdim=True)
        t = t / t.norm(dim=-1, keepdim=True)
        return i @ t.T


def get_activation(outputs: List[torch.Tensor], mode: str):
    """utils.py:27-52.  avg: unwrap tuple, NCHW -> spatial mean, [B,T,D] -> CLS token,
    [B,D] passthrough.  max: the same with amax and WITHOUT the tuple unwrap."""
    if mode not in ("avg", "max"):
        # the reference falls through to `return hook` with hook unbound (UnboundLocalError)
        raise ValueError("mode must be 'avg' or 'max'")

    def hook(module, inputs, output):
        if mode == "avg" and type(output) is tuple:
            output = output[0]
        nd = len(output.shape)
        if nd == 4:
            pooled = output.mean(dim=[2, 3]) if mode == "avg" else output.amax(dim=[2, 3])
            outputs.append(pooled.detach())
        elif nd == 3:
            outputs.append(output[:, 0].clone())
        elif nd == 2:
            outputs.append(output.detach())
    return hook
