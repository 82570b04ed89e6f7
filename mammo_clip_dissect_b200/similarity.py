"""B200-native neuron->concept scoring ops with the call surface of the reference's
``concept_vit/similarity.py`` (same names, positional order, defaults and return shape), so
``eval("similarity.<name>")`` in the reference drivers (describe_clip_neurons.py:41) resolves to
these functions unchanged.

    soft_wpmi            reference similarity.py:49-73
    wpmi                 reference similarity.py:75-97
    cos_similarity_cubed reference similarity.py:7-31
    cos_similarity       reference similarity.py:33-47

The arithmetic runs in hand-written sm_100a kernels behind the C ABI of include/mcd_b200.h
(ctypes, see _lib.py).  PyTorch is used for device buffers and the current stream only.
There is no CPU path: ``device`` must name a CUDA device and the library must be built.

Differences from the reference, all supersets or stated rules:
  * top-k ties: value desc, then probe-image index asc; NaN largest; -0.0 == +0.0
    (torch.topk leaves both tie order and tie membership unspecified);
  * ``top_k`` is accepted and ignored by the cos functions (the reference's utils.py:602 passes it
    to every similarity function and its cos functions raise TypeError on it);
  * nothing is printed, no tqdm bar, no empty_cache() calls;
  * numerics of the log-sums: the term 1 + p(s - 1) + eps is one FMA and the terms of 4 consecutive ranks are multiplied
    before one MUFU lg2 (S is this module's own softmax, entries in [0, 1]): ~5e-8 relative on L against the reference's
    fp32 operation order, the same size as the reference's own distance from its fp64 run (stated tolerance 1e-5; DESIGN.md
    section 3).  `log_sums(..., probabilities=False)` / `mcd_wpmi_accum_f32` evaluate in the reference's order.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

__all__ = ["soft_wpmi", "wpmi", "cos_similarity", "cos_similarity_cubed", "cos_similarity_cubed_single", "rank_reorder",
           "soft_wpmi_layers", "wpmi_layers", "soft_wpmi_top", "wpmi_top", "top_concepts", "topk_cols",
           "concept_probabilities", "pmi_scores", "pmi_logsums"]

_S_ALIGN = 32  # leading dimension of the probability matrix: rows start on 128-byte boundaries

# Optional per-stage timing for bench.py: set PROFILE = [] and every stage of pmi_scores records a pair
# of CUDA events on the current stream (no synchronisation); profile_summary() reads them back.
PROFILE = None


class _Stage:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.append((self.name, self.e0, e1))
        return False


def profile_summary():
    """Median milliseconds per stage over everything recorded since PROFILE was set (a stage that happens to contain an
    allocation or a clock ramp must not move the figure)."""
    if not PROFILE:
        return {}
    torch.cuda.synchronize()
    acc = {}
    for name, e0, e1 in PROFILE:
        acc.setdefault(name, []).append(e0.elapsed_time(e1))
    return {k: sorted(v)[len(v) // 2] for k, v in acc.items()}


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _cuda_device(device):
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(
            "mammo_clip_dissect_b200 has no CPU path: device=%r; pass a CUDA device (B200, sm_100a)" % (device,))
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def _as_f32_matrix(t, dev, name):
    """fp32, on `dev`, unit stride along the last axis (any row stride). Never modifies the input."""
    if t.dim() != 2:
        raise RuntimeError("%s must be 2-D, got shape %s" % (name, tuple(t.shape)))
    if t.device != dev:
        t = t.to(dev, non_blocking=True)
    if t.dtype != torch.float32:
        t = t.float()
    if t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t


def _ld(t):
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)


def _workspace(nbytes, dev):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)


_CALL_WS = {}          # (device, stream) -> grow-only workspace for the single-entry call, small calls only
_CALL_WS_MAX = 256 << 20


def _call_workspace(nbytes, dev):
    """Workspace of mcd_pmi_scores_f32.  Calls on one stream are ordered, so small workspaces are kept and reused
    (a job scores dozens of small layers back to back); big ones come from the caching allocator per call."""
    if nbytes > _CALL_WS_MAX:
        return torch.empty(nbytes, dtype=torch.uint8, device=dev)
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    buf = _CALL_WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = _CALL_WS[key] = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)
    return buf


# ------------------------------------------------------------------------------------------------
# building blocks (also used by the multi-GPU path and the tests)
# ------------------------------------------------------------------------------------------------
def concept_probabilities(clip_feats, a, device="cuda"):
    """S = softmax(a * clip_feats, dim=1) as a [N, C] view of a row-padded buffer (K1b)."""
    dev = _cuda_device(device)
    P = _as_f32_matrix(clip_feats, dev, "clip_feats")
    N, C = P.shape
    if N < 1 or C < 1:
        raise RuntimeError("clip_feats must be non-empty")
    lds = (C + _S_ALIGN - 1) // _S_ALIGN * _S_ALIGN
    buf = torch.empty((N, lds), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().mcd_softmax_rows_f32(_ptr(P), _ld(P), _ptr(buf), lds, N, C, float(a), _stream(dev)),
                   "mcd_softmax_rows_f32")
    return buf[:, :C]


def topk_cols(target_feats, k, device="cuda", want_values=False, want_int32=False):
    """Per-column top-k over axis 0 (K2).  Returns int64 indices [k, K] (like torch.topk(..., dim=0)[1]);
    optionally (values, indices) and/or an extra int32 copy used by the accumulate kernel."""
    dev = _cuda_device(device)
    A = _as_f32_matrix(target_feats, dev, "target_feats")
    N, K = A.shape
    k = int(k)
    if k < 1 or k > N:
        raise RuntimeError("selected index k out of range")   # torch.topk's message for k > N
    if K < 1:
        raise RuntimeError("target_feats has no neurons")
    lib = _lib.lib()
    idx64 = torch.empty((k, K), dtype=torch.int64, device=dev)
    idx32 = torch.empty((k, K), dtype=torch.int32, device=dev) if want_int32 else None
    vals = torch.empty((k, K), dtype=torch.float32, device=dev) if want_values else None
    need = lib.mcd_topk_cols_workspace_bytes(N, K, k)
    if need == 0:
        raise RuntimeError("top_k=%d is outside the supported range of the column top-k kernel (<= 16384)" % k)
    ws = _workspace(need, dev)
    with torch.cuda.device(dev):
        _lib.check(lib.mcd_topk_cols_f32(_ptr(A), _ld(A), N, K, k, _ptr(idx64), _ptr(idx32), _ptr(vals), _ptr(ws),
                                         ws.numel(), _stream(dev)), "mcd_topk_cols_f32")
    out = (vals, idx64) if want_values else idx64
    return (out, idx32) if want_int32 else out


def _topk_int32(A, k, dev):
    """int32 indices only (what K3 consumes)."""
    N, K = A.shape
    if k < 1 or k > N:
        raise RuntimeError("selected index k out of range")
    lib = _lib.lib()
    idx32 = torch.empty((k, K), dtype=torch.int32, device=dev)
    need = lib.mcd_topk_cols_workspace_bytes(N, K, k)
    if need == 0:
        raise RuntimeError("top_k=%d is outside the supported range of the column top-k kernel (<= 16384)" % k)
    with torch.cuda.device(dev):
        ws = _workspace(need, dev)
        _lib.check(lib.mcd_topk_cols_f32(_ptr(A), _ld(A), N, K, k, None, _ptr(idx32), None, _ptr(ws), ws.numel(),
                                         _stream(dev)), "mcd_topk_cols_f32")
    return idx32


def _reference_ramp(top_k, p_start, p_end):
    """The rank weights with the reference's exact fp32 rounding sequence (similarity.py:58);
    evaluated with the same host-side tensor expression, then shipped to the device."""
    steps = torch.arange(start=0, end=top_k) / top_k * (p_start - p_end)
    return (p_start - steps).to(torch.float32).contiguous()


_ramp_cache = {}


def _device_ramp(ramp, top_k, p_start, p_end, dev):
    """The ramp on the device, cached per (top_k, p_start, p_end, device): a job scores dozens of layers with the same
    arguments, and a pageable host-to-device copy per call would be a host synchronisation per call."""
    key = (int(top_k), float(p_start), float(p_end), str(dev))
    t = _ramp_cache.get(key)
    if t is None:
        if len(_ramp_cache) > 64:
            _ramp_cache.clear()
        t = _ramp_cache[key] = ramp.to(dev)
    return t


def log_sums(S, idx32, weights, min_prob, out=None, probabilities=True):
    """K3: L[j,c] = sum_r log(1 + w_r (S[idx[r,j],c]-1) + eps)  (weights=None: sum_r log(S+eps)).
    probabilities=True (S is a softmax output, entries in [0, 1]): grouped-log evaluation (mcd_wpmi_accum_prob_f32);
    False: any S, the reference's operation order per term (mcd_wpmi_accum_f32)."""
    dev = S.device
    N, C = S.shape
    k, K = idx32.shape
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty((K, C), dtype=torch.float32, device=dev)
        fn = _lib.lib().mcd_wpmi_accum_prob_f32 if probabilities else _lib.lib().mcd_wpmi_accum_f32
        _lib.check(fn(_ptr(S), _ld(S), N, C, _ptr(idx32), K, k, _ptr(weights), float(min_prob), _ptr(out), _ld(out),
                      _stream(dev)), "mcd_wpmi_accum_f32")
    return out


def lse_partials(L):
    """K3b part 1: per-256-neuron-block (max, sum exp) partials [nb, 2, C]."""
    K, C = L.shape
    nb = (K + _lib.LSE_BLOCK - 1) // _lib.LSE_BLOCK
    with torch.cuda.device(L.device):
        part = torch.empty((nb, 2, C), dtype=torch.float32, device=L.device)
        _lib.check(_lib.lib().mcd_col_lse_partials_f32(_ptr(L), _ld(L), K, C, _ptr(part), _stream(L.device)),
                   "mcd_col_lse_partials_f32")
    return part


def pmi_finalize(L, partials_all, K_total, lam, out=None):
    """K3b part 2: out = L - lam * (logsumexp over ALL blocks - log K_total).  In place by default."""
    K, C = L.shape
    if out is None:
        out = L
    with torch.cuda.device(L.device):
        prob_d = torch.empty((C,), dtype=torch.float32, device=L.device)
        _lib.check(_lib.lib().mcd_pmi_finalize_f32(_ptr(L), _ld(L), K, C, _ptr(partials_all), partials_all.shape[0],
                                                   int(K_total), float(lam), _ptr(prob_d), _ptr(out), _ld(out),
                                                   _stream(L.device)), "mcd_pmi_finalize_f32")
    return out, prob_d


def pmi_finalize_top(L, partials_all, K_total, lam, t, out=None):
    """K3b part 2 fused with the per-neuron top concepts: (out, log p(d) [C], top values [K, t], top concept indices
    [K, t] int64), out = L - lam * log p(d) in place by default; one pass over the matrix (mcd_pmi_finalize_topk_f32)."""
    K, C = L.shape
    if out is None:
        out = L
    t = int(t)
    if t < 1 or t > min(C, 64) or C > 1024:
        raise RuntimeError("top concepts: 1 <= t <= min(C, 64) and C <= 1024 (got t=%d, C=%d)" % (t, C))
    with torch.cuda.device(L.device):
        prob_d = torch.empty((C,), dtype=torch.float32, device=L.device)
        vals = torch.empty((K, t), dtype=torch.float32, device=L.device)
        idx = torch.empty((K, t), dtype=torch.int64, device=L.device)
        _lib.check(_lib.lib().mcd_pmi_finalize_topk_f32(_ptr(L), _ld(L), K, C, _ptr(partials_all), partials_all.shape[0],
                                                        int(K_total), float(lam), _ptr(prob_d), _ptr(out), _ld(out), t,
                                                        _ptr(vals), _ptr(idx), _stream(L.device)),
                   "mcd_pmi_finalize_topk_f32")
    return out, prob_d, vals, idx


def pmi_finalize_bcast(L, partials_all, K_total, lam, dest_ptrs, row_offset):
    """K3b fused with the score all-gather: finalize this rank's contiguous [K, C] log-sums and store the slice into
    rows [row_offset, row_offset + K) of every [K_total, C] matrix in dest_ptrs (device pointers: the peers' symmetric
    buffers and this rank's own).  Returns log p(d) [C]."""
    import ctypes
    K, C = L.shape
    if not L.is_contiguous():
        raise RuntimeError("pmi_finalize_bcast needs contiguous log-sums")
    with torch.cuda.device(L.device):
        prob_d = torch.empty((C,), dtype=torch.float32, device=L.device)
        arr = (ctypes.c_void_p * len(dest_ptrs))(*[int(p) for p in dest_ptrs])
        _lib.check(_lib.lib().mcd_pmi_finalize_bcast_f32(_ptr(L), K, C, _ptr(partials_all), partials_all.shape[0],
                                                         int(K_total), float(lam), _ptr(prob_d), arr, len(dest_ptrs),
                                                         int(row_offset), _stream(L.device)), "mcd_pmi_finalize_bcast_f32")
    return prob_d


def bcast_rows(src, dest_ptrs, dest_offset):
    """src (contiguous fp32 CUDA tensor) to element offset `dest_offset` of every buffer in dest_ptrs (device pointers:
    the peers' symmetric buffers and this rank's own)."""
    import ctypes
    if not src.is_contiguous() or src.dtype != torch.float32 or not src.is_cuda:
        raise RuntimeError("bcast_rows needs a contiguous fp32 CUDA tensor")
    if src.numel() == 0:
        return
    with torch.cuda.device(src.device):
        arr = (ctypes.c_void_p * len(dest_ptrs))(*[int(p) for p in dest_ptrs])
        _lib.check(_lib.lib().mcd_bcast_f32(_ptr(src), src.numel(), arr, len(dest_ptrs), int(dest_offset), _stream(src.device)),
                   "mcd_bcast_f32")


def pmi_logsums(clip_feats, target_feats, top_k, a, device, min_prob, ramp):
    """(L [K, C], partials [ceil(K/256), 2, C]) of one device's neurons behind one FFI entry point
    (mcd_pmi_logsums_f32: softmax -> column top-k -> gather / log-sum -> block partials, column chunks pipelined) --
    what a rank of the neuron-sharded call computes before the partials are exchanged."""
    dev = _cuda_device(device)
    with torch.no_grad(), torch.cuda.device(dev):
        A = _as_f32_matrix(target_feats, dev, "target_feats")
        P = _as_f32_matrix(clip_feats, dev, "clip_feats")
        if P.shape[0] != A.shape[0]:
            raise RuntimeError("clip_feats %s and target_feats %s must share the probe-image axis"
                               % (tuple(P.shape), tuple(A.shape)))
        N, C = P.shape
        K = A.shape[1]
        top_k = int(top_k)
        if C < 1 or K < 1:
            raise RuntimeError("clip_feats / target_feats must be non-empty")
        if top_k < 1 or top_k > N:
            raise RuntimeError("selected index k out of range")
        lib = _lib.lib()
        need = int(lib.mcd_pmi_scores_workspace_bytes(N, K, C, top_k))
        if need == 0:
            raise RuntimeError("top_k=%d is outside the supported range of the column top-k kernel (<= 16384)" % top_k)
        ws = _call_workspace(need, dev)
        weights = ramp.to(dev) if ramp is not None else None
        L = torch.empty((K, C), dtype=torch.float32, device=dev)
        part = torch.empty(((K + _lib.LSE_BLOCK - 1) // _lib.LSE_BLOCK, 2, C), dtype=torch.float32, device=dev)
        _lib.check(lib.mcd_pmi_logsums_f32(_ptr(P), _ld(P), _ptr(A), _ld(A), N, K, C, top_k, float(a), _ptr(weights),
                                           float(min_prob), _ptr(L), _ld(L), _ptr(part), _ptr(ws), ws.numel(),
                                           _stream(dev)), "mcd_pmi_logsums_f32")
    return L, part


def pmi_scores(clip_feats, target_feats, top_k, a, lam, device, min_prob, ramp, return_parts=False):
    """Shared body of soft_wpmi / wpmi on one device."""
    dev = _cuda_device(device)
    with torch.no_grad(), torch.cuda.device(dev):
        A = _as_f32_matrix(target_feats, dev, "target_feats")
        if clip_feats.dim() != 2 or clip_feats.shape[0] != A.shape[0]:
            raise RuntimeError("clip_feats %s and target_feats %s must share the probe-image axis"
                               % (tuple(clip_feats.shape), tuple(A.shape)))
        top_k = int(top_k)
        if PROFILE is None and not return_parts:
            # the whole call behind one FFI entry point (same kernels, intermediates in one workspace): a small layer
            # is launch-bound from Python otherwise
            P = _as_f32_matrix(clip_feats, dev, "clip_feats")
            N, C = P.shape
            K = A.shape[1]
            if C < 1 or K < 1:
                raise RuntimeError("clip_feats / target_feats must be non-empty")
            if top_k < 1 or top_k > N:
                raise RuntimeError("selected index k out of range")
            lib = _lib.lib()
            need = int(lib.mcd_pmi_scores_workspace_bytes(N, K, C, top_k))
            if need == 0:
                raise RuntimeError("top_k=%d is outside the supported range of the column top-k kernel (<= 16384)" % top_k)
            ws = _call_workspace(need, dev)
            weights = ramp.to(dev) if ramp is not None else None
            out = torch.empty((K, C), dtype=torch.float32, device=dev)
            _lib.check(lib.mcd_pmi_scores_f32(_ptr(P), _ld(P), _ptr(A), _ld(A), N, K, C, top_k, float(a), float(lam),
                                              _ptr(weights), float(min_prob), _ptr(out), _ld(out), _ptr(ws), ws.numel(),
                                              _stream(dev)), "mcd_pmi_scores_f32")
            return out
        with _Stage("softmax_rows"):
            S = concept_probabilities(clip_feats, a, dev)
        with _Stage("topk_cols"):
            idx32 = _topk_int32(A, top_k, dev)
        weights = ramp.to(dev) if ramp is not None else None      # no-op for a cached device ramp
        with _Stage("wpmi_accum"):
            L = log_sums(S, idx32, weights, min_prob)
        if return_parts:
            raw = L.clone()
        with _Stage("lse_finalize"):
            part = lse_partials(L)
            out, _ = pmi_finalize(L, part, A.shape[1], lam)
    if return_parts:
        return out, raw, idx32
    return out


_seg_cache = {}


def _segment_tables(Ks, dev):
    """Device tables for the segmented K3b kernels (cached per layer-width tuple): every layer's 256-neuron LSE
    blocks start at the layer's first row, exactly as in a separate call."""
    key = (tuple(int(k) for k in Ks), str(dev))
    hit = _seg_cache.get(key)
    if hit is not None:
        return hit
    import math
    blocks, segs, logs, row, row_seg = [], [], [], 0, []
    for s_id, K in enumerate(key[0]):
        first = len(blocks)
        for j0 in range(0, K, _lib.LSE_BLOCK):
            blocks.append((row + j0, min(_lib.LSE_BLOCK, K - j0), s_id))
        segs.append((first, len(blocks) - first))
        logs.append(math.log(float(K)))          # libm log of a double, like the single-layer entry point
        row += K
        row_seg.append(torch.full((K,), s_id, dtype=torch.int32))
    out = (torch.tensor(blocks, dtype=torch.int32).to(dev), torch.tensor(segs, dtype=torch.int32).to(dev),
           torch.tensor(logs, dtype=torch.float64).to(dev), len(blocks), torch.cat(row_seg).to(dev))
    if len(_seg_cache) > 16:
        _seg_cache.clear()
    _seg_cache[key] = out
    return out


def pmi_scores_layers(clip_feats, target_feats_list, top_k, a, lam, device, min_prob, ramp, top_concepts=None):
    """[pmi_scores(clip_feats, t, ...) for t in target_feats_list], bit for bit, with ONE pass of every kernel over
    the layers stacked along the neuron axis: one softmax, one column top-k and one
    gather/log-sum over [N, sum K_l], and log p(d) per layer from segment-wise 256-neuron blocks."""
    dev = _cuda_device(device)
    stacked = hasattr(target_feats_list, "matrix") and hasattr(target_feats_list, "widths")   # hooks.ActivationStack
    if not stacked and len(target_feats_list) == 0:
        return []
    with torch.no_grad(), torch.cuda.device(dev):
        if stacked:
            # the layers already sit side by side in one [N, sum K_l] matrix: nothing to concatenate
            if hasattr(target_feats_list, "complete") and not target_feats_list.complete():
                raise RuntimeError("ActivationStack is not full: some layer has not seen all probe images")
            A = _as_f32_matrix(target_feats_list.matrix, dev, "target_feats")
            Ks = [int(w) for w in target_feats_list.widths]
            if sum(Ks) != A.shape[1] or min(Ks) < 1:
                raise RuntimeError("layer widths %s do not add up to the stacked matrix width %d" % (Ks, A.shape[1]))
            N = A.shape[0]
        else:
            mats = [_as_f32_matrix(t, dev, "target_feats") for t in target_feats_list]
            N = mats[0].shape[0]
            for t in mats:
                if t.shape[0] != N or t.shape[1] < 1:
                    raise RuntimeError("every layer must be [N, K_l] with the same N and K_l >= 1")
            Ks = [t.shape[1] for t in mats]
            A = mats[0] if len(mats) == 1 else torch.cat(mats, dim=1)
        if clip_feats.dim() != 2 or clip_feats.shape[0] != N:
            raise RuntimeError("clip_feats %s and target_feats [%d, .] must share the probe-image axis"
                               % (tuple(clip_feats.shape), N))
        with _Stage("softmax_rows"):
            S = concept_probabilities(clip_feats, a, dev)
        with _Stage("topk_cols"):
            idx32 = _topk_int32(A, int(top_k), dev)
        with _Stage("wpmi_accum"):
            L = log_sums(S, idx32, ramp, min_prob)
        C = L.shape[1]
        block_tab, seg_tab, seg_log, n_blocks, row_seg = _segment_tables(Ks, dev)
        lib = _lib.lib()
        with _Stage("lse_finalize"):
            part = torch.empty((n_blocks, 2, C), dtype=torch.float32, device=dev)
            prob_d = torch.empty((len(Ks), C), dtype=torch.float32, device=dev)
            _lib.check(lib.mcd_col_lse_partials_seg_f32(_ptr(L), _ld(L), C, _ptr(block_tab), n_blocks, _ptr(part),
                                                        _stream(dev)), "mcd_col_lse_partials_seg_f32")
            if top_concepts is None:
                _lib.check(lib.mcd_pmi_finalize_seg_f32(_ptr(L), _ld(L), C, _ptr(part), _ptr(block_tab), n_blocks,
                                                        _ptr(seg_tab), _ptr(seg_log), len(Ks), float(lam), _ptr(prob_d),
                                                        _ptr(L), _ld(L), _stream(dev)), "mcd_pmi_finalize_seg_f32")
            else:
                t = int(top_concepts)
                if t < 1 or t > min(C, 64) or C > 1024:
                    raise RuntimeError("top concepts: 1 <= t <= min(C, 64) and C <= 1024 (got t=%d, C=%d)" % (t, C))
                Kt = L.shape[0]
                vals = torch.empty((Kt, t), dtype=torch.float32, device=dev)
                idx = torch.empty((Kt, t), dtype=torch.int64, device=dev)
                _lib.check(lib.mcd_pmi_finalize_seg_topk_f32(_ptr(L), _ld(L), Kt, C, _ptr(part), _ptr(row_seg), n_blocks,
                                                             _ptr(seg_tab), _ptr(seg_log), len(Ks), float(lam),
                                                             _ptr(prob_d), _ptr(L), _ld(L), t, _ptr(vals), _ptr(idx),
                                                             _stream(dev)), "mcd_pmi_finalize_seg_topk_f32")
    if top_concepts is not None:
        return list(torch.split(L, Ks, dim=0)), list(torch.split(vals, Ks, dim=0)), list(torch.split(idx, Ks, dim=0))
    return list(torch.split(L, Ks, dim=0))


def soft_wpmi_layers(clip_feats, target_feats_list, top_k=100, a=10, lam=1, device='cuda',
                     min_prob=1e-7, p_start=0.998, p_end=0.97, top_concepts=None):
    """soft_wpmi for all layers of a model in one pass: returns [soft_wpmi(clip_feats, t, ...) for t in the list]
    (views of one [sum K_l, C] matrix), identical bits.  The reference scores layer by layer
    (describe_broad_neurons.py:83-119); this is the same loop with the per-call work done once.
    top_concepts=t: also the t best (value, concept) pairs of every neuron, emitted by the finalize pass itself -- what
    the caller computes next with torch.topk(similarities, k=10, dim=1) (describe_broad_neurons.py:101); returns
    (scores per layer, top values per layer [K_l, t], top concept indices per layer [K_l, t])."""
    dev = _cuda_device(device)
    ramp = _device_ramp(_reference_ramp(int(top_k), p_start, p_end), top_k, p_start, p_end, dev)
    return pmi_scores_layers(clip_feats, target_feats_list, top_k, a, lam, dev, min_prob, ramp, top_concepts)


def wpmi_layers(clip_feats, target_feats_list, top_k=28, a=2, lam=0.6, device='cuda', min_prob=1e-7, top_concepts=None):
    """wpmi for all layers of a model in one pass (see soft_wpmi_layers)."""
    return pmi_scores_layers(clip_feats, target_feats_list, top_k, a, lam, device, min_prob, None, top_concepts)


def soft_wpmi_top(clip_feats, target_feats, top_concepts=10, top_k=100, a=10, lam=1, device='cuda',
                  min_prob=1e-7, p_start=0.998, p_end=0.97):
    """(soft_wpmi scores [K, C], top values [K, t], top concept indices [K, t]): the scores of soft_wpmi(...) bit for bit,
    with the t best concepts of every neuron emitted by the finalize pass (no second read of the matrix) -- the
    reference's callers' torch.topk(similarities, 10, 1) / torch.max(similarities, 1)."""
    dev = _cuda_device(device)
    ramp = _device_ramp(_reference_ramp(int(top_k), p_start, p_end), top_k, p_start, p_end, dev)
    L, part = pmi_logsums(clip_feats, target_feats, top_k, a, dev, min_prob, ramp)
    out, _, vals, idx = pmi_finalize_top(L, part, L.shape[0], lam, top_concepts)
    return out, vals, idx


def wpmi_top(clip_feats, target_feats, top_concepts=10, top_k=28, a=2, lam=0.6, device='cuda', min_prob=1e-7):
    """wpmi with the per-neuron top concepts (see soft_wpmi_top)."""
    dev = _cuda_device(device)
    L, part = pmi_logsums(clip_feats, target_feats, top_k, a, dev, min_prob, None)
    out, _, vals, idx = pmi_finalize_top(L, part, L.shape[0], lam, top_concepts)
    return out, vals, idx


def top_concepts(scores, k=10):
    """(values [K, k], indices [K, k] int64) of the k best concepts of every neuron, best first -- what the reference's
    callers compute with torch.topk(similarities, k=10, dim=1) (describe_broad_neurons.py:101) or torch.max(similarities,
    1) (describe_clip_neurons.py:64), under the stated order (value desc, concept index asc, NaN largest)."""
    if scores.dim() != 2 or not scores.is_cuda:
        raise RuntimeError("top_concepts expects a 2-D CUDA score matrix (mammo_clip_dissect_b200 has no CPU path)")
    X = _as_f32_matrix(scores, scores.device, "scores")
    K, C = X.shape
    k = int(k)
    if k < 1 or k > C:
        raise RuntimeError("selected index k out of range")
    vals = torch.empty((K, k), dtype=torch.float32, device=X.device)
    idx = torch.empty((K, k), dtype=torch.int64, device=X.device)
    with torch.cuda.device(X.device):
        _lib.check(_lib.lib().mcd_row_topk_f32(_ptr(X), _ld(X), K, C, k, _ptr(vals), _ptr(idx), _stream(X.device)),
                   "mcd_row_topk_f32")
    return vals, idx


# ------------------------------------------------------------------------------------------------
# the reference's call surface
# ------------------------------------------------------------------------------------------------
def soft_wpmi(clip_feats, target_feats, top_k=100, a=10, lam=1, device='cuda',
              min_prob=1e-7, p_start=0.998, p_end=0.97):
    """Soft-WPMI neuron x concept scores [K, C] on `device` (reference similarity.py:49-73)."""
    dev = _cuda_device(device)
    ramp = _device_ramp(_reference_ramp(int(top_k), p_start, p_end), top_k, p_start, p_end, dev)
    return pmi_scores(clip_feats, target_feats, top_k, a, lam, dev, min_prob, ramp)


def wpmi(clip_feats, target_feats, top_k=28, a=2, lam=0.6, device='cuda', min_prob=1e-7):
    """WPMI neuron x concept scores [K, C] on `device` (reference similarity.py:75-97)."""
    return pmi_scores(clip_feats, target_feats, top_k, a, lam, device, min_prob, None)


def _cos(clip_feats, target_feats, device, cubed, min_norm):
    dev = _cuda_device(device)
    lib = _lib.lib()
    with torch.no_grad(), torch.cuda.device(dev):
        P = _as_f32_matrix(clip_feats, dev, "clip_feats")
        A = _as_f32_matrix(target_feats, dev, "target_feats")
        if P.shape[0] != A.shape[0]:
            raise RuntimeError("clip_feats and target_feats must share the probe-image axis")
        N, C = P.shape
        K = A.shape[1]
        out = torch.empty((K, C), dtype=torch.float32, device=dev)
        ws = _call_workspace(int(lib.mcd_cos_similarity_workspace_bytes(N, K, C)), dev)
        _lib.check(lib.mcd_cos_similarity_f32(_ptr(P), _ld(P), _ptr(A), _ld(A), N, K, C, int(cubed), float(min_norm),
                                              _ptr(out), _ld(out), _ptr(ws), ws.numel(), _stream(dev)),
                   "mcd_cos_similarity_f32")
    return out


def last_cos_path():
    """Which kernel the last cos_similarity / cos_similarity_cubed call ran: 'tcgen05' or 'fp32_ffma'."""
    return {1: "tcgen05", 3: "fp32_ffma"}.get(int(_lib.lib().mcd_last_cos_path()), "none")


def cos_similarity_cubed(clip_feats, target_feats, device='cuda', batch_size=10000, min_norm=1e-3, top_k=None):
    """Reference similarity.py:7-31.  `batch_size` only blocked the reference's matmul; ignored."""
    return _cos(clip_feats, target_feats, device, True, min_norm)


def cos_similarity(clip_feats, target_feats, device='cuda', top_k=None):
    """Reference similarity.py:33-47 (zero columns give NaN, as there)."""
    return _cos(clip_feats, target_feats, device, False, 0.0)


def cos_similarity_cubed_single(clip_feats, target_feats, device='cuda', min_norm=1e-3):
    """NOT in the reference tree (named by BASELINE.json's north_star; see SURVEY.md section 8 a-note).
    Convenience: the matched-pair (diagonal) cos^3 similarity for equally-shaped inputs."""
    if tuple(clip_feats.shape) != tuple(target_feats.shape):
        raise RuntimeError("cos_similarity_cubed_single needs equally shaped inputs")
    return torch.diagonal(_cos(clip_feats, target_feats, device, True, min_norm)).clone()


# ---- the reference's RNG stream for rank_reorder -------------------------------------------------------------------
# The reference draws 5 x torch.randperm(top_n) per neuron from the GLOBAL CPU generator (similarity.py:119).  On the CPU
# torch.randperm(n) is a Fisher-Yates shuffle that consumes one raw 32-bit Mersenne-Twister output per step
# (z = draw % (n - i); swap i, i + z; n - 1 steps), so instead of 5 K Python-level randperm calls the generator's raw
# outputs are produced in one vectorised numpy call (same MT19937 state), the generator is advanced by exactly that many
# draws, and the shuffles run on the device (mcd_rank_reorder_draws_f32).  The replay is verified against torch.randperm
# itself the first time it is used; if the installed torch ever draws differently, the per-call loop is used instead.
_MT_N = 624
_replay_ok = None


def _mt_state_to_numpy(state):
    """torch CPU generator state (legacy layout: seed u64, left i32, seeded i32, next u64, state u64[624], ...) ->
    (numpy MT19937 key, pos)."""
    import numpy as np
    raw = state.numpy().tobytes()
    left = int(np.frombuffer(raw, dtype=np.int32, count=1, offset=8)[0])
    nxt = int(np.frombuffer(raw, dtype=np.uint64, count=1, offset=16)[0])
    key = np.frombuffer(raw, dtype=np.uint64, count=_MT_N, offset=24).astype(np.uint32)
    # ATen twists when --left hits 0; after the twist left + next == 625
    pos = _MT_N if left == 1 else nxt
    if not (0 <= pos <= _MT_N) or (left != 1 and left + nxt != _MT_N + 1):
        raise RuntimeError("unexpected CPU generator state layout")
    return key, pos


def _mt_state_from_numpy(state, key, pos):
    import numpy as np
    buf = bytearray(state.numpy().tobytes())
    np.frombuffer(buf, dtype=np.int32, count=1, offset=8)[0] = _MT_N + 1 - pos
    np.frombuffer(buf, dtype=np.uint64, count=1, offset=16)[0] = pos
    np.frombuffer(buf, dtype=np.uint64, count=_MT_N, offset=24)[:] = key.astype(np.uint64)
    return torch.frombuffer(buf, dtype=torch.uint8).clone()


def _raw_draws(count, generator=None):
    """The next `count` raw 32-bit outputs of the (global) CPU generator as a uint32 numpy array; advances the generator."""
    import numpy as np
    gen = torch.default_generator if generator is None else generator
    state = gen.get_state()
    key, pos = _mt_state_to_numpy(state)
    mt = np.random.MT19937()
    st = mt.state
    st["state"]["key"] = key
    st["state"]["pos"] = pos
    mt.state = st
    draws = mt.random_raw(int(count)).astype(np.uint32)
    gen.set_state(_mt_state_from_numpy(state, mt.state["state"]["key"], int(mt.state["state"]["pos"])))
    return draws


def _replay_works():
    """One-time self-check on a scratch generator: raw draws + Fisher-Yates == torch.randperm, and the generator ends in
    the same state."""
    global _replay_ok
    if _replay_ok is None:
        try:
            ok = True
            for seed, n, reps in ((7, 10, 3), (123, 257, 2), (5, 1000, 1)):
                g1, g2 = torch.Generator().manual_seed(seed), torch.Generator().manual_seed(seed)
                torch.randn(5, generator=g1), torch.randn(5, generator=g2)          # not at a fresh-seed state
                want = [torch.randperm(n, generator=g1) for _ in range(reps)]
                d = _raw_draws(reps * (n - 1), g2).reshape(reps, n - 1)
                for r in range(reps):
                    perm = list(range(n))
                    for i in range(n - 1):
                        z = i + int(d[r, i]) % (n - i)
                        perm[i], perm[z] = perm[z], perm[i]
                    ok = ok and perm == want[r].tolist()
                ok = ok and torch.equal(torch.randperm(9, generator=g1), torch.randperm(9, generator=g2))
            _replay_ok = bool(ok)
        except Exception:
            _replay_ok = False
    return _replay_ok


_device_replay_ok = {}
_side_streams = {}


def _side_stream(dev):
    key_ = (dev.type, dev.index)
    if key_ not in _side_streams:
        # high priority: its few CTAs (generator, shuffles) must get SM slots while the rank pass has thousands queued
        _side_streams[key_] = torch.cuda.Stream(device=dev, priority=-1)
    return _side_streams[key_]


def _device_replay_works(dev):
    """One-time self-check per device: the MT19937 kernel continues a (non-fresh) generator state exactly like the host
    replay -- same raw draws across several twists, same final state."""
    import numpy as np
    key_ = (dev.type, dev.index)
    if key_ not in _device_replay_ok:
        try:
            lib = _lib.lib()
            g = torch.Generator().manual_seed(20240229)
            torch.randn(7, generator=g)
            state = g.get_state()
            key, pos = _mt_state_to_numpy(state)
            count = 3 * _MT_N + 101
            st_dev = torch.from_numpy(np.append(key, np.uint32(pos)).view(np.int32)).to(dev)
            draws = torch.empty((count,), dtype=torch.int32, device=dev)
            _lib.check(lib.mcd_mt19937_draws(_ptr(st_dev), count, _ptr(draws), _stream(dev)), "mcd_mt19937_draws")
            want = _raw_draws(count, g)
            after_key, after_pos = _mt_state_to_numpy(g.get_state())
            got_state = st_dev.cpu().numpy().view(np.uint32)
            ok = np.array_equal(draws.cpu().numpy().view(np.uint32), want)
            ok = ok and np.array_equal(got_state[:_MT_N], after_key) and int(got_state[_MT_N]) == int(after_pos)
            _device_replay_ok[key_] = bool(ok)
        except Exception:
            _device_replay_ok[key_] = False
    return _device_replay_ok[key_]


def rank_reorder(clip_feats, target_feats, device="cuda", p=3, top_fraction=0.05, scale_p=0.5, top_k=None):
    """Reference similarity.py:99-132 on the GPU.  The reference draws 5 x torch.randperm(top_n) per neuron from the
    GLOBAL CPU generator (in neuron order); the same stream is consumed here (see _raw_draws above), so under the same
    torch.manual_seed the two implementations see identical permutations and leave the generator in the same state.
    Limits: top_n = int(N * top_fraction) <= 8192 (N <= 163 840 at the default 5 %) and K <= 65535 per call."""
    import numpy as np
    dev = _cuda_device(device)
    lib = _lib.lib()
    with torch.no_grad(), torch.cuda.device(dev):
        P = _as_f32_matrix(clip_feats, dev, "clip_feats")
        A = _as_f32_matrix(target_feats, dev, "target_feats")
        if P.shape[0] != A.shape[0]:
            raise RuntimeError("clip_feats and target_feats must share the probe-image axis")
        N, C = P.shape
        K = A.shape[1]
        top_n = int(N * top_fraction)
        if top_n < 1:
            raise RuntimeError("rank_reorder: top_fraction selects no probe image")
        if top_n > 8192 or K > 65535:
            raise NotImplementedError("rank_reorder on the B200 path supports top_n <= 8192 and K <= 65535 "
                                      "(got top_n=%d, K=%d)" % (top_n, K))
        out = torch.empty((K, C), dtype=torch.float32, device=dev)
        ws = _workspace(int(lib.mcd_rank_reorder_workspace_bytes(K, top_n, C)), dev)
        count = K * 5 * (top_n - 1)
        pending = None
        main = torch.cuda.current_stream(dev)
        if top_n > 1 and _replay_works() and _device_replay_works(dev):
            # The generator continues on the device (MT19937 kernel) on a side stream, beside the column top-k and the
            # rank pass; its final state comes back while those are being enqueued and is put into the CPU generator
            # before this function returns.
            side = _side_stream(dev)
            gen = torch.default_generator
            state = gen.get_state()
            key, pos = _mt_state_to_numpy(state)
            st_dev = torch.from_numpy(np.append(key, np.uint32(pos)).view(np.int32)).to(dev)
            draws = torch.empty((count,), dtype=torch.int32, device=dev)
            back = torch.empty((_MT_N + 1,), dtype=torch.int32, pin_memory=True)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                _lib.check(lib.mcd_mt19937_draws(_ptr(st_dev), count, _ptr(draws), side.cuda_stream), "mcd_mt19937_draws")
                back.copy_(st_dev, non_blocking=True)
                done = torch.cuda.Event()
                done.record(side)
            pending = (gen, state, back, done, side, draws)
        (vals, _), idx32 = topk_cols(A, top_n, dev, want_values=True, want_int32=True)
        if pending is not None:
            side, draws = pending[4], pending[5]
            side.wait_stream(main)                                  # the baseline needs the top-n activations
            with torch.cuda.stream(side):
                _lib.check(lib.mcd_rank_baseline_draws_f32(_ptr(vals), K, top_n, C, _ptr(draws), float(p), _ptr(ws), ws.numel(),
                                                           side.cuda_stream), "mcd_rank_baseline_draws_f32")
        elif top_n > 1 and _replay_works():
            draws = torch.from_numpy(_raw_draws(count).view(np.int32)).to(dev)
            _lib.check(lib.mcd_rank_baseline_draws_f32(_ptr(vals), K, top_n, C, _ptr(draws), float(p), _ptr(ws), ws.numel(),
                                                       _stream(dev)), "mcd_rank_baseline_draws_f32")
        else:
            # the reference's RNG stream call by call: for every neuron, five permutations of range(top_n)
            perms = torch.stack([torch.stack([torch.randperm(top_n) for _ in range(5)]) for _ in range(K)]).to(torch.int32)
            perms = perms.to(dev)
            _lib.check(lib.mcd_rank_baseline_perms_f32(_ptr(vals), K, top_n, C, _ptr(perms), float(p), _ptr(ws), ws.numel(),
                                                       _stream(dev)), "mcd_rank_baseline_perms_f32")
        _lib.check(lib.mcd_rank_errors_f32(_ptr(P), _ld(P), N, C, _ptr(idx32), _ptr(vals), K, top_n, float(p), float(scale_p),
                                           _ptr(ws), ws.numel(), _ptr(out), _ld(out), _stream(dev)), "mcd_rank_errors_f32")
        if pending is not None:
            main.wait_stream(pending[4])
        _lib.check(lib.mcd_rank_finish_f32(K, C, top_n, _ptr(ws), ws.numel(), _ptr(out), _ld(out), _stream(dev)),
                   "mcd_rank_finish_f32")
        if pending is not None:
            gen, state, back, done = pending[:4]
            done.synchronize()
            words = back.numpy().view(np.uint32)
            gen.set_state(_mt_state_from_numpy(state, words[:_MT_N], int(words[_MT_N])))
    return out
