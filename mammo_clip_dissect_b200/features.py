"""Image x text similarity matrix over the concept set (reference concept_vit/utils.py:570-594,
CLIP_og_utils.py:155-160, og_utils.py:478-506): float(), L2-normalise the rows of the image and
text features, clip_feats = I @ T.T -- as one call into K1 (include/mcd_b200.h:
mcd_gemm_nt_softmax_f32), optionally fused with the temperature softmax that soft_wpmi / wpmi
apply next.  Inputs are never modified (the reference normalises its loaded copies in place).

``get_similarity_from_activations`` mirrors the reference function of the same name so a driver
can be pointed at it unchanged.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .similarity import _as_f32_matrix, _cuda_device, _ld, _ptr, _stream, _S_ALIGN


def similarity_matrix(image_features, text_features, device="cuda", normalize=True, softmax_scale=None):
    """clip_feats [N, C] = normalise(I) @ normalise(T).T on `device`.
    With softmax_scale=a also returns S = softmax(a * clip_feats, dim=1) as (P, S)."""
    dev = _cuda_device(device)
    lib = _lib.lib()
    with torch.no_grad(), torch.cuda.device(dev):
        I = _as_f32_matrix(image_features, dev, "image_features")
        T = _as_f32_matrix(text_features, dev, "text_features")
        if I.shape[1] != T.shape[1]:
            raise RuntimeError("image features %s and text features %s differ in embedding width"
                               % (tuple(I.shape), tuple(T.shape)))
        N, D = I.shape
        C = T.shape[0]
        if N < 1 or C < 1 or D < 1:
            raise RuntimeError("empty feature matrix")
        P = torch.empty((N, C), dtype=torch.float32, device=dev)
        S = None
        lds = 0
        if softmax_scale is not None:
            lds = (C + _S_ALIGN - 1) // _S_ALIGN * _S_ALIGN
            S = torch.empty((N, lds), dtype=torch.float32, device=dev)
        ws = torch.empty(int(lib.mcd_gemm_nt_softmax_workspace_bytes(N, C, D)), dtype=torch.uint8, device=dev)
        code = lib.mcd_gemm_nt_softmax_f32(_ptr(I), _ld(I), _ptr(T), _ld(T), N, C, D, int(bool(normalize)),
                                           float(softmax_scale if softmax_scale is not None else 1.0),
                                           _ptr(P), C, _ptr(S), lds, _ptr(ws), ws.numel(), _stream(dev))
        _lib.check(code, "mcd_gemm_nt_softmax_f32")
    if S is not None:
        return P, S[:, :C]
    return P


def last_gemm_path():
    """Which K1 kernel the last similarity_matrix call ran: 'tcgen05' (3xTF32 tensor-core GEMM + stand-alone softmax),
    'tcgen05_fused_softmax', or 'fp32_ffma' (the exact CUDA-core kernel: forced by tunable gemm_variant = 1, or shapes the
    tensor path does not take)."""
    return {1: "tcgen05", 2: "tcgen05_fused_softmax", 3: "fp32_ffma"}.get(int(_lib.lib().mcd_last_gemm_path()), "none")


def get_similarity_from_activations(target_save_name, clip_save_name, text_save_name, similarity_fn,
                                    return_target_feats=True, device="cuda", d_probe=None, top_k=None,
                                    target_on_device=False):
    """Drop-in for get_similarity_from_activations of all three reference variants: CLIP_og_utils.py:153-175
    (no extra arguments; target_feats handed back on the CPU), og_utils.py:478-521 (`d_probe`; target_feats loaded onto
    `device`) and utils.py:566-612 (`d_probe` and `top_k`, the latter forwarded to `similarity_fn` exactly as
    utils.py:602 does).  Loads the three cached .pt tensors, builds clip_feats on the GPU (K1) and calls
    `similarity_fn`.  `d_probe` only chose where the reference ran its matmul; it is accepted and not needed here."""
    image_features = torch.load(clip_save_name, map_location='cpu', weights_only=True)
    text_features = torch.load(text_save_name, map_location='cpu', weights_only=True)
    clip_feats = similarity_matrix(image_features, text_features, device=device)
    del image_features, text_features
    target_feats = torch.load(target_save_name, map_location=device if target_on_device else 'cpu', weights_only=True)
    if top_k is not None:
        similarity = similarity_fn(clip_feats, target_feats, device=device, top_k=top_k)
    else:
        similarity = similarity_fn(clip_feats, target_feats, device=device)
    del clip_feats
    if return_target_feats:
        return similarity, target_feats
    return similarity
