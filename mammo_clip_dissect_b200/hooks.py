"""The activation-summary forward hook with the reference's factory signature
(``get_activation(outputs, mode)`` -> ``hook(model, input, output)``; reference
concept_vit/utils.py:27-52 = og_utils.py:31-56 = CLIP_og_utils.py:13-36).

4-D (NCHW) activations are pooled by the sm_100a kernel K4 (include/mcd_b200.h: mcd_pool_nchw);
[B,T,D] activations yield the CLS token and [B,D] activations pass through, exactly as in the
reference (those two branches move no arithmetic).  As in the reference, ``avg`` unwraps tuple
outputs and ``max`` does not.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

_DTYPES = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}


_WS = {}   # grow-only partials buffer per (device, stream): the hook fires once per layer per batch


def _workspace(device, nbytes):
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf


def pool_nchw(x: torch.Tensor, mode: str, out: torch.Tensor = None) -> torch.Tensor:
    """[B,C,H,W] CUDA tensor -> [B,C] spatial mean ('avg') or max ('max').

    out=None: a new [B,C] tensor of x's dtype (what the reference's hook appends).  Otherwise `out` is a [B,C] view with
    unit column stride and any row stride, of x's dtype or fp32 -- e.g. rows of the stacked activation matrix --
    and the kernel writes straight into it (mcd_pool_nchw_to).  Contiguous and channels-last activations are pooled
    in place; other layouts are repacked once."""
    if x.dim() != 4:
        raise RuntimeError("pool_nchw expects a 4-D tensor")
    if not x.is_cuda:
        raise RuntimeError("mammo_clip_dissect_b200 has no CPU path: the hooked activation lives on %s" % x.device)
    if x.dtype not in _DTYPES:
        raise RuntimeError("pool_nchw supports float32/float16/bfloat16, got %s" % x.dtype)
    B, C, H, W = x.shape
    if B * C == 0 or H * W == 0:
        raise RuntimeError("pool_nchw: empty activation %s" % (tuple(x.shape),))
    x = x.detach()
    channels_last = 0
    if not x.is_contiguous():
        if x.is_contiguous(memory_format=torch.channels_last):
            channels_last = 1           # memory order [B, H, W, C]: pooled in place by the channels-last kernel
        else:
            x = x.contiguous()          # sliced / permuted activations: one repack, then the NCHW kernel
    lib = _lib.lib()
    if out is None:
        out = torch.empty((B, C), dtype=x.dtype, device=x.device)
    else:
        if tuple(out.shape) != (B, C) or out.device != x.device or out.dtype not in (x.dtype, torch.float32) or \
                (C > 1 and out.stride(1) != 1) or (B > 1 and out.stride(0) < C):
            raise RuntimeError("pool_nchw: out must be a [%d, %d] view on %s with unit column stride, dtype %s or float32"
                               % (B, C, x.device, x.dtype))
    out_ld = out.stride(0) if B > 1 else max(C, 1)
    need = int(lib.mcd_pool_nchw_workspace_bytes(B, C, H, W))
    with torch.cuda.device(x.device):
        ws = _workspace(x.device, need) if need else None      # only planes split across CTAs need partials
        code = lib.mcd_pool_nchw_to(ctypes.c_void_p(x.data_ptr()), _DTYPES[x.dtype], B, C, H, W, channels_last,
                                    _lib.POOL_MEAN if mode == "avg" else _lib.POOL_MAX, ctypes.c_void_p(out.data_ptr()),
                                    _DTYPES[out.dtype], out_ld, ctypes.c_void_p(ws.data_ptr() if ws is not None else 0),
                                    need, ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
    _lib.check(code, "mcd_pool_nchw_to")
    return out


def get_activation(outputs, mode):
    '''
    mode: how to pool activations: one of avg, max
    for fc or ViT neurons does no pooling
    '''
    if mode not in ("avg", "max"):
        raise ValueError("mode must be 'avg' or 'max', got %r" % (mode,))

    def hook(model, input, output):
        if mode == "avg" and type(output) is tuple:
            output = output[0]
        ndim = len(output.shape)
        if ndim == 4:      # CNN layers
            outputs.append(pool_nchw(output, mode))
        elif ndim == 3:    # ViT: CLS token
            outputs.append(output[:, 0].clone())
        elif ndim == 2:    # FC layers
            outputs.append(output.detach())
    return hook


class ActivationStack:
    """Pooled activations of all hooked layers, kept on the device in ONE [n_images, sum K_l] fp32 matrix
    (SURVEY.md section 8 f2): layer l owns columns [offset_l, offset_l + K_l), the hook of layer l writes the rows of
    the current batch.  Replaces the reference's per-layer Python lists + ``torch.cat`` + ``torch.save`` / ``torch.load``
    round trip (utils.py:151-200) when the activations are scored in the same process:

        stack = ActivationStack(n_images, [24, 40, 64], "cuda")
        for l, layer in enumerate(layers): layer.register_forward_hook(stack.hook(l, "avg"))
        for batch in loader: model(batch)                         # hooks fill the matrix batch by batch
        scores = similarity.soft_wpmi_layers(clip_feats, stack)   # no concatenation, no copy

    ``stack.layer(l)`` is the [n_images, K_l] view the reference would have saved as one .pt file."""

    def __init__(self, n_images, widths, device="cuda"):
        self.widths = [int(w) for w in widths]
        if n_images < 1 or not self.widths or min(self.widths) < 1:
            raise ValueError("ActivationStack needs n_images >= 1 and positive layer widths")
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("mammo_clip_dissect_b200 has no CPU path: device=%r" % (device,))
        self.offsets = [0]
        for w in self.widths:
            self.offsets.append(self.offsets[-1] + w)
        self.matrix = torch.zeros((int(n_images), self.offsets[-1]), dtype=torch.float32, device=dev)
        self.rows = [0] * len(self.widths)

    def reset(self):
        self.rows = [0] * len(self.widths)

    def layer(self, l):
        return self.matrix[:, self.offsets[l]:self.offsets[l + 1]]

    def complete(self):
        return all(r == self.matrix.shape[0] for r in self.rows)

    def hook(self, l, mode):
        """Forward hook for layer l with the summary rules of get_activation(outputs, mode)."""
        if mode not in ("avg", "max"):
            raise ValueError("mode must be 'avg' or 'max', got %r" % (mode,))
        lo, hi = self.offsets[l], self.offsets[l + 1]

        def hook(model, input, output):
            if mode == "avg" and type(output) is tuple:
                output = output[0]
            ndim = len(output.shape)
            if ndim not in (2, 3, 4):
                return
            B = output.shape[0]
            width = output.shape[1] if ndim != 3 else output.shape[2]
            r = self.rows[l]
            if width != hi - lo or r + B > self.matrix.shape[0]:
                raise RuntimeError("ActivationStack: layer %d produced [%d, %d], expected width %d and at most %d more rows"
                                   % (l, B, width, hi - lo, self.matrix.shape[0] - r))
            dst = self.matrix[r:r + B, lo:hi]
            if ndim == 4:
                pool_nchw(output, mode, out=dst)        # K4 writes the rows of the stacked matrix itself
            elif ndim == 3:
                dst.copy_(output.detach()[:, 0])        # ViT: CLS token
            else:
                dst.copy_(output.detach())              # FC layers
            self.rows[l] = r + B
        return hook
