"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

Run from the repo root:   python tests/golden/make_golden.py
Needs /root/reference (present only in the authoring container, never on the GPU box).
The reference ships no golden vectors of its own, so these files
are "reference source x installed torch" outputs; torch.__version__ is stored in each.

Reference entry points executed here:
  concept_vit/similarity.py      soft_wpmi, wpmi, cos_similarity, cos_similarity_cubed,
                                 rank_reorder                         (imported as-is)
  concept_vit/CLIP_og_utils.py   get_activation, get_similarity_from_activations
                                 (imported with empty stub modules for `clip` and
                                 `data_utils`, which need network/extra packages)
"""
import contextlib
import io
import os
import sys
import tempfile
import types

import numpy as np
import torch

REF = "/root/reference/concept_vit"
OUT = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    sys.path.insert(0, REF)
    for name in ("clip", "data_utils"):
        sys.modules.setdefault(name, types.ModuleType(name))
    import similarity  # noqa: E402
    import CLIP_og_utils  # noqa: E402
    sys.path.pop(0)
    return similarity, CLIP_og_utils


def _quiet(fn, *a, **kw):
    """The reference prints shapes and tqdm bars; keep the generator's output clean."""
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        return fn(*a, **kw)


def _rand(gen, *shape):
    return torch.rand(*shape, generator=gen)


def _randn(gen, *shape):
    return torch.randn(*shape, generator=gen)


def main():
    sim, og = _import_reference()
    meta = dict(torch_version=np.array(torch.__version__))

    # ---- 1. the survey's tiny known-answer case -------------------
    g = torch.Generator().manual_seed(1234)
    cf = _rand(g, 8, 3) * 2 - 1
    tg = _randn(g, 8, 2)
    np.savez(os.path.join(OUT, "kat_8x3.npz"), clip_feats=cf.numpy(), target_feats=tg.numpy(),
             topk4=torch.topk(tg, dim=0, k=4)[1].numpy(),
             soft_wpmi_k4=_quiet(sim.soft_wpmi, cf, tg, top_k=4, device="cpu").numpy(),
             wpmi_k4=_quiet(sim.wpmi, cf, tg, top_k=4, device="cpu").numpy(),
             cos_cubed=_quiet(sim.cos_similarity_cubed, cf, tg, device="cpu").numpy(),
             cos=_quiet(sim.cos_similarity, cf, tg, device="cpu").numpy(), **meta)

    # ---- 2. 763-concept slice of config c1 (tie-free randn activations) ----------------
    g = torch.Generator().manual_seed(0)
    img = _randn(g, 256, 512)
    g = torch.Generator().manual_seed(1)
    txt = _randn(g, 763, 512)
    img = img / img.norm(dim=-1, keepdim=True)
    txt = txt / txt.norm(dim=-1, keepdim=True)
    P = img @ txt.T
    g = torch.Generator().manual_seed(2)
    A = _randn(g, 256, 48)
    np.savez(os.path.join(OUT, "c1_slice_256x763x48.npz"), clip_feats=P.numpy(), target_feats=A.numpy(),
             topk100=torch.topk(A, dim=0, k=100)[1].numpy(),
             soft_wpmi=_quiet(sim.soft_wpmi, P, A, device="cpu").numpy(),
             soft_wpmi_f64=_quiet(sim.soft_wpmi, P.double(), A.double(), device="cpu").numpy(),
             wpmi=_quiet(sim.wpmi, P, A, device="cpu").numpy(),
             wpmi_k100=_quiet(sim.wpmi, P, A, top_k=100, device="cpu").numpy(),
             cos_cubed=_quiet(sim.cos_similarity_cubed, P, A, device="cpu").numpy(),
             cos=_quiet(sim.cos_similarity, P, A, device="cpu").numpy(), **meta)

    # ---- 3. odd sizes: C not a multiple of 4, K=1 and K=33, k=1 / k=N, non-default scalars
    g = torch.Generator().manual_seed(3)
    P2 = (0.22 + 0.03 * _randn(g, 130, 37)).clamp(-1, 1)       # "CLIP-like" sharp cosines
    A2 = _randn(g, 130, 33)
    np.savez(os.path.join(OUT, "odd_130x37x33.npz"), clip_feats=P2.numpy(), target_feats=A2.numpy(),
             soft_wpmi_k10=_quiet(sim.soft_wpmi, P2, A2, top_k=10, device="cpu").numpy(),
             soft_wpmi_k1=_quiet(sim.soft_wpmi, P2, A2, top_k=1, device="cpu").numpy(),
             soft_wpmi_kN=_quiet(sim.soft_wpmi, P2, A2, top_k=130, device="cpu").numpy(),
             soft_wpmi_params=_quiet(sim.soft_wpmi, P2, A2, top_k=17, a=4, lam=0.5, device="cpu",
                                     min_prob=1e-6, p_start=0.9, p_end=0.6).numpy(),
             soft_wpmi_K1=_quiet(sim.soft_wpmi, P2, A2[:, :1], top_k=10, device="cpu").numpy(),
             wpmi_default=_quiet(sim.wpmi, P2, A2, device="cpu").numpy(),
             wpmi_params=_quiet(sim.wpmi, P2, A2, top_k=5, a=7, lam=1.5, device="cpu", min_prob=1e-5).numpy(),
             **meta)

    # ---- 4. rank_reorder with the global CPU RNG seeded (reference draws 5 randperm/neuron)
    g = torch.Generator().manual_seed(5)
    P3 = _rand(g, 200, 21) * 0.5 + 0.05          # positive cosines -> no NaN from avg**0.5
    A3 = _randn(g, 200, 6)
    torch.manual_seed(77)
    rr = _quiet(sim.rank_reorder, P3, A3, device="cpu")
    np.savez(os.path.join(OUT, "rank_reorder_200x21x6.npz"), clip_feats=P3.numpy(), target_feats=A3.numpy(),
             seed=np.array(77), rank_reorder=rr.numpy(), **meta)

    # ---- 5. pooling hook (CLIP_og_utils.get_activation) --------------------------------
    g = torch.Generator().manual_seed(6)
    x4 = _randn(g, 3, 5, 7, 9)
    x3 = _randn(g, 3, 4, 6)
    x2 = _randn(g, 3, 8)
    res = {}
    for mode in ("avg", "max"):
        got = []
        hook = og.get_activation(got, mode)
        hook(None, None, x4)
        hook(None, None, x3)
        hook(None, None, x2)
        if mode == "avg":
            hook(None, None, (x4, "ignored"))
        for i, t in enumerate(got):
            res["%s_%d" % (mode, i)] = t.numpy()
    np.savez(os.path.join(OUT, "hook_cases.npz"), x4=x4.numpy(), x3=x3.numpy(), x2=x2.numpy(), **res, **meta)

    # ---- 6. image x text similarity matrix through get_similarity_from_activations ------
    g = torch.Generator().manual_seed(8)
    I = _randn(g, 40, 512) * 3.0
    T = _randn(g, 29, 512) * 0.5
    At = _randn(g, 40, 5)
    captured = {}

    def spy(clip_feats, target_feats, device="cpu"):
        captured["clip_feats"] = clip_feats.clone()
        return torch.zeros(target_feats.shape[1], clip_feats.shape[1])

    with tempfile.TemporaryDirectory() as d:
        torch.save(I, os.path.join(d, "i.pt"))
        torch.save(T, os.path.join(d, "t.pt"))
        torch.save(At, os.path.join(d, "a.pt"))
        og.get_similarity_from_activations(os.path.join(d, "a.pt"), os.path.join(d, "i.pt"),
                                           os.path.join(d, "t.pt"), spy, device="cpu")
    np.savez(os.path.join(OUT, "itt_40x29.npz"), image_features=I.numpy(), text_features=T.numpy(),
             clip_feats=captured["clip_feats"].numpy(), **meta)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
