"""The oracle restatement against the reference's own outputs (tests/golden/*.npz,
produced by tests/golden/make_golden.py from /root/reference).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import similarity_oracle as orc


def same_bits(a, b):
    return a.shape == b.shape and a.dtype == b.dtype and torch.equal(
        torch.nan_to_num(a, nan=12345.0), torch.nan_to_num(b, nan=12345.0))


def test_kat(golden):
    g = golden("kat_8x3.npz")
    cf, tg = g["clip_feats"], g["target_feats"]
    assert torch.equal(orc.topk_cols(tg, 4)[1], g["topk4"])
    assert g["topk4"].T.tolist() == [[3, 2, 0, 5], [4, 3, 7, 6]]
    assert same_bits(orc.soft_wpmi(cf, tg, top_k=4), g["soft_wpmi_k4"])
    assert same_bits(orc.wpmi(cf, tg, top_k=4), g["wpmi_k4"])
    assert same_bits(orc.cos_similarity_cubed(cf, tg), g["cos_cubed"])
    assert same_bits(orc.cos_similarity(cf, tg), g["cos"])
    assert abs(float(g["soft_wpmi_k4"][0, 0]) - (-1.1462039947509766)) < 1e-6


def test_c1_slice_loop_form_is_bit_exact(golden):
    g = golden("c1_slice_256x763x48.npz")
    P, A = g["clip_feats"], g["target_feats"]
    assert torch.equal(orc.topk_cols(A, 100)[1], g["topk100"])        # tie-free: == torch.topk
    assert same_bits(orc.soft_wpmi(P, A), g["soft_wpmi"])
    assert same_bits(orc.wpmi(P, A), g["wpmi"])
    assert same_bits(orc.wpmi(P, A, top_k=100), g["wpmi_k100"])
    assert same_bits(orc.cos_similarity_cubed(P, A), g["cos_cubed"])
    assert same_bits(orc.cos_similarity(P, A), g["cos"])


def test_c1_slice_chunked_and_f64(golden):
    g = golden("c1_slice_256x763x48.npz")
    P, A = g["clip_feats"], g["target_feats"]
    out, L, _ = orc.soft_wpmi_fast(P, A, return_parts=True)
    # chunked form: same terms, possibly another summation order -> tolerance of SURVEY 8a
    assert (out - g["soft_wpmi"]).abs().max() <= 1e-5 * L.abs().max()
    f64 = orc.soft_wpmi_fast(P, A.double(), dtype=torch.float64)
    assert torch.allclose(f64, g["soft_wpmi_f64"], rtol=0, atol=1e-9)
    # the reference's own fp32 run sits ~1e-4 from its fp64 run; record the bound we rely on
    assert (g["soft_wpmi"].double() - g["soft_wpmi_f64"]).abs().max() < 1e-3


@pytest.mark.parametrize("key,kw", [
    ("soft_wpmi_k10", dict(top_k=10)),
    ("soft_wpmi_k1", dict(top_k=1)),
    ("soft_wpmi_kN", dict(top_k=130)),
    ("soft_wpmi_params", dict(top_k=17, a=4, lam=0.5, min_prob=1e-6, p_start=0.9, p_end=0.6)),
])
def test_odd_sizes_soft(golden, key, kw):
    g = golden("odd_130x37x33.npz")
    assert same_bits(orc.soft_wpmi(g["clip_feats"], g["target_feats"], **kw), g[key])


def test_odd_sizes_misc(golden):
    g = golden("odd_130x37x33.npz")
    P, A = g["clip_feats"], g["target_feats"]
    assert same_bits(orc.soft_wpmi(P, A[:, :1], top_k=10), g["soft_wpmi_K1"])
    assert same_bits(orc.wpmi(P, A), g["wpmi_default"])
    assert same_bits(orc.wpmi(P, A, top_k=5, a=7, lam=1.5, min_prob=1e-5), g["wpmi_params"])


def test_rank_reorder_rng_replay(golden):
    g = golden("rank_reorder_200x21x6.npz")
    torch.manual_seed(int(g["seed"]))
    got = orc.rank_reorder(g["clip_feats"], g["target_feats"])
    assert same_bits(got, g["rank_reorder"])


def test_hook(golden):
    g = golden("hook_cases.npz")
    for mode, n in (("avg", 4), ("max", 3)):
        got = []
        hook = orc.get_activation(got, mode)
        hook(None, None, g["x4"]); hook(None, None, g["x3"]); hook(None, None, g["x2"])
        if mode == "avg":
            hook(None, None, (g["x4"], "ignored"))
        assert len(got) == n
        for i, t in enumerate(got):
            assert same_bits(t, g["%s_%d" % (mode, i)])
    with pytest.raises(Exception):
        orc.get_activation([], "max")(None, None, (g["x4"],))          # no tuple unwrap in max mode


def test_similarity_matrix(golden):
    g = golden("itt_40x29.npz")
    assert same_bits(orc.similarity_matrix(g["image_features"], g["text_features"]), g["clip_feats"])


def test_tie_rule():
    col = torch.tensor([[1.0], [float("nan")], [-0.0], [0.0], [1.0], [float("inf")], [float("nan")], [-1.0]])
    vals, idx = orc.topk_cols(col, 7)
    assert idx[:, 0].tolist() == [1, 6, 5, 0, 4, 2, 3]     # NaN first (by index), inf, ties by index, -0 == +0
    with pytest.raises(RuntimeError):
        orc.topk_cols(col, 9)


def test_block_lse_matches_logsumexp():
    g = torch.Generator().manual_seed(11)
    L = -400 - 50 * torch.rand(700, 19, generator=g)
    lse = orc.lse_combine(orc.lse_block_partials(L))
    assert torch.allclose(lse, torch.logsumexp(L.double(), dim=0), rtol=0, atol=1e-4)


def test_rank_weights_have_the_reference_bits():
    """The ramp p_r of soft_wpmi (reference similarity.py:58): the values probed from the reference expression
     and equality of the oracle's and the product's host-side construction."""
    import torch
    from mammo_clip_dissect_b200.similarity import _reference_ramp
    from oracle import similarity_oracle as orc
    p = _reference_ramp(100, 0.998, 0.97)
    assert p.dtype == torch.float32 and p.shape == (100,)
    probed = {0: "0x1.fef9dcp-1", 1: "0x1.fed528p-1", 50: "0x1.f7cedap-1", 99: "0x1.f0c88cp-1"}
    for r, h in probed.items():
        assert float(p[r]) == float.fromhex(h), (r, float(p[r]).hex(), h)
    assert torch.equal(p, orc.p_ramp(100, 0.998, 0.97).to(torch.float32).reshape(-1))
    ref = 0.998 - torch.arange(start=0, end=100) / 100 * (0.998 - 0.97)         # the reference's expression, verbatim
    assert torch.equal(p, ref.to(torch.float32))
    for k, ps, pe in ((28, 0.9, 0.5), (1, 0.998, 0.97), (7, 1.0, 0.0)):
        assert torch.equal(_reference_ramp(k, ps, pe), orc.p_ramp(k, ps, pe).to(torch.float32).reshape(-1))
