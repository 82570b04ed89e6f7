"""GPU parity: the CUDA path (through the C ABI, via the host mirror of the reference surface)
against the CPU oracle and the golden fixtures produced by the reference itself.

Tolerances (SURVEY.md section 8a / DESIGN.md "Numerics"):
  * top-k indices: bit-exact under the stated total order;
  * pre-normalisation scores L: |L - L_ref| <= 1e-5 * |L_ref| element-wise;
  * final scores: |out - out_ref| <= 1e-5 * max|L|  (out is a difference of O(|L|) numbers);
  * per-neuron top concept: identical wherever the fp64 top-1/top-2 gap exceeds that noise.
"""
import numpy as np
import pytest
import torch

from oracle import similarity_oracle as orc

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def sim():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from mammo_clip_dissect_b200 import _lib, similarity
    assert _lib.lib().mcd_device_check() == 0, "not an sm_100 device"
    return similarity


def gen(seed):
    return torch.Generator().manual_seed(seed)


class topk_path:
    """K2 has two implementations: the radix select for short columns (N <= 16384, "auto"; "cluster1/2/8" force the
    number of CTAs that share a column group's rows) and the streaming scan ("stream": tunable topk_small = 1 switches
    the short-column kernel off)."""

    def __init__(self, path):
        self.path = path

    def __enter__(self):
        from mammo_clip_dissect_b200 import _lib
        _lib.set_tunable("topk_small", {"stream": 1, "cluster1": 2, "cluster2": 4, "cluster8": 16}.get(self.path, 0))

    def __exit__(self, *exc):
        from mammo_clip_dissect_b200 import _lib
        _lib.set_tunable("topk_small", 0)


def check_scores(out, L, ref_out, ref_L, f64_out=None, label=""):
    out, L = out.cpu(), L.cpu()
    scale = ref_L.abs().max().item()
    relL = ((L - ref_L).abs() / ref_L.abs().clamp_min(1e-30)).max().item()
    assert relL <= 1e-5, "%s: L rel err %.3g" % (label, relL)
    err = (out - ref_out).abs().max().item()
    assert err <= 1e-5 * scale, "%s: out abs err %.3g > %.3g" % (label, err, 1e-5 * scale)
    # top concept per neuron: mismatches are only acceptable inside the fp32 noise band
    mism = (out.argmax(1) != ref_out.argmax(1)).nonzero().flatten()
    truth = f64_out if f64_out is not None else ref_out.double()
    for j in mism.tolist():
        top2 = truth[j].topk(2).values
        assert (top2[0] - top2[1]).item() <= 2e-5 * scale, "%s: neuron %d top concept differs outside noise" % (label, j)
    return err, len(mism)


# ------------------------------------------------------------------------------------------------
# K2: column top-k
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,K,k", [(2000, 256, 100), (777, 33, 28), (300, 7, 1), (130, 33, 130), (5000, 40, 10),
                                   (1024, 64, 48), (1500, 129, 112), (3000, 68, 200), (4000, 36, 300),
                                   (2048, 12, 496), (64, 4, 64), (1024, 70, 512), (700, 33, 129), (900, 40, 257),
                                   (16384, 40, 100), (16385, 33, 100), (31, 5, 31), (10000, 768, 100)])
@pytest.mark.parametrize("path", ["auto", "stream", "cluster8"])
def test_topk_tie_free(sim, N, K, k, path):
    A = torch.randn(N, K, generator=gen(N + K + k))
    with topk_path(path):
        vals, idx = sim.topk_cols(A, k, device=DEV, want_values=True)
    rv, ri = orc.topk_cols(A, k)
    assert torch.equal(idx.cpu(), ri)
    assert torch.equal(vals.cpu(), rv)
    assert torch.equal(ri, torch.topk(A, k, dim=0)[1])         # tie-free: the rule coincides with torch.topk


@pytest.mark.parametrize("kind", ["round1", "relu", "const", "nan_inf", "signed_zero", "sorted_up", "sorted_down"])
@pytest.mark.parametrize("path", ["auto", "stream", "cluster1", "cluster2", "cluster8"])
def test_topk_ties_and_specials(sim, kind, path):
    N, K, k = 3000, 96, 100
    A = torch.randn(N, K, generator=gen(5))
    if kind == "round1":
        A = (A * 10).round() / 10
    elif kind == "relu":
        A = torch.relu(A - 1.5)                    # ~93 % exact zeros: the k-th value is a tie
    elif kind == "const":
        A = torch.full((N, K), 0.25)
    elif kind == "nan_inf":
        m = torch.rand(N, K, generator=gen(6))
        A[m < 0.01] = float("nan")
        A[(m >= 0.01) & (m < 0.02)] = float("inf")
        A[(m >= 0.02) & (m < 0.03)] = float("-inf")
        A[:, 3] = float("nan")
        A[:, 4] = float("-inf")
    elif kind == "signed_zero":
        A = torch.where(torch.rand(N, K, generator=gen(7)) < 0.5, torch.tensor(-0.0), torch.tensor(0.0))
        A[::7] = -1.0
    elif kind == "sorted_up":
        A = torch.sort(A, dim=0).values            # every element beats the running threshold
    elif kind == "sorted_down":
        A = torch.sort(A, dim=0, descending=True).values
    with topk_path(path):
        idx = sim.topk_cols(A, k, device=DEV)
    assert torch.equal(idx.cpu(), orc.topk_cols(A, k)[1]), kind


def test_topk_strided_unaligned_and_forced_splits(sim):
    from mammo_clip_dissect_b200 import _lib
    base = torch.randn(2500, 203, generator=gen(9)).to(DEV)
    A = base[:, 3:200]                             # row stride 203, 12-byte offset: element-copy producer path
    ref = orc.topk_cols(A.cpu(), 100)[1]
    assert torch.equal(sim.topk_cols(A, 100, device=DEV).cpu(), ref)
    B = torch.randn(6000, 260, generator=gen(10))
    refB = orc.topk_cols(B, 100)[1]
    try:
        for feed in (0, 1):                                   # 0: TMA tensor tiles, 1: element copies by the warp
            _lib.set_tunable("topk_variant", feed)
            for splits in (1, 2, 7, 15):
                _lib.set_tunable("topk_splits", splits)
                assert torch.equal(sim.topk_cols(B, 100, device=DEV).cpu(), refB), (feed, splits)
    finally:
        _lib.set_tunable("topk_splits", 0)
        _lib.set_tunable("topk_variant", 0)


@pytest.mark.parametrize("k", [1, 5, 31, 64, 100, 128])
def test_topk_finish_register_and_generic_sorters_agree(sim, k):
    """splits * k <= 128 takes the register-resident bitonic network; topk_variant = 2 forces the shared-memory one."""
    from mammo_clip_dissect_b200 import _lib
    A = torch.randn(1500, 77, generator=gen(12)).round(decimals=1)         # heavy ties: the index order matters
    A[5, 3] = float("nan")
    ref_v, ref_i = orc.topk_cols(A, k)
    got = sim.topk_cols(A, k, device=DEV, want_values=True)
    try:
        _lib.set_tunable("topk_variant", 2)
        gen_i = sim.topk_cols(A, k, device=DEV).cpu()
    finally:
        _lib.set_tunable("topk_variant", 0)
    assert torch.equal(got[1].cpu(), ref_i) and torch.equal(gen_i, ref_i)
    assert torch.equal(got[0].cpu().nan_to_num(7.0), ref_v.nan_to_num(7.0))


@pytest.mark.parametrize("kind", ["randn", "spikes_on_sampled_rows", "relu", "const", "nan_cols", "sorted_up"])
def test_topk_pre_threshold_and_exact_redo(sim, kind):
    """Long single-split scans start from a sampled threshold (k'-th largest of every 32nd row); column groups
    where fewer than k elements beat it are redone exactly.  Results must not depend on whether the scheme is on."""
    from mammo_clip_dissect_b200 import _lib
    N, K, k = 20000, 200, 100
    A = torch.randn(N, K, generator=gen(77))
    if kind == "spikes_on_sampled_rows":
        A[::32] += 50.0                      # the sample only sees the spikes: threshold far too high -> redo
    elif kind == "relu":
        A = torch.relu(A - 2.5)              # 99.4 % exact zeros
    elif kind == "const":
        A = torch.full((N, K), -3.0)
    elif kind == "nan_cols":
        A[:, ::3] = float("nan")
        A[::5, 1] = float("inf")
    elif kind == "sorted_up":
        A = torch.sort(A, dim=0).values
    ref = orc.topk_cols(A, k)[1]
    try:
        _lib.set_tunable("topk_splits", 1)
        n0 = _lib.launch_count()
        got = sim.topk_cols(A, k, device=DEV)
        assert _lib.launch_count() - n0 == 5            # sample tile maxima, select, scan, redo pass, finish
        assert torch.equal(got.cpu(), ref), kind
        _lib.set_tunable("topk_pre", 2)                  # scheme off
        n0 = _lib.launch_count()
        assert torch.equal(sim.topk_cols(A, k, device=DEV).cpu(), ref)
        assert _lib.launch_count() - n0 == 2
    finally:
        _lib.set_tunable("topk_splits", 0)
        _lib.set_tunable("topk_pre", 0)


def _filter_case(kind, N, K, seed=77):
    A = torch.randn(N, K, generator=gen(seed))
    tile = torch.arange(N) // 32
    if kind == "spikes_on_sampled_rows":
        A[::32] += 50.0                      # the sample only sees the spikes: threshold far too high -> short lists -> redo
    elif kind == "overflow":
        A[(tile % 8 == 5) | (tile % 8 == 7)] += 50.0        # a quarter of the rows, none in a sampled tile: lists overflow -> redo
    elif kind == "relu":
        A = torch.relu(A - 2.5)              # 99.4 % exact zeros
    elif kind == "const":
        A = torch.full((N, K), -3.0)
    elif kind == "nan_cols":
        A[:, ::3] = float("nan")
        A[::5, 1] = float("inf")
    elif kind == "sorted_up":
        A = torch.sort(A, dim=0).values      # every survivor of a column sits in the last rows: bags fill within a tile
    elif kind == "sorted_down":
        A = torch.sort(A, dim=0, descending=True).values
    elif kind == "round1":
        A = (A * 10).round() / 10            # ties inside the top k and at its boundary
    elif kind == "mixed":
        A[:, : K // 2] = torch.sort(A[:, : K // 2], dim=0).values
        A[:, 1::4] = 0.5
    return A


@pytest.mark.parametrize("kind", ["randn", "spikes_on_sampled_rows", "overflow", "relu", "const", "nan_cols", "sorted_up",
                                  "sorted_down", "round1", "mixed"])
@pytest.mark.parametrize("N,K,k", [(40000, 200, 100), (25017, 131, 28), (33333, 260, 10), (60000, 100, 256)])
def test_topk_filter_form(sim, kind, N, K, k):
    """Long TMA-aligned columns take the filter form (sample threshold -> filter scan -> select, exact redo of flagged
    column groups).  Same bits as the oracle and as the kept-set scan (tunable topk_filter = 1), values included."""
    from mammo_clip_dissect_b200 import _lib
    if K % 4:
        K += 4 - K % 4                       # TMA needs a 16-byte row pitch (other pitches take the element-copy scan)
    A = _filter_case(kind, N, K)
    ref_v, ref_i = orc.topk_cols(A, k)
    n0 = _lib.launch_count()
    vals, idx = sim.topk_cols(A, k, device=DEV, want_values=True)
    # sample tile maxima, sample select, filter scan (+ the last N % 8 rows), select, redo scan, redo finish
    assert _lib.launch_count() - n0 == 6 + (N % 8 != 0)
    assert torch.equal(idx.cpu(), ref_i), kind
    assert torch.equal(vals.cpu().nan_to_num(7.0), ref_v.nan_to_num(7.0))
    try:
        _lib.set_tunable("topk_filter", 1)
        n0 = _lib.launch_count()
        assert torch.equal(sim.topk_cols(A, k, device=DEV).cpu(), ref_i)
        assert _lib.launch_count() - n0 == 5
    finally:
        _lib.set_tunable("topk_filter", 0)


@pytest.mark.parametrize("kind", ["randn", "mixed"])
def test_topk_filter_form_beyond_128_sampled_tiles(sim, kind):
    """N = 140 000 rows give 136 sampled tiles: the thread-per-column threshold select (the warp-per-column one holds 128)."""
    A = _filter_case(kind, 140000, 36, seed=9)
    assert torch.equal(sim.topk_cols(A, 100, device=DEV).cpu(), orc.topk_cols(A, 100)[1])


@pytest.mark.parametrize("stages,chunk_tiles", [(2, 1), (3, 1000000), (4, 7)])
def test_topk_filter_ring_and_item_shapes(sim, stages, chunk_tiles):
    from mammo_clip_dissect_b200 import _lib
    A = _filter_case("mixed", 40011, 388, seed=5)
    ref = orc.topk_cols(A, 100)[1]
    try:
        _lib.set_tunable("filter_stages", stages)
        _lib.set_tunable("filter_chunk_tiles", chunk_tiles)
        assert torch.equal(sim.topk_cols(A, 100, device=DEV).cpu(), ref)
    finally:
        _lib.set_tunable("filter_stages", 0)
        _lib.set_tunable("filter_chunk_tiles", 0)


@pytest.mark.parametrize("chunks", [-1, 1, 2, 4, 8])
def test_pipelined_call_equals_the_staged_path(sim, chunks):
    """mcd_pmi_scores_f32 cuts the neurons into column chunks and runs chunk q's select + K3 + partials on a side stream
    under the scan of chunk q + 1: same bits as the staged single-stream path, for any chunk count."""
    from mammo_clip_dissect_b200 import _lib
    N, K, C = 40000, 2304, 763
    A = torch.randn(N, K, generator=gen(91)).to(DEV)
    A[:, 700] = 1.0                                       # a flagged column inside a chunk
    P = (torch.randn(N, C, generator=gen(92)) * 0.05).to(DEV)
    staged, raw, idx = sim.pmi_scores(P, A, 100, 10, 1, DEV, 1e-7, sim._reference_ramp(100, 0.998, 0.97).to(DEV), return_parts=True)
    try:
        _lib.set_tunable("pipe_chunks", chunks)
        for _ in range(2):
            got = sim.soft_wpmi(P, A, device=DEV)
            assert torch.equal(got, staged)
        L, part = sim.pmi_logsums(P, A, 100, 10, DEV, 1e-7, sim._reference_ramp(100, 0.998, 0.97).to(DEV))
        assert torch.equal(L, raw) and torch.equal(part, sim.lse_partials(raw))
        w = sim.wpmi(P, A, device=DEV)
    finally:
        _lib.set_tunable("pipe_chunks", 0)
    _lib.set_tunable("pipe_chunks", -1)
    try:
        assert torch.equal(w, sim.wpmi(P, A, device=DEV))
    finally:
        _lib.set_tunable("pipe_chunks", 0)



def test_topk_errors(sim):
    A = torch.randn(50, 4)
    with pytest.raises(RuntimeError):
        sim.topk_cols(A, 51, device=DEV)
    with pytest.raises(RuntimeError):
        sim.topk_cols(A, 0, device=DEV)
    with pytest.raises(RuntimeError):
        sim.soft_wpmi(torch.randn(50, 9), A, top_k=100, device=DEV)    # reference: torch.topk raises when k > N


# ------------------------------------------------------------------------------------------------
# K1b softmax, K3 accumulate, K3b log-sum-exp, and the two scoring functions end to end
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,C,a", [(257, 763, 10.0), (64, 37, 2.0), (33, 128, 10.0), (9, 1500, 4.0), (5, 2500, 1.0)])
def test_softmax_rows(sim, N, C, a):
    P = torch.randn(N, C, generator=gen(C)) * 0.2
    S = sim.concept_probabilities(P, a, device=DEV)
    ref = torch.softmax(a * P, dim=1)
    assert S.shape == (N, C)
    assert ((S.cpu() - ref).abs() <= 4e-7 * ref + 1e-12).all()
    assert S.stride(0) % 32 == 0 and float(S.cpu().sum(1).sub(1).abs().max()) < 1e-5


def test_golden_kat(sim, golden):
    g = golden("kat_8x3.npz")
    cf, tg = g["clip_feats"], g["target_feats"]
    assert torch.allclose(sim.soft_wpmi(cf, tg, top_k=4, device=DEV).cpu(), g["soft_wpmi_k4"], rtol=0, atol=5e-6)
    assert torch.allclose(sim.wpmi(cf, tg, top_k=4, device=DEV).cpu(), g["wpmi_k4"], rtol=0, atol=5e-6)
    assert torch.allclose(sim.cos_similarity_cubed(cf, tg, device=DEV).cpu(), g["cos_cubed"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(sim.cos_similarity(cf, tg, device=DEV).cpu(), g["cos"], rtol=1e-5, atol=1e-6)


def test_golden_c1_slice(sim, golden):
    """The reference's own outputs (fp32 and fp64 runs) on a 256 x 763 x 48 slice of config c1."""
    g = golden("c1_slice_256x763x48.npz")
    P, A = g["clip_feats"], g["target_feats"]
    out, L, idx = sim.pmi_scores(P, A, 100, 10, 1, DEV, 1e-7, sim._reference_ramp(100, 0.998, 0.97), return_parts=True)
    assert torch.equal(idx.cpu().long(), g["topk100"])
    _, refL, _ = orc.soft_wpmi(P, A, return_parts=True)
    err, mism = check_scores(out, L, g["soft_wpmi"], refL, g["soft_wpmi_f64"], "c1 slice soft_wpmi")
    ours_vs_truth = (out.cpu().double() - g["soft_wpmi_f64"]).abs().max().item()
    ref_vs_truth = (g["soft_wpmi"].double() - g["soft_wpmi_f64"]).abs().max().item()
    print("soft_wpmi vs fp64 truth: ours %.3g, reference fp32 %.3g; vs reference %.3g; argmax mismatches %d"
          % (ours_vs_truth, ref_vs_truth, err, mism))
    assert ours_vs_truth <= 3 * ref_vs_truth + 1e-4
    for key, kw in (("wpmi", {}), ("wpmi_k100", dict(top_k=100))):
        w = sim.wpmi(P, A, device=DEV, **kw).cpu()
        scale = orc.wpmi(P, A, return_parts=True, **kw)[1].abs().max().item()
        assert (w - g[key]).abs().max().item() <= 1e-5 * scale, key
    assert torch.allclose(sim.cos_similarity_cubed(P, A, device=DEV).cpu(), g["cos_cubed"], rtol=1e-4, atol=2e-6)
    assert torch.allclose(sim.cos_similarity(P, A, device=DEV).cpu(), g["cos"], rtol=1e-4, atol=2e-6)


@pytest.mark.parametrize("key,kw", [
    ("soft_wpmi_k10", dict(top_k=10)),
    ("soft_wpmi_k1", dict(top_k=1)),
    ("soft_wpmi_kN", dict(top_k=130)),
    ("soft_wpmi_params", dict(top_k=17, a=4, lam=0.5, min_prob=1e-6, p_start=0.9, p_end=0.6)),
])
def test_golden_odd_sizes(sim, golden, key, kw):
    g = golden("odd_130x37x33.npz")
    P, A = g["clip_feats"], g["target_feats"]
    scale = orc.soft_wpmi(P, A, return_parts=True, **kw)[1].abs().max().item()
    got = sim.soft_wpmi(P, A, device=DEV, **kw).cpu()
    assert (got - g[key]).abs().max().item() <= 1e-5 * max(scale, 1.0), key


def test_golden_odd_misc(sim, golden):
    g = golden("odd_130x37x33.npz")
    P, A = g["clip_feats"], g["target_feats"]
    k1 = sim.soft_wpmi(P, A[:, :1], top_k=10, device=DEV).cpu()          # a single neuron: log p(d) == its own score
    assert k1.shape == (1, 37) and (k1 - g["soft_wpmi_K1"]).abs().max().item() <= 1e-4
    assert (sim.wpmi(P, A, device=DEV).cpu() - g["wpmi_default"]).abs().max().item() <= 2e-3
    assert (sim.wpmi(P, A, top_k=5, a=7, lam=1.5, min_prob=1e-5, device=DEV).cpu() - g["wpmi_params"]).abs().max().item() <= 1e-3


@pytest.mark.parametrize("N,K,C", [(5000, 512, 763), (20011, 300, 763), (130, 37, 33), (4096, 129, 1), (33, 5, 200)])
@pytest.mark.parametrize("cubed", [False, True])
def test_cos_similarities_on_the_tensor_cores(sim, N, K, C, cubed):
    """cos / cos^3 run as a tcgen05 3xTF32 GEMM with the contraction over the probe images (accumulators flushed into
    fp32 registers every 8 k-blocks): fp32-grade against the reference's formula in fp64 for any N, and the same
    numbers (to fp32 noise) as the exact CUDA-core kernel."""
    from mammo_clip_dissect_b200 import _lib
    P = torch.randn(N, C, generator=gen(N + 1)) * 0.05 + 0.2
    A = torch.randn(N, K, generator=gen(N + 2)) * torch.rand(1, K, generator=gen(N + 3)) * 3 + 0.7
    fn = sim.cos_similarity_cubed if cubed else sim.cos_similarity
    got = fn(P, A, device=DEV).cpu()
    assert sim.last_cos_path() == "tcgen05"
    ref64 = (orc.cos_similarity_cubed if cubed else orc.cos_similarity)(P.double(), A.double())
    try:
        _lib.set_tunable("gemm_variant", 1)
        ffma = fn(P, A, device=DEV).cpu()
        assert sim.last_cos_path() == "fp32_ffma"
    finally:
        _lib.set_tunable("gemm_variant", 0)
    scale = ref64.abs().max().item()
    e_tc, e_ffma = (got.double() - ref64).abs().max().item(), (ffma.double() - ref64).abs().max().item()
    # the error floor is the fp32 evaluation of the column statistics and of f(x), shared by both kernels: the tensor-core
    # product must be no worse than true-fp32 arithmetic, and within 1e-5 of the largest similarity where that is wider
    assert e_tc <= max(1e-5 * scale, 1.5 * e_ffma + 1e-7), (e_tc, e_ffma, scale)


def test_cos_similarity_cubed_single_is_the_matched_pair_diagonal(sim):
    """Not in the reference tree (BASELINE's north star names it): the diagonal of cos_similarity_cubed for equally shaped
    inputs."""
    X = torch.randn(700, 41, generator=gen(71))
    Y = X * 0.5 + torch.randn(700, 41, generator=gen(72))
    d = sim.cos_similarity_cubed_single(X, Y, device=DEV).cpu()
    full = orc.cos_similarity_cubed(X.double(), Y.double())
    assert tuple(d.shape) == (41,) and (d.double() - torch.diagonal(full)).abs().max().item() <= 1e-5
    with pytest.raises(RuntimeError):
        sim.cos_similarity_cubed_single(X, Y[:, :40], device=DEV)


def test_config_c1_full(sim):
    """BASELINE config c1: clip_feats 2000 x 763, target_feats 2000 x 2048, top_k = 100."""
    I = torch.randn(2000, 512, generator=gen(0))
    T = torch.randn(763, 512, generator=gen(1))
    P = orc.similarity_matrix(I, T)
    A = torch.randn(2000, 2048, generator=gen(2))
    out, L, idx = sim.pmi_scores(P, A, 100, 10, 1, DEV, 1e-7, sim._reference_ramp(100, 0.998, 0.97), return_parts=True)
    ref, refL, ridx = orc.soft_wpmi_fast(P, A, return_parts=True)
    assert torch.equal(idx.cpu().long(), ridx)
    f64 = orc.soft_wpmi_fast(P, A, dtype=torch.float64)
    err, mism = check_scores(out, L, ref, refL, f64, "c1")
    print("c1: max |out - oracle| = %.3g, argmax mismatches within noise: %d / 2048" % (err, mism))
    # device-resident, non-contiguous and half-precision inputs go through the same path
    out2 = sim.soft_wpmi(P.to(DEV), A.to(DEV), device=DEV)
    assert torch.equal(out2, out)
    wide = torch.zeros(2000, 2100)
    wide[:, 10:2058] = A
    assert torch.equal(sim.soft_wpmi(P, wide[:, 10:2058], device=DEV), out)
    h = sim.soft_wpmi(P, A.half(), device=DEV).cpu()
    assert (h - orc.soft_wpmi_fast(P, A.half().float())).abs().max().item() <= 1e-5 * refL.abs().max().item()
    assert A.dtype == torch.float32 and torch.equal(A, torch.randn(2000, 2048, generator=gen(2)))   # inputs untouched


@pytest.mark.parametrize("tile", [0, 128, 192, 256, 384])
def test_accumulate_concept_tiles(sim, tile):
    from mammo_clip_dissect_b200 import _lib
    P = torch.randn(500, 763, generator=gen(3)) * 0.05
    A = torch.randn(500, 70, generator=gen(4))
    ref, refL, _ = orc.soft_wpmi_fast(P, A, top_k=50, return_parts=True)
    try:
        _lib.set_tunable("accum_tile", tile)
        out = sim.soft_wpmi(P, A, top_k=50, device=DEV).cpu()
    finally:
        _lib.set_tunable("accum_tile", 0)
    assert (out - ref).abs().max().item() <= 1e-5 * refL.abs().max().item()


@pytest.mark.parametrize("k", [1, 3, 4, 7, 8, 12, 50, 101])
@pytest.mark.parametrize("case", ["soft", "hard", "tiny_eps", "weights_above_one", "per_term", "reference_order"])
def test_accumulate_grouped_logs_and_fallbacks(sim, k, case):
    """K3 multiplies the terms of 4 ranks before one lg2 when eps >= 1e-9 and every weight is in [0,1] (default: each
    term as one FMA; accum_unroll = 2: in the reference's operation order); otherwise (and with accum_unroll = 1) it
    takes one lg2 per term in reference order.  All paths against the oracle, incl. NaN where the reference takes the
    log of a negative number."""
    from mammo_clip_dissect_b200 import _lib
    S = torch.softmax(10 * torch.randn(400, 131, generator=gen(31)) * 0.3, dim=1)
    idx = torch.stack([torch.randperm(400, generator=gen(32 + j))[:k] for j in range(45)], dim=1)   # [k, 45]
    eps = 1e-12 if case == "tiny_eps" else 1e-7
    w = None if case == "hard" else orc.p_ramp(k, 1.3 if case == "weights_above_one" else 0.998, 0.97)
    ref = orc.log_sums_chunked(S, idx, w, eps)
    try:
        _lib.set_tunable("accum_unroll", {"per_term": 1, "reference_order": 2}.get(case, 0))
        out = sim.log_sums(S.to(DEV), idx.to(DEV).int(), None if w is None else w.to(DEV), eps).cpu()
    finally:
        _lib.set_tunable("accum_unroll", 0)
    if case == "weights_above_one":
        assert ref.isnan().any() or k < 3                      # the case is meant to reach log(negative)
        assert torch.equal(out.isnan(), ref.isnan())
    fin = ~ref.isnan()
    assert ((out[fin] - ref[fin]).abs() <= 1e-5 * ref[fin].abs() + 1e-6).all()


def test_accumulate_plain_entry_keeps_the_reference_order_for_any_matrix(sim):
    """mcd_wpmi_accum_f32 (no promise about S) evaluates every term in the reference's order: entries outside [0, 1] --
    where the grouped evaluation could turn two negative terms into a positive product -- give the reference's NaN / -inf
    pattern and values."""
    S = torch.randn(300, 77, generator=gen(35)) * 0.6          # not a probability matrix: negative entries, entries > 1
    k = 8
    idx = torch.stack([torch.randperm(300, generator=gen(36 + j))[:k] for j in range(21)], dim=1)
    w = orc.p_ramp(k, 0.998, 0.97)
    for weights in (w, None):
        ref = orc.log_sums_chunked(S, idx, weights, 1e-7)
        assert ref.isnan().any()
        out = sim.log_sums(S.to(DEV), idx.to(DEV).int(), None if weights is None else weights.to(DEV), 1e-7,
                           probabilities=False).cpu()
        assert torch.equal(out.isnan(), ref.isnan())
        fin = torch.isfinite(ref)
        assert torch.equal(torch.isinf(out), torch.isinf(ref))
        assert ((out[fin] - ref[fin]).abs() <= 1e-5 * ref[fin].abs() + 1e-5).all()


def test_tied_activations_use_the_stated_rule(sim):
    """relu activations: scores must match the oracle run with the stated tie order spliced in."""
    P = torch.randn(800, 763, generator=gen(12)) * 0.05
    A = torch.relu(torch.randn(800, 64, generator=gen(13)) - 1.0)
    ref, refL, ridx = orc.soft_wpmi_fast(P, A, return_parts=True)
    out, L, idx = sim.pmi_scores(P, A, 100, 10, 1, DEV, 1e-7, sim._reference_ramp(100, 0.998, 0.97), return_parts=True)
    assert torch.equal(idx.cpu().long(), ridx)
    check_scores(out, L, ref, refL, None, "relu ties")


def test_lse_blocks_are_sharding_invariant(sim):
    """Concatenating per-shard partials in block order reproduces the unsharded result bit for bit."""
    L = (-400 - 60 * torch.rand(1280, 763, generator=gen(14))).to(DEV)
    full = sim.lse_partials(L)
    halves = torch.cat([sim.lse_partials(L[:512].contiguous()), sim.lse_partials(L[512:].contiguous())])
    assert torch.equal(full, halves)
    out_full, pd = sim.pmi_finalize(L.clone(), full, 1280, 1.0)
    out_a, _ = sim.pmi_finalize(L[:512].clone(), halves, 1280, 1.0)
    out_b, _ = sim.pmi_finalize(L[512:].clone(), halves, 1280, 1.0)
    assert torch.equal(out_full, torch.cat([out_a, out_b]))
    ref_pd = torch.logsumexp(L.double().cpu(), dim=0) - np.log(1280)
    assert (pd.cpu().double() - ref_pd).abs().max().item() < 1e-4
    assert torch.allclose(orc.lse_block_partials(L.cpu())[:, 0], full.cpu()[:, 0])


@pytest.mark.parametrize("K,C", [(512, 763), (300, 131), (256, 5)])
def test_finalize_broadcast_kernel_matches_finalize(sim, K, C):
    """mcd_pmi_finalize_bcast_f32 (the fused finalize + score all-gather) on one GPU: every destination matrix
    receives exactly mcd_pmi_finalize_f32's bits in rows [row_offset, row_offset + K), and nothing else is touched."""
    from mammo_clip_dissect_b200 import _lib
    L = (-400 - 60 * torch.rand(K, C, generator=gen(15))).to(DEV)
    other = (-400 - 60 * torch.rand(256, C, generator=gen(16))).to(DEV)         # a second rank's neurons
    parts = torch.cat([sim.lse_partials(other), sim.lse_partials(L)])
    K_total, row0 = 256 + K, 256
    want, pd = sim.pmi_finalize(L.clone(), parts, K_total, 0.7)
    dests = [torch.full((K_total, C), 7.0, device=DEV) for _ in range(3)]
    pd2 = sim.pmi_finalize_bcast(L, parts, K_total, 0.7, [d.data_ptr() for d in dests], row0)
    assert torch.equal(pd, pd2)
    for d in dests:
        assert torch.equal(d[row0:], want) and bool((d[:row0] == 7.0).all())
    # argument validation: slice outside the destination, too many destinations, misaligned slice
    lib = _lib.lib()
    import ctypes
    arr = (ctypes.c_void_p * 1)(dests[0].data_ptr())
    pdp, st = pd2.data_ptr(), torch.cuda.current_stream().cuda_stream
    assert lib.mcd_pmi_finalize_bcast_f32(L.data_ptr(), K, C, parts.data_ptr(), parts.shape[0], K_total, 0.7, pdp, arr, 1,
                                          row0 + 1, st) == -1
    many = (ctypes.c_void_p * 17)(*[dests[0].data_ptr()] * 17)
    assert lib.mcd_pmi_finalize_bcast_f32(L.data_ptr(), K, C, parts.data_ptr(), parts.shape[0], K_total, 0.7, pdp, many,
                                          17, row0, st) == -2
    if C % 4:
        odd = (ctypes.c_void_p * 1)(dests[0].data_ptr())
        assert lib.mcd_pmi_finalize_bcast_f32(L.data_ptr(), K - 1, C, parts.data_ptr(), parts.shape[0], K_total, 0.7, pdp,
                                              odd, 1, row0 + 1, st) == -2


@pytest.mark.parametrize("N,K,k", [(700, 130, 50), (20000, 70, 100), (300, 5, 300)])
def test_single_entry_call_equals_the_staged_path(sim, N, K, k):
    """soft_wpmi / wpmi normally run behind ONE C entry point (mcd_pmi_scores_f32); with per-stage profiling on, the
    Python wrapper launches the same kernels stage by stage.  Same bits; workspace errors are reported."""
    from mammo_clip_dissect_b200 import _lib
    P = torch.randn(N, 763, generator=gen(61)) * 0.05
    A = torch.randn(N, K, generator=gen(62))
    fast, fast_w = sim.soft_wpmi(P, A, top_k=k, device=DEV), sim.wpmi(P, A, top_k=min(k, 28), device=DEV)
    try:
        sim.PROFILE = []
        staged, staged_w = sim.soft_wpmi(P, A, top_k=k, device=DEV), sim.wpmi(P, A, top_k=min(k, 28), device=DEV)
        assert {"softmax_rows", "topk_cols", "wpmi_accum", "lse_finalize"} <= set(sim.profile_summary())
    finally:
        sim.PROFILE = None
    assert torch.equal(fast, staged) and torch.equal(fast_w, staged_w)
    lib = _lib.lib()
    need = lib.mcd_pmi_scores_workspace_bytes(N, K, 763, k)
    assert need > 0 and lib.mcd_pmi_scores_workspace_bytes(N, K, 763, N + 1) == 0
    Pd, Ad, out = P.to(DEV), A.to(DEV), torch.empty(K, 763, device=DEV)
    ws = torch.empty(need, dtype=torch.uint8, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    assert lib.mcd_pmi_scores_f32(Pd.data_ptr(), 763, Ad.data_ptr(), K, N, K, 763, k, 10.0, 1.0, None, 1e-7,
                                  out.data_ptr(), 763, ws.data_ptr(), need - 1, st) == -3
    assert lib.mcd_pmi_scores_f32(Pd.data_ptr(), 762, Ad.data_ptr(), K, N, K, 763, k, 10.0, 1.0, None, 1e-7,
                                  out.data_ptr(), 763, ws.data_ptr(), need, st) == -1


@pytest.mark.parametrize("N", [1500, 20000])
def test_layers_in_one_pass_equal_separate_calls(sim, N):
    """soft_wpmi_layers / wpmi_layers: every layer gets exactly the bits of its own call -- the LSE
    blocks restart at each layer's first neuron -- and the first layer also matches the oracle."""
    widths = [24, 40, 300, 513, 7, 256]
    P = torch.randn(N, 763, generator=gen(41)) * 0.05
    layers = [torch.randn(N, w, generator=gen(50 + i)) for i, w in enumerate(widths)]
    many = sim.soft_wpmi_layers(P, layers, top_k=50, device=DEV)
    assert [tuple(m.shape) for m in many] == [(w, 763) for w in widths]
    for t, m in zip(layers, many):
        assert torch.equal(m, sim.soft_wpmi(P, t, top_k=50, device=DEV))
    for t, m in zip(layers, sim.wpmi_layers(P, layers, device=DEV)):
        assert torch.equal(m, sim.wpmi(P, t, device=DEV))
    one = sim.soft_wpmi_layers(P, layers[2:3], top_k=50, device=DEV)
    assert len(one) == 1 and torch.equal(one[0], many[2])
    if N <= 2000:
        ref, refL, _ = orc.soft_wpmi_fast(P, layers[0], top_k=50, return_parts=True)
        assert (many[0].cpu() - ref).abs().max().item() <= 1e-5 * refL.abs().max().item()
    assert sim.soft_wpmi_layers(P, [], device=DEV) == []
    with pytest.raises(RuntimeError):
        sim.soft_wpmi_layers(P, [layers[0], layers[1][:-1]], device=DEV)


def test_activation_stack_feeds_the_layers_call(sim):
    """hooks.ActivationStack: the hooks of all layers fill one [N, sum K] matrix batch by batch; its
    layer views equal what get_activation's lists + torch.cat give, and scoring the stack equals scoring the list."""
    from mammo_clip_dissect_b200 import hooks
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 24, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(24, 40, 3, stride=2),
                              torch.nn.ReLU(), torch.nn.Conv2d(40, 9, 1)).to(DEV).eval()
    hooked, widths, n_img = [0, 2, 4], [24, 40, 9], 300
    stack = hooks.ActivationStack(n_img, widths, DEV)
    lists = [[] for _ in hooked]
    handles = []
    for l, i in enumerate(hooked):
        handles.append(net[i].register_forward_hook(stack.hook(l, "avg")))
        handles.append(net[i].register_forward_hook(hooks.get_activation(lists[l], "avg")))
    x = torch.randn(n_img, 3, 20, 24, generator=gen(44))
    with torch.no_grad():
        for b0 in range(0, n_img, 64):                              # ragged last batch
            net(x[b0:b0 + 64].to(DEV))
            assert not stack.complete() or b0 + 64 >= n_img
    for h in handles:
        h.remove()
    assert stack.complete()
    cats = [torch.cat(v) for v in lists]
    for l in range(3):
        assert torch.equal(stack.layer(l), cats[l])
    P = torch.randn(n_img, 763, generator=gen(45)) * 0.05
    a = sim.soft_wpmi_layers(P, stack, top_k=20, device=DEV)
    b = sim.soft_wpmi_layers(P, cats, top_k=20, device=DEV)
    assert all(torch.equal(u, v) for u, v in zip(a, b)) and [t.shape[0] for t in a] == widths
    stack.reset()
    with pytest.raises(RuntimeError):
        sim.soft_wpmi_layers(P, stack, top_k=20, device=DEV)       # not refilled yet


@pytest.mark.parametrize("K,C,t", [(300, 763, 10), (33, 763, 1), (70, 29, 29), (5, 1000, 64), (257, 200, 5)])
def test_top_concepts_per_neuron(sim, K, C, t):
    """mcd_row_topk_f32 = torch.topk(scores, t, dim=1) of the reference's callers, under the stated order."""
    X = torch.randn(K, C, generator=gen(71))
    v, i = sim.top_concepts(X.to(DEV), t)
    rv, ri = torch.topk(X, t, dim=1)
    assert torch.equal(i.cpu(), ri) and torch.equal(v.cpu(), rv)                  # tie-free: equals torch.topk
    Y = (X * 4).round() / 4                                                     # heavy ties + specials
    Y[0, :3] = float("nan")
    Y[1, 5] = float("inf")
    Y[2] = 0.5
    want = torch.sort(Y, dim=1, descending=True, stable=True)                   # NaN first (largest), ties by index
    v, i = sim.top_concepts(Y.to(DEV), t)
    assert torch.equal(i.cpu(), want.indices[:, :t])
    assert torch.equal(v.cpu().nan_to_num(9.0), want.values[:, :t].nan_to_num(9.0))
    wide = torch.zeros(K, C + 7)
    wide[:, 3:C + 3] = X
    assert torch.equal(sim.top_concepts(wide.to(DEV)[:, 3:C + 3], t)[1].cpu(), ri)      # strided rows
    with pytest.raises(RuntimeError):
        sim.top_concepts(X.to(DEV), C + 1)


# ------------------------------------------------------------------------------------------------
# K1 similarity matrix, K4 hook
# ------------------------------------------------------------------------------------------------
def test_finalize_pass_emits_the_top_concepts(sim):
    """soft_wpmi_top / soft_wpmi_layers(top_concepts=t): scores bit-identical to soft_wpmi, and the (value, concept)
    pairs equal the stand-alone row top-k of that matrix (stated order; NaN-free here), for one layer and for stacked
    layers."""
    N, C = 3000, 763
    P = torch.randn(N, C, generator=gen(81)) * 0.05
    layers = [torch.randn(N, w, generator=gen(82 + w)) for w in (24, 300, 513)]
    for t in (1, 10, 64):
        out, vals, idx = sim.soft_wpmi_top(P, layers[1], top_concepts=t, device=DEV)
        assert torch.equal(out, sim.soft_wpmi(P, layers[1], device=DEV))
        rv, ri = sim.top_concepts(out, t)
        assert torch.equal(vals, rv) and torch.equal(idx, ri)
        assert torch.equal(vals, out.gather(1, idx))
    w_out, w_vals, w_idx = sim.wpmi_top(P, layers[0], top_concepts=5, device=DEV)
    assert torch.equal(w_out, sim.wpmi(P, layers[0], device=DEV)) and torch.equal(w_idx, sim.top_concepts(w_out, 5)[1])
    scores, vals_l, idx_l = sim.soft_wpmi_layers(P, layers, device=DEV, top_concepts=10)
    plain = sim.soft_wpmi_layers(P, layers, device=DEV)
    for l in range(3):
        assert torch.equal(scores[l], plain[l])
        rv, ri = sim.top_concepts(scores[l], 10)
        assert torch.equal(vals_l[l], rv) and torch.equal(idx_l[l], ri)
    with pytest.raises(RuntimeError):
        sim.soft_wpmi_top(P, layers[0], top_concepts=65, device=DEV)


def test_similarity_matrix(sim, golden):
    from mammo_clip_dissect_b200 import features
    g = golden("itt_40x29.npz")
    I, T = g["image_features"], g["text_features"]
    keepI = I.clone()
    P = features.similarity_matrix(I, T, device=DEV)
    assert torch.allclose(P.cpu(), g["clip_feats"], rtol=0, atol=2e-6)
    assert torch.equal(I, keepI)                                   # the reference normalises in place; we must not
    I2 = torch.randn(1000, 512, generator=gen(20))
    T2 = torch.randn(763, 512, generator=gen(21))
    P2, S2 = features.similarity_matrix(I2, T2, device=DEV, softmax_scale=10)
    ref = orc.similarity_matrix(I2.double(), T2.double()) if False else (
        (I2.double() / I2.double().norm(dim=-1, keepdim=True)) @ (T2.double() / T2.double().norm(dim=-1, keepdim=True)).T)
    assert (P2.cpu().double() - ref).abs().max().item() <= 2e-6
    refS = torch.softmax(10 * ref, dim=1)
    assert ((S2.cpu().double() - refS).abs() / refS).max().item() <= 1e-5      # tcgen05 3xTF32 with 4 accumulators (fp32 path: 2e-6)


@pytest.mark.parametrize("N,C,D,normalize", [(1000, 763, 512, True), (130, 29, 512, True), (257, 300, 96, False),
                                               (64, 763, 40, True)])
def test_similarity_matrix_tensor_core_vs_fp32(sim, N, C, D, normalize):
    """K1: the tcgen05 3xTF32 path and the exact-fp32 CUDA-core kernel against fp64."""
    from mammo_clip_dissect_b200 import _lib, features
    I = torch.randn(N, D, generator=gen(N)) * 1.7
    T = torch.randn(C, D, generator=gen(C)) * 0.3
    Id, Td = I.double(), T.double()
    if normalize:
        Id, Td = Id / Id.norm(dim=-1, keepdim=True), Td / Td.norm(dim=-1, keepdim=True)
    ref = Id @ Td.T
    scale = ref.abs().max().item()
    n0 = _lib.launch_count()
    P_tc = features.similarity_matrix(I, T, device=DEV, normalize=normalize).cpu()
    assert _lib.launch_count() - n0 == 3                 # 2 x prepare_rows + the streaming GEMM
    assert features.last_gemm_path() == "tcgen05"        # ... and the library says which GEMM it ran (the CUDA-core
    try:                                                 # fallback is 3 launches too: 2 x row_norm + sgemm)
        _lib.set_tunable("gemm_variant", 1)
        P_32 = features.similarity_matrix(I, T, device=DEV, normalize=normalize).cpu()
        assert features.last_gemm_path() == "fp32_ffma"
    finally:
        _lib.set_tunable("gemm_variant", 0)
    assert not torch.equal(P_tc, P_32) or N * C < 64     # two different arithmetic paths: not the same bits
    e_tc = (P_tc.double() - ref).abs().max().item() / scale
    e_32 = (P_32.double() - ref).abs().max().item() / scale
    print("K1 max err / max|P|: tcgen05 3xTF32 %.3g, fp32 CUDA-core %.3g" % (e_tc, e_32))
    assert e_32 <= 2e-6 and e_tc <= 3e-6
    # and through the temperature softmax (a = 10 amplifies operand rounding)
    if normalize and C > 100:
        _, S = features.similarity_matrix(I, T, device=DEV, softmax_scale=10)
        refS = torch.softmax(10 * ref, dim=1)
        e_s = ((S.cpu().double() - refS).abs() / refS).max().item()
        print("   softmax(10 P) max rel err through the tensor-core path: %.3g" % e_s)
        assert e_s <= 1e-5

    def softmax_of(P32):
        """The kernels' softmax of THEIR OWN P: fp32-rounded a * P and (a * P - max) as they compute them, the rest in fp64."""
        z = P32 * 10.0
        d = z - z.max(dim=1, keepdim=True).values
        e = d.double().exp()
        return e / e.sum(dim=1, keepdim=True)

    # the tensor-core variants: streaming kernel + the stand-alone softmax (default), one CTA per tile + the stand-alone
    # softmax (2), the band kernel that rescales in place (3), the streaming kernel whose epilogue keeps the row softmax
    # pairs + a normalising pass (4)
    got = {}
    for variant, launches, path in ((0, 4, "tcgen05"), (2, 4, "tcgen05"), (3, 3, "tcgen05_fused_softmax"),
                                    (4, 4, "tcgen05_fused_softmax")):
        try:
            _lib.set_tunable("gemm_variant", variant)
            n0 = _lib.launch_count()
            P_v, S_v = features.similarity_matrix(I, T, device=DEV, normalize=normalize, softmax_scale=10)
            assert _lib.launch_count() - n0 == launches, variant
            assert features.last_gemm_path() == path, variant
        finally:
            _lib.set_tunable("gemm_variant", 0)
        P_v, S_v = P_v.cpu(), S_v.cpu()
        assert (P_v.double() - ref).abs().max().item() / scale <= 3e-6, variant
        want = softmax_of(P_v)
        big = want > 1e-30                                # below that fp32 exponentials are denormal or zero
        rel = ((S_v.double() - want).abs()[big] / want[big]).max().item()
        assert S_v[~big].max().item() <= 2e-30 if (~big).any() else True
        print("   variant %d: softmax against the fp64 softmax of its own P: max rel %.3g" % (variant, rel))
        assert rel <= 2e-6 and abs(S_v.sum(dim=1) - 1).max().item() < 1e-5, (variant, rel)
        got[variant] = P_v
    assert torch.equal(got[0], P_tc) and torch.equal(got[4], P_tc)       # with or without the softmax: the same P bits
    assert torch.equal(got[2], got[3])                   # the two one-tile-at-a-time kernels share their accumulation order
    for tiles in (1, 2, 4):                              # fewer column tiles per CTA: any split gives the same bits
        try:
            _lib.set_tunable("gemm_tiles_per_cta", tiles)
            assert torch.equal(features.similarity_matrix(I, T, device=DEV, normalize=normalize).cpu(), P_tc)
        finally:
            _lib.set_tunable("gemm_tiles_per_cta", 0)


def test_hook_matches_reference(sim, golden):
    from mammo_clip_dissect_b200.hooks import get_activation
    g = golden("hook_cases.npz")
    for mode, n in (("avg", 4), ("max", 3)):
        got = []
        hook = get_activation(got, mode)
        hook(None, None, g["x4"].to(DEV)); hook(None, None, g["x3"].to(DEV)); hook(None, None, g["x2"].to(DEV))
        if mode == "avg":
            hook(None, None, (g["x4"].to(DEV), "ignored"))
        assert len(got) == n
        for i, t in enumerate(got):
            want = g["%s_%d" % (mode, i)]
            assert t.shape == want.shape and t.is_cuda
            assert torch.allclose(t.cpu(), want, rtol=0, atol=1e-6 if mode == "avg" else 0)
    with pytest.raises(Exception):
        get_activation([], "max")(None, None, (g["x4"].to(DEV),))


@pytest.mark.parametrize("shape", [(4, 24, 760, 456), (3, 40, 380, 228), (5, 128, 95, 57), (7, 304, 48, 29),
                                   (2, 3, 1, 1), (1, 5, 7, 9), (2, 512, 33, 31)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
def test_pooling_shapes(sim, shape, dtype):
    from mammo_clip_dissect_b200.hooks import pool_nchw
    if dtype != torch.float32 and shape[2] > 400:
        pytest.skip("large planes: fp32 only")
    x = (torch.randn(*shape, generator=gen(sum(shape))) + 0.3).to(dtype)
    xd = x.to(DEV)
    mean = pool_nchw(xd, "avg").cpu()
    ref = x.double().mean(dim=[2, 3])
    tol = 1e-5 * x.double().abs().mean().item() if dtype == torch.float32 else 4e-3
    assert mean.dtype == dtype and (mean.double() - ref).abs().max().item() <= tol
    mx = pool_nchw(xd, "max").cpu()
    assert torch.equal(mx, x.amax(dim=[2, 3]))
    # views: channels_last and a sliced batch go through the same kernel after one repack
    assert torch.equal(pool_nchw(xd.to(memory_format=torch.channels_last), "max").cpu(), mx)
    if shape[0] > 1:
        assert torch.equal(pool_nchw(xd[1:], "max").cpu(), mx[1:])


def test_pooling_nan_propagates(sim):
    from mammo_clip_dissect_b200.hooks import pool_nchw
    x = torch.randn(2, 4, 50, 50, generator=gen(30))
    x[0, 1, 7, 9] = float("nan")
    got = pool_nchw(x.to(DEV), "max").cpu()
    assert torch.isnan(got[0, 1]) and torch.equal(torch.isnan(got), torch.isnan(x.amax(dim=[2, 3])))


@pytest.mark.parametrize("shape", [(4, 24, 190, 114), (3, 40, 95, 57), (5, 128, 48, 29), (2, 304, 12, 9), (1, 7, 33, 5),
                                   (6, 20, 3, 3)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("mode", ["avg", "max"])
def test_pooling_channels_last_and_strided_output(sim, shape, dtype, mode):
    """Channels-last activations are pooled in place (no repack), and the result may be written straight into rows of a
    wider fp32 matrix (hooks.ActivationStack): same values as the NCHW kernel / torch."""
    from mammo_clip_dissect_b200.hooks import pool_nchw
    B, C, H, W = shape
    x = (torch.randn(shape, generator=gen(B * C + H)) * 2 + 0.5).to(dtype).to(DEV)
    ref = x.float().mean(dim=[2, 3]) if mode == "avg" else x.float().amax(dim=[2, 3])
    tol = 1e-5 * x.float().abs().mean().item() if mode == "avg" else 0.0
    xl = x.contiguous(memory_format=torch.channels_last)
    assert not xl.is_contiguous() or C == 1 or H * W == 1
    big = torch.full((B + 2, C + 11), -7.0, dtype=torch.float32, device=DEV)
    for src in (x, xl):
        got = pool_nchw(src, mode)
        assert got.dtype == dtype and tuple(got.shape) == (B, C)
        lim = tol if dtype == torch.float32 else tol + 2 ** -7 * ref.abs().max().item()
        assert (got.float() - ref).abs().max().item() <= lim
        big.fill_(-7.0)
        view = big[1:1 + B, 5:5 + C]
        assert pool_nchw(src, mode, out=view) is view
        assert (view - ref).abs().max().item() <= tol                     # fp32 output: no rounding to the input dtype
        assert bool((big[0] == -7).all() and (big[-1] == -7).all() and (big[:, :5] == -7).all() and (big[:, 5 + C:] == -7).all())
    with pytest.raises(RuntimeError):
        pool_nchw(x, mode, out=big[:B, :C].t().t()[:, ::1].to(torch.float64))


# ------------------------------------------------------------------------------------------------
# size-independent properties at a larger shape (the oracle would take minutes here)
# ------------------------------------------------------------------------------------------------
def test_properties_at_scale(sim):
    N, K, C, k = 20000, 4096, 763, 100
    A = torch.randn(N, K, generator=gen(40)).to(DEV)
    P = (torch.randn(N, C, generator=gen(41)) * 0.05).to(DEV)
    vals, idx = sim.topk_cols(A, k, device=DEV, want_values=True)
    assert torch.equal(vals, A.gather(0, idx))                               # values are the selected elements
    assert bool((vals[:-1] >= vals[1:]).all())                               # sorted descending
    assert bool((idx.sort(dim=0).values[1:] != idx.sort(dim=0).values[:-1]).all())   # no index twice
    kth = vals[-1]
    assert int((A > kth).sum(0).max()) <= k - 1 and bool(((A >= kth).sum(0) >= k).all())   # exactly the top k
    assert torch.equal(idx[:5], torch.topk(A, 5, dim=0)[1])                  # caller-side topk(A, 5, 0) falls out
    out = sim.soft_wpmi(P, A, device=DEV)
    # permuting neurons permutes rows (log p(d) is order-free up to the fixed-block summation order)
    perm = torch.randperm(K, generator=gen(42)).to(DEV)
    out_p = sim.soft_wpmi(P, A[:, perm], device=DEV)
    assert (out_p - out[perm]).abs().max().item() <= 2e-3
    # lam = 0 switches the coupling off: rows equal the raw log-sums, and a sub-layer reproduces them exactly
    raw = sim.soft_wpmi(P, A, lam=0, device=DEV)
    sub = sim.soft_wpmi(P, A[:, 1000:1300], lam=0, device=DEV)
    assert torch.equal(sub, raw[1000:1300])
    # out = raw - lam * prob_d with a per-concept prob_d
    d = raw - out
    assert (d - d[0:1]).abs().max().item() <= 1e-3
    assert bool(torch.isfinite(out).all())


# ------------------------------------------------------------------------------------------------
# the benched configuration (c4: N = 100000 probe images, the stride-32 sample plan) against the oracle
# ------------------------------------------------------------------------------------------------
def test_c4_rows_against_the_oracle(sim):
    """N = 100000 x K = 4096 (an eighth of the bench's width, the same per-column work and the same K2 plan): the
    top-k indices of 288 sampled columns -- 32 of them with planted fp32 ties inside the top k and at its boundary --
    must equal the oracle's bit for bit, their soft-WPMI log-sums (lam = 0) must match the oracle within 1e-5 relative,
    and log p(d) over all 4096 neurons must match an fp64 logsumexp."""
    N, K, C, k = 100000, 4096, 763, 100
    A = torch.randn(N, K, generator=gen(2))
    P = torch.randn(N, C, generator=gen(0)) * 0.044
    cols = torch.randperm(K, generator=gen(3))[:288].sort().values
    tied = cols[:32]
    for c in tied.tolist():
        v, r = torch.topk(A[:, c], k)
        free = torch.tensor([i for i in range(0, 200) if i not in set(r.tolist())][:3])
        A[free[0], c] = v[49]                 # a tie inside the top k: order by image index decides ranks 50 / 51
        A[free[1], c] = v[k - 1]              # a tie at the boundary: the lower image index makes the cut
        A[free[2], c] = v[k - 1]
    sub = A[:, cols].contiguous()
    ref_i = orc.topk_cols(sub, k)[1]
    Ad, Pd = A.to(DEV), P.to(DEV)
    idx = sim.topk_cols(Ad, k, device=DEV)
    assert torch.equal(idx[:, cols.to(DEV)].cpu(), ref_i)
    raw = sim.soft_wpmi(Pd, Ad, lam=0, device=DEV)                 # lam = 0: the rows are the log-sums L
    ref_L = orc.soft_wpmi_fast(P, sub, top_k=k, lam=0, inds=ref_i)
    got = raw[cols.to(DEV)].cpu()
    rel = ((got - ref_L).abs() / ref_L.abs().clamp_min(1e-30)).max().item()
    assert rel <= 1e-5, rel
    assert bool((got.argmax(1) == ref_L.argmax(1)).all())
    # log p(d) over all neurons: block LSE (fp32 blocks, fp64 combine) against fp64 logsumexp of the same L
    out = sim.soft_wpmi(Pd, Ad, device=DEV)
    prob_d = (raw - out)[0]
    truth = (torch.logsumexp(raw.double(), dim=0) - torch.log(torch.tensor(float(K), dtype=torch.float64))).cpu()
    assert (prob_d.cpu().double() - truth).abs().max().item() <= 1e-5 * truth.abs().max().item()
    assert bool(((raw - out) - prob_d[None, :]).abs().max().item() <= 1e-3)


# ------------------------------------------------------------------------------------------------
# the drivers' call sequence (describe_clip_neurons.py:41-66, CLIP_og_utils.py:60-75,153-175) end to end
# ------------------------------------------------------------------------------------------------
def test_driver_call_sequence(sim, tmp_path):
    import mammo_clip_dissect_b200.similarity as similarity      # noqa: F401  (the name the driver's eval() uses)
    from mammo_clip_dissect_b200 import features
    from mammo_clip_dissect_b200.hooks import get_activation      # noqa: F401  (used inside the eval below)

    class Target(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.layer1 = torch.nn.Conv2d(3, 24, 3, padding=1)
            self.layer2 = torch.nn.Conv2d(24, 40, 3, stride=2)
            self.fc = torch.nn.Linear(40 * 7 * 7, 12)

        def forward(self, x):
            return self.fc(torch.relu(self.layer2(torch.relu(self.layer1(x)))).flatten(1))

    torch.manual_seed(0)
    target_model = Target().to(DEV).eval()
    images = torch.randn(300, 3, 16, 16)
    img_feats, txt_feats = torch.randn(300, 64) * 2, torch.randn(29, 64)
    # hooks are registered exactly like the reference does it: through eval() on a layer string
    all_features = {name: [] for name in ("layer1", "layer2", "fc")}
    hooks = {}
    for target_layer in all_features:
        command = "target_model.{}.register_forward_hook(get_activation(all_features[target_layer], 'avg'))".format(target_layer)
        hooks[target_layer] = eval(command)
    # stock hooks next to ours record the raw layer outputs, so the pooled values are checked on the very same
    # activations (the GPU convolutions themselves run in TF32 and differ from a CPU forward at the 1e-4 level)
    raw = {name: [] for name in all_features}
    raw_hooks = [getattr(target_model, name).register_forward_hook(
        lambda m, i, o, name=name: raw[name].append(o.detach().clone())) for name in all_features]
    with torch.no_grad():
        for i in range(0, 300, 100):
            target_model(images[i:i + 100].to(DEV))
    for h in raw_hooks:
        h.remove()
    torch.save(img_feats, tmp_path / "img.pt")
    torch.save(txt_feats, tmp_path / "txt.pt")
    acts = {}
    for name in all_features:
        out = torch.cat(raw[name]).cpu()
        acts[name] = out.double().mean(dim=[2, 3]).float() if out.dim() == 4 else out
    similarity_fn = eval("similarity.{}".format("soft_wpmi"))
    P_ref = orc.similarity_matrix(img_feats, txt_feats)
    for layer in all_features:
        pooled = torch.cat(all_features[layer])
        hooks[layer].remove()
        assert pooled.is_cuda and pooled.shape == acts[layer].shape
        assert torch.allclose(pooled.cpu(), acts[layer], rtol=0, atol=1e-6)
        torch.save(pooled.cpu(), tmp_path / ("%s.pt" % layer))
        sims, target_feats = features.get_similarity_from_activations(
            str(tmp_path / ("%s.pt" % layer)), str(tmp_path / "img.pt"), str(tmp_path / "txt.pt"), similarity_fn,
            return_target_feats=True, device=DEV)
        vals, ids = torch.max(sims, dim=1)
        _, top_ids = torch.topk(target_feats, k=5, dim=0)
        ref, refL, _ = orc.soft_wpmi_fast(P_ref, pooled.cpu(), return_parts=True)
        assert sims.shape == ref.shape and sims.is_cuda
        assert (sims.cpu() - ref).abs().max().item() <= 1e-5 * refL.abs().max().item()
        assert (ids.cpu() == ref.argmax(1)).float().mean().item() >= 0.99
        assert top_ids.shape == (5, pooled.shape[1])
    torch.manual_seed(5)
    rr = similarity.rank_reorder(P_ref, acts["fc"], device=DEV)
    assert rr.shape == (12, 29) and rr.is_cuda


# ------------------------------------------------------------------------------------------------
# rank_reorder (scope row f4): RNG stream replayed on the host, NaN where the mean cosine is negative
# ------------------------------------------------------------------------------------------------
def _same_with_nan(a, b, rtol):
    assert torch.equal(torch.isnan(a), torch.isnan(b))
    m = ~torch.isnan(a)
    return bool(((a[m] - b[m]).abs() <= rtol * b[m].abs() + 1e-7).all())


def test_rank_reorder_golden(sim, golden):
    g = golden("rank_reorder_200x21x6.npz")
    torch.manual_seed(int(g["seed"]))
    got = sim.rank_reorder(g["clip_feats"], g["target_feats"], device=DEV).cpu()
    assert _same_with_nan(got, g["rank_reorder"], 2e-5)


@pytest.mark.parametrize("N,C,K,kw", [(2000, 763, 40, {}), (5000, 100, 24, {}), (1000, 37, 9, dict(p=2, top_fraction=0.1, scale_p=1.0)),
                                       (3000, 64, 16, dict(p=2.5, scale_p=0.25)),
                                       (9000, 50, 12, {}),                          # top_n = 450: ranks in registers, 16 words per lane
                                       (12000, 41, 9, {}),                          # top_n = 600: shared-memory sort; K2 by radix select
                                       (100000, 33, 5, {}),                         # the c4 probe count: top_n = 5000
                                       (40, 7, 3, {})])                             # top_n = 2
@pytest.mark.parametrize("replay", ["device_mt", "raw_draws", "randperm_calls"])
def test_rank_reorder_vs_oracle(sim, N, C, K, kw, replay):
    """The reference's RNG stream is consumed by continuing MT19937 on the device, as raw generator outputs produced on
    the host (shuffles on the device in both cases), or call by call; all match the oracle under the same seed and leave
    the global generator in the oracle's state."""
    if replay != "device_mt" and N > 12000:
        pytest.skip("the host replays are fallbacks; covered at the smaller sizes")
    P = torch.randn(N, C, generator=gen(N)) * 0.05 + 0.04          # mixed-sign means: some concepts give NaN
    P[::3][: P[1::3].shape[0]] = P[1::3]                           # duplicated rows: equal cosines -> ties in the ranks
    A = torch.randn(N, K, generator=gen(K))
    saved, saved_dev = sim._replay_ok, dict(sim._device_replay_ok)
    try:
        assert sim._replay_works() and sim._device_replay_works(torch.device(DEV))
        if replay == "randperm_calls":
            sim._replay_ok = False
        elif replay == "raw_draws":
            for k_ in list(sim._device_replay_ok):
                sim._device_replay_ok[k_] = False
        torch.manual_seed(123)
        got = sim.rank_reorder(P, A, device=DEV, **kw).cpu()
        tail = torch.rand(4)
    finally:
        sim._replay_ok = saved
        sim._device_replay_ok.update(saved_dev)
    torch.manual_seed(123)
    ref = orc.rank_reorder(P, A, **kw)
    assert torch.equal(tail, torch.rand(4))                        # same generator state afterwards
    assert got.shape == ref.shape == (K, C)
    assert _same_with_nan(got, ref, 5e-5)


@pytest.mark.parametrize("N,K", [(5000, 12), (9000, 7)])
@pytest.mark.parametrize("kind", ["crowded", "constant", "two_values"])
def test_rank_reorder_cosines_that_agree_in_their_leading_bits(sim, N, K, kind):
    """The register sort orders by the leading 24 (23) bits of the cosine; elements that agree there are placed by an exact
    count.  Crowded columns (thousands of cosines within 4096 ulps), constant columns and two-valued columns must rank
    exactly like the oracle's stable argsort."""
    C = 19
    g = gen(N + K)
    if kind == "crowded":
        P = 0.25 + torch.randint(0, 4096, (N, C), generator=g).float() * 2.0 ** -25
        P[:, 1] = -P[:, 1]                                         # negative cosines: reversed key order, NaN output
    elif kind == "constant":
        P = torch.full((N, C), 0.125)
    else:
        P = torch.where(torch.rand(N, C, generator=g) < 0.5, torch.tensor(0.3), torch.tensor(0.3000001))
    A = torch.randn(N, K, generator=g)
    torch.manual_seed(5)
    got = sim.rank_reorder(P, A, device=DEV).cpu()
    torch.manual_seed(5)
    ref = orc.rank_reorder(P, A)
    assert _same_with_nan(got, ref, 5e-5)


def test_mt19937_on_the_device_continues_the_cpu_generator(sim):
    """mcd_mt19937_draws against numpy's MT19937 started from the same torch generator state: a million draws from a
    mid-block position, and a start exactly at a block boundary."""
    import numpy as np
    from mammo_clip_dissect_b200 import _lib
    lib = _lib.lib()
    for warm, count in ((7, 1_000_003), (0, 624), (312, 5)):
        g = torch.Generator().manual_seed(77 + warm)
        if warm:
            torch.randint(0, 10, (warm,), generator=g)
        key, pos = sim._mt_state_to_numpy(g.get_state())
        st = torch.from_numpy(np.append(key, np.uint32(pos)).view(np.int32)).to(DEV)
        draws = torch.empty((count,), dtype=torch.int32, device=DEV)
        assert lib.mcd_mt19937_draws(st.data_ptr(), count, draws.data_ptr(), None) == 0
        want = sim._raw_draws(count, g)
        key2, pos2 = sim._mt_state_to_numpy(g.get_state())
        assert np.array_equal(draws.cpu().numpy().view(np.uint32), want)
        got = st.cpu().numpy().view(np.uint32)
        assert np.array_equal(got[:624], key2) and int(got[624]) == int(pos2)


@pytest.mark.parametrize("K", [24, 176, 512])
@pytest.mark.parametrize("top_k", [10, 28, 50, 100, 200])
def test_c5_sweep_against_the_oracle(sim, K, top_k):
    """BASELINE configs[4]: the similarity functions x top_k 10 - 200 on EfficientNet-B5-shaped layers (N = 5000 probes,
    24 / 176 / 512 channels), each against the oracle: wpmi and soft_wpmi at every top_k (the cos functions and
    rank_reorder take no top_k and are swept over the layer widths)."""
    N, C = 5000, 763
    P = torch.randn(N, C, generator=gen(500 + K)) * 0.044
    A = torch.randn(N, K, generator=gen(600 + K))
    for name, ours, ref in (("wpmi", sim.wpmi, orc.wpmi_fast), ("soft_wpmi", sim.soft_wpmi, orc.soft_wpmi_fast)):
        out = ours(P, A, top_k=top_k, device=DEV).cpu()
        want, L, _ = ref(P, A, top_k=top_k, return_parts=True)
        err = (out - want).abs().max().item()
        assert err <= 1e-5 * L.abs().max().item(), (name, K, top_k, err)
        gap = want.topk(2, dim=1).values
        sure = (gap[:, 0] - gap[:, 1]) > 2e-5 * L.abs().max().item()
        assert bool((out.argmax(1) == want.argmax(1))[sure].all()), (name, K, top_k)
    if top_k == 10:
        for name, ours, ref in (("cos_similarity_cubed", sim.cos_similarity_cubed, orc.cos_similarity_cubed),
                                ("cos_similarity", sim.cos_similarity, orc.cos_similarity)):
            out = ours(P, A, device=DEV, top_k=top_k).cpu()            # top_k is accepted and ignored (utils.py:602 passes it)
            want = ref(P.double(), A.double())
            assert (out.double() - want).abs().max().item() <= 1e-5 * want.abs().max().item() + 2e-7, (name, K)
        Pp = P.abs() + 0.01                                             # positive cosines: no NaN rows
        torch.manual_seed(9)
        out = sim.rank_reorder(Pp, A, device=DEV, top_k=top_k).cpu()
        torch.manual_seed(9)
        assert _same_with_nan(out, orc.rank_reorder(Pp, A), 5e-5), ("rank_reorder", K)


def test_rank_reorder_limits(sim):
    with pytest.raises(NotImplementedError):
        sim.rank_reorder(torch.randn(170000, 8), torch.randn(170000, 2), device=DEV)    # top_n = 8500 > 8192


@pytest.mark.parametrize("N,K,k", [(12000, 20, 600), (100000, 8, 5000), (3000, 33, 1500), (20000, 5, 513)])
def test_topk_large_k_by_radix_select(sim, N, K, k):
    """k beyond the scan's kept sets (rank_reorder takes 5 % of the probe images): radix select over the column."""
    A = torch.randn(N, K, generator=gen(N + k)).round(decimals=2)           # plenty of ties
    vals, idx = sim.topk_cols(A, k, device=DEV, want_values=True)
    rv, ri = orc.topk_cols(A, k)
    assert torch.equal(idx.cpu(), ri) and torch.equal(vals.cpu(), rv)
