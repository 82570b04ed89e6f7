"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol that
include/mcd_b200.h declares, and the host mirror refuses to run without CUDA (no fallback).
No kernel is launched here."""
import ctypes
import inspect
import os
import subprocess
import sys

import pytest
import torch

from mammo_clip_dissect_b200 import _lib, hooks, similarity


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.lib()


def test_header_and_bindings_agree():
    assert _lib.declared_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol(lib):
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in _lib.declared_symbols():
        assert hasattr(raw, name), name
    assert lib.mcd_abi_version() == 1
    assert b"sm_100a" in lib.mcd_build_info()
    assert lib.mcd_strerror(0) == b"ok" and lib.mcd_strerror(-3) == b"workspace too small"


def test_cubin_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = {ln.split(".")[-2] for ln in out.stdout.splitlines() if ".cubin" in ln}
    assert archs == {"sm_100a"}, archs


def test_workspace_queries_are_pure_host_functions(lib):
    # headline shape: 100k probes x 32768 neurons, k = 100
    ws = lib.mcd_topk_cols_workspace_bytes(100_000, 32_768, 100)
    assert ws >= 100 * 32_768 * 8
    assert lib.mcd_topk_cols_workspace_bytes(50, 8, 100) == 0          # k > N
    # the filter form (survivor lists: ~2000 words per column at c4, a few hundred for short columns) is planned wherever a
    # row sample applies, not for short or wide-k problems
    assert lib.mcd_topk_cols_workspace_bytes(100_000, 32_768, 100) >= 32_768 * 2000 * 8
    assert lib.mcd_topk_cols_workspace_bytes(10_000, 9_216, 100) >= 9_216 * 500 * 8
    assert lib.mcd_topk_cols_workspace_bytes(5_000, 512, 100) < 512 * 500 * 8
    assert lib.mcd_topk_cols_workspace_bytes(40_000, 200, 300) < 200 * 1000 * 8
    assert lib.mcd_topk_cols_workspace_bytes(10_000, 8, 1000) == 1000 * 8 * 8      # k > 512: radix select, candidates only
    assert lib.mcd_topk_cols_workspace_bytes(100_000, 8, 20_000) == 0      # beyond the radix select (k <= 16384)
    assert lib.mcd_pool_nchw_workspace_bytes(4, 24, 760, 456) > 0      # large planes are split
    assert lib.mcd_pool_nchw_workspace_bytes(64, 512, 12, 9) == 0      # small planes, many of them: no partials in either memory order
    assert lib.mcd_gemm_nt_softmax_workspace_bytes(2000, 763, 512) >= (2000 + 763) * 4
    # rank_reorder: baseline partials + the [K, C] denominators + the shuffles' working arrays
    assert lib.mcd_rank_reorder_workspace_bytes(512, 250, 763) >= 512 * 5 * 4 + 512 * 763 * 4 + 512 * 5 * 250 * 4
    assert lib.mcd_rank_reorder_workspace_bytes(0, 250, 763) == 0


def test_argument_validation_happens_before_any_launch(lib):
    n0 = lib.mcd_launch_count()
    assert lib.mcd_softmax_rows_f32(None, 4, None, 4, 1, 4, 1.0, None) == -1
    assert lib.mcd_topk_cols_f32(None, 4, 10, 4, 2, None, None, None, None, 0, None) == -1
    assert lib.mcd_wpmi_accum_f32(None, 4, 4, 4, None, 1, 1, None, 1e-7, None, 4, None) == -1
    assert lib.mcd_pool_nchw(None, 0, 1, 1, 1, 1, 0, None, None, 0, None) == -1
    assert lib.mcd_set_tunable(b"no_such_knob", 1) == -1
    assert lib.mcd_mt19937_draws(None, 10, None, None) == -1
    assert lib.mcd_bcast_f32(None, 4, None, 1, 0, None) == -1
    assert lib.mcd_rank_errors_f32(None, 4, 4, 4, None, None, 1, 1, 3.0, 0.5, None, 0, None, 4, None) == -1
    assert lib.mcd_rank_finish_f32(1, 4, 1, None, 0, None, 4, None) == -1
    assert lib.mcd_launch_count() == n0


def test_reference_call_surface():
    """Names, positional order and defaults of reference concept_vit/similarity.py."""
    def params(fn):
        return [(p.name, p.default) for p in inspect.signature(fn).parameters.values()]
    E = inspect.Parameter.empty
    assert params(similarity.soft_wpmi) == [("clip_feats", E), ("target_feats", E), ("top_k", 100), ("a", 10),
                                            ("lam", 1), ("device", "cuda"), ("min_prob", 1e-7),
                                            ("p_start", 0.998), ("p_end", 0.97)]
    assert params(similarity.wpmi) == [("clip_feats", E), ("target_feats", E), ("top_k", 28), ("a", 2),
                                       ("lam", 0.6), ("device", "cuda"), ("min_prob", 1e-7)]
    assert params(similarity.cos_similarity_cubed)[:5] == [("clip_feats", E), ("target_feats", E), ("device", "cuda"),
                                                           ("batch_size", 10000), ("min_norm", 1e-3)]
    assert params(similarity.cos_similarity)[:3] == [("clip_feats", E), ("target_feats", E), ("device", "cuda")]
    assert params(hooks.get_activation) == [("outputs", E), ("mode", E)]


def test_no_cpu_fallback():
    P, A = torch.zeros(8, 3), torch.zeros(8, 2)
    for fn in (similarity.soft_wpmi, similarity.wpmi, similarity.cos_similarity, similarity.cos_similarity_cubed):
        with pytest.raises(RuntimeError, match="no CPU path"):
            fn(P, A, device="cpu")
    got = []
    hook = hooks.get_activation(got, "avg")
    with pytest.raises(RuntimeError, match="no CPU path"):
        hook(None, None, torch.zeros(2, 3, 4, 5))
    hook(None, None, torch.ones(2, 5, 4))          # ViT / FC branches move no arithmetic
    hook(None, None, (torch.ones(2, 7), "x"))
    assert got[0].shape == (2, 4) and got[1].shape == (2, 7)
    with pytest.raises(ValueError):
        hooks.get_activation([], "median")


def test_product_does_not_import_the_oracle():
    pkg = os.path.dirname(_lib.__file__)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
    code = "import sys; import mammo_clip_dissect_b200.similarity, mammo_clip_dissect_b200.hooks, " \
           "mammo_clip_dissect_b200.features; sys.exit(any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules))"
    assert subprocess.run([sys.executable, "-c", code], cwd=os.path.dirname(pkg)).returncode == 0


def test_segment_tables_restart_the_lse_blocks_at_every_layer():
    """Host logic of soft_wpmi_layers: block / segment tables of the stacked call (no GPU needed)."""
    import math
    import torch
    from mammo_clip_dissect_b200 import similarity as sim
    blocks, segs, logs, n, row_seg = sim._segment_tables([24, 300, 513, 256], torch.device("cpu"))
    assert row_seg.dtype == torch.int32 and row_seg.tolist() == [0] * 24 + [1] * 300 + [2] * 513 + [3] * 256
    assert n == 7 and blocks.dtype == torch.int32 and logs.dtype == torch.float64
    assert blocks.tolist() == [[0, 24, 0], [24, 256, 1], [280, 44, 1], [324, 256, 2], [580, 256, 2], [836, 1, 2],
                               [837, 256, 3]]
    assert segs.tolist() == [[0, 1], [1, 2], [3, 3], [6, 1]]
    assert logs.tolist() == [math.log(24.0), math.log(300.0), math.log(513.0), math.log(256.0)]
    # rows are covered exactly once, in order
    assert sum(b[1] for b in blocks.tolist()) == 24 + 300 + 513 + 256


def test_activation_stack_and_exchange_refuse_cpu():
    import pytest
    from mammo_clip_dissect_b200 import hooks
    with pytest.raises(RuntimeError):
        hooks.ActivationStack(10, [4, 5], "cpu")
    with pytest.raises(ValueError):
        hooks.ActivationStack(0, [4], "cuda")
