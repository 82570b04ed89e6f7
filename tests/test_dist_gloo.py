"""world_size-2 (and 3) CPU tests of the neuron-sharded path over gloo.  The exchange logic of
mammo_clip_dissect_b200.distributed is driven with an oracle-backed compute backend (test-only:
the product backend is CUDA); the sharded result must equal the unsharded one bit for bit."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mammo_clip_dissect_b200 import distributed as mdist
from oracle import similarity_oracle as orc


class OracleBackend:
    def log_sums(self, clip_feats, target_shard, top_k, a, min_prob, ramp):
        probs = torch.softmax(a * clip_feats, dim=1)
        inds = orc.topk_cols(target_shard, int(top_k))[1]
        w = ramp.reshape(-1, 1) if ramp is not None else None
        return orc.log_sums_chunked(probs, inds, w, min_prob)

    def lse_partials(self, L):
        return orc.lse_block_partials(L)

    def finalize(self, L, partials_all, K_total, lam):
        prob_d = (orc.lse_combine(partials_all) - torch.log(torch.tensor(float(K_total), dtype=torch.float64))).float()
        return L - lam * prob_d


class HostExchange:
    """Stands in for PeerScoreExchange (whose buffers are CUDA symmetric memory) to cover the wiring of the exchange
    branch of pmi_scores_sharded: same attributes, same call, the push replaced by a gloo all_gather."""

    class _Sim:
        pass

    def __init__(self, sizes, C):
        self.sizes, self.C, self.calls = [int(x) for x in sizes], int(C), 0

    def exchange(self, L, partials_all, lam, sim, wait=True):
        self.calls += 1
        local = OracleBackend().finalize(L, partials_all, sum(self.sizes), lam)
        full = mdist._all_gather_var(local, self.sizes, None)
        return full if wait else mdist.GatheredScores(full, None)


class HostPartialsExchange:
    """Stands in for PeerPartialsExchange (CUDA symmetric memory): same attributes and call, the peer stores replaced by
    a gloo all_gather of the variable-length partial tables."""

    def __init__(self, sizes, C):
        self.sizes, self.C, self.calls = [int(x) for x in sizes], int(C), 0

    def gather(self, part, sim):
        self.calls += 1
        nblocks = [(s + mdist.LSE_BLOCK - 1) // mdist.LSE_BLOCK for s in self.sizes]
        return mdist._all_gather_var(part, nblocks, None)


def _inputs(K):
    g = torch.Generator().manual_seed(3)
    return torch.randn(300, 41, generator=g) * 0.1, torch.randn(300, K, generator=g)


def _unsharded(P, A, soft):
    from mammo_clip_dissect_b200.similarity import _reference_ramp
    be = OracleBackend()
    ramp = _reference_ramp(20, 0.998, 0.97) if soft else None
    L = be.log_sums(P, A, 20, 10 if soft else 2, 1e-7, ramp)
    return be.finalize(L, be.lse_partials(L), A.shape[1], 1 if soft else 0.6)


def _worker(rank, world, port, K, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P, A = _inputs(K)
        b = mdist.shard_bounds(K, world)
        sizes = [b[i + 1] - b[i] for i in range(world)]
        shard = A[:, b[rank]:b[rank + 1]].contiguous()
        full = mdist.soft_wpmi_sharded(P, shard, sizes, top_k=20, backend=OracleBackend())
        local = mdist.soft_wpmi_sharded(P, shard, sizes, top_k=20, backend=OracleBackend(), gather_scores=False)
        w = mdist.wpmi_sharded(P, shard, sizes, top_k=20, backend=OracleBackend())
        be = OracleBackend()
        be.sim = None
        ex = HostExchange(sizes, P.shape[1])
        via_ex = mdist.soft_wpmi_sharded(P, shard, sizes, top_k=20, backend=be, exchange=ex)
        handle = mdist.soft_wpmi_sharded(P, shard, sizes, top_k=20, backend=be, exchange=ex, wait=False)
        assert ex.calls == 2 and torch.equal(via_ex, full) and torch.equal(handle.wait(), full)
        pex = HostPartialsExchange(sizes, P.shape[1])
        via_pex = mdist.soft_wpmi_sharded(P, shard, sizes, top_k=20, backend=be, partials_exchange=pex)
        both = mdist.soft_wpmi_sharded(P, shard, sizes, top_k=20, backend=be, exchange=ex, partials_exchange=pex)
        assert pex.calls == 2 and ex.calls == 3 and torch.equal(via_pex, full) and torch.equal(both, full)
        try:
            mdist.soft_wpmi_sharded(P, shard, sizes, top_k=20, backend=be, partials_exchange=HostPartialsExchange([1] * world, 41))
            raise AssertionError("a partials exchange built for other shard sizes must be refused")
        except RuntimeError:
            pass
        try:
            mdist.soft_wpmi_sharded(P, shard, sizes, top_k=20, backend=be, exchange=HostExchange([1] * world, 41))
            raise AssertionError("an exchange built for other shard sizes must be refused")
        except RuntimeError:
            pass
        q.put((rank, full.numpy(), local.numpy(), w.numpy(), b))      # numpy: no shared-memory handles to outlive the child
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,K", [(2, 1024), (2, 700), (3, 1300), (3, 300), (2, 100)])   # last two: empty shards
def test_sharded_equals_unsharded(world, K):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, K, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    P, A = _inputs(K)
    want, want_w = _unsharded(P, A, True), _unsharded(P, A, False)
    for rank, full, local, w, b in got:
        full, local, w = torch.from_numpy(full), torch.from_numpy(local), torch.from_numpy(w)
        assert torch.equal(full, want), rank                       # G-invariant, bit for bit
        assert torch.equal(local, want[b[rank]:b[rank + 1]])
        assert torch.equal(w, want_w)
    # and the block-LSE formulation agrees with the reference's logsumexp formulation
    ref = orc.soft_wpmi_fast(P, A, top_k=20)
    assert (want - ref).abs().max().item() < 1e-3


def test_shard_bounds():
    assert mdist.shard_bounds(32768, 8) == [4096 * i for i in range(9)]
    b = mdist.shard_bounds(1300, 3)
    assert b[0] == 0 and b[-1] == 1300 and all(x % 256 == 0 for x in b[1:-1]) and b == sorted(b)
    assert mdist.shard_bounds(100, 4) == [0, 100, 100, 100, 100]
    assert mdist.shard_bounds(300, 4) == [0, 256, 300, 300, 300]      # K < 256 * (world - 1): trailing empty shards
