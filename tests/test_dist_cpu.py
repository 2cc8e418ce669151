"""world_size-2 gloo test of the post-sharding host logic (partitioning + packed-spike all-gather).
The CUDA kernels cannot run here, so each rank's propagate step is played by the oracle on its
shard; what is checked is that shards + exchange reproduce the unsharded result exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O
from brainevent_b200 import dist as bd
from tests import cases


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, n_conn, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        f = cases.fcn_case(3, n, n, n_conn)
        spikes = cases.spikes(4, n, 0.1)
        shards = bd.partition_fcn_by_post(f["indices"], f["weights"], n, world)
        ip, ix, dw, (lo, hi) = shards[rank]
        x = bd.SpikeExchange(n, device="cpu")
        assert x.local == (lo, hi)
        local = spikes[lo:hi]
        words = O.bitpack(local, 0).view(np.int32)
        slot = np.zeros(x.words_per_rank, np.int32)
        slot[: words.size] = words
        packed_all = x.exchange(torch.from_numpy(slot)).numpy().view(np.uint32)
        np.testing.assert_array_equal(packed_all, O.bitpack(spikes, 0))
        bits = ((packed_all[:, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool).reshape(-1)[:n]
        np.testing.assert_array_equal(bits, spikes)
        out_local = O.binary_csrmv(dw, ix, ip, bits, shape=(n, hi - lo), transpose=True)
        full = O.binary_fcnmv(f["weights"], f["indices"], spikes, shape=(n, n), transpose=True)
        np.testing.assert_allclose(out_local, full[lo:hi], rtol=1e-6, atol=1e-6)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,n_conn", [(4096, 16), (1000, 7)])
def test_post_sharding_world2(n, n_conn):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, n_conn, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_bounds_are_word_aligned():
    for n, w in [(10_000_000, 8), (1000, 2), (33, 4), (1_250_016, 1)]:
        b = bd.post_block_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n
        for (lo, hi), (lo2, _) in zip(b, b[1:]):
            assert hi == lo2 and (hi % 32 == 0 or hi == n)


def test_partition_csr_preserves_everything():
    c = cases.csr_case(5, 300, 257, 9, 0.1, ragged=True)
    shards = bd.partition_csr_by_post(c["indptr"], c["indices"], c["data"], 257, 3)
    v = cases.spikes(6, 300, 0.2)
    full = O.binary_csrmv(c["data"], c["indices"], c["indptr"], v, shape=(300, 257), transpose=True)
    got = np.concatenate([O.binary_csrmv(d, ix, ip, v, shape=(300, hi - lo), transpose=True)
                          for ip, ix, d, (lo, hi) in shards])
    np.testing.assert_allclose(got, full, rtol=1e-6, atol=1e-6)
    assert sum(ix.size for _, ix, _, _ in shards) == c["indices"].size


def test_c5_generator_shards_union_is_the_single_gpu_matrix():
    """bench.py's SURVEY 8(d) generator: every rank draws the same chunk stream and keeps its post block; the union of
    the ragged CSR shards is the FixedNumPerPre matrix a single GPU builds (strong scaling compares the same network)."""
    import bench
    old = bench.C5_CHUNK_ROWS
    bench.C5_CHUNK_ROWS = 128
    try:
        n, conn = 1000, 12
        idx, w = bench.build_c5_ell(torch, "cpu", n, conn)
        assert idx.shape == (n, conn) and int(idx.min()) >= 0 and int(idx.max()) < n
        total = 0
        for lo, hi in bd.post_block_bounds(n, 3):
            ip, ix, ww, counts = bench.build_c5_shard(torch, "cpu", n, conn, lo, hi)
            m = (idx >= lo) & (idx < hi)
            assert torch.equal(ix, idx[m] - lo) and torch.equal(ww, w[m]) and torch.equal(counts, m.sum(1))
            assert ip.dtype == torch.int32 and int(ip[-1]) == ix.numel()
            # the shard's product equals the block of the unsharded product (oracle on both sides)
            v = cases.spikes(9, n, 0.1)
            full = O.binary_fcnmv(w.numpy(), idx.numpy(), v, shape=(n, n), transpose=True)
            part = O.binary_csrmv(ww.numpy(), ix.numpy(), ip.numpy(), v, shape=(n, hi - lo), transpose=True)
            np.testing.assert_allclose(part, full[lo:hi], rtol=1e-6, atol=1e-6)
            total += ix.numel()
        assert total == n * conn
    finally:
        bench.C5_CHUNK_ROWS = old


def test_sharded_projection_refuses_mismatched_layouts():
    """the exchange and the projection must agree on the neuron count (checked before any device work)"""
    x = bd.SpikeExchange(4096, device="cpu", world=2, rank=1)
    assert x.local == (2048, 4096) and x.words_per_rank == 64
    with pytest.raises(AssertionError):
        bd.PostShardedProjection("csr", 4000, 4096, x, data=None, indices=None, indptr=None)
    p = bd.PostShardedProjection("csr", 4096, 4096, x, data=None, indices=None, indptr=None)
    with pytest.raises(AssertionError):
        p.pack_local(torch.zeros(100, dtype=torch.bool))
