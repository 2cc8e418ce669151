"""Post-sharded projection object (csrc/sharded.cu) on ONE GPU: world = 1, and `world` ranks emulated on the
same device (every rank's publish, then every rank's consume, so no kernel ever waits on a later launch).
Each rank's output block is compared with the oracle's UNSHARDED product; the gloo tests in
test_dist_cpu.py cover the host logic, tools/mgpu_check.py and bench.py's parity block the real N > 1 run."""
import numpy as np
import pytest
import torch

import oracle as O
from tests import cases

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 2e-5


@pytest.fixture(scope="module")
def G():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import brainevent_b200 as be
    from brainevent_b200 import _lib
    _lib.check(_lib.lib().be_b200_check_device(), "check_device")
    return be


def t(a, dtype=None):
    x = torch.as_tensor(np.ascontiguousarray(a), device="cuda")
    return x if dtype is None else x.to(dtype)


def _csr_shards(bd, c, n_post, world):
    return bd.partition_csr_by_post(c["indptr"], c["indices"], c["data"], n_post, world)


def _run_emulated(projs, bounds, spikes):
    """one timestep over all emulated ranks: publish round, then consume round"""
    outs = [p.publish(t(spikes[lo:hi])) for p, (lo, hi) in zip(projs, bounds)]
    return [p.consume(o).cpu().numpy() for p, o in zip(projs, outs)]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("fused", [1, 0])
def test_sharded_csr_ragged(G, world, fused):
    """ragged CSR shards (incl. empty rows and rows above the whole-CTA threshold) behind the exchange"""
    from brainevent_b200 import dist as bd
    n = 20_000 + 37                      # not a multiple of 32 * world: the last block is short
    c = cases.csr_case(11, n, n, 24, 0.0, ragged=True)
    # one very heavy row (> 8192 synapses in a single shard at world 1..3)
    rng = np.random.default_rng(5)
    heavy = rng.integers(0, n, 30_000).astype(np.int32)
    counts = np.diff(c["indptr"]).copy()
    r = 4321
    indices = np.concatenate([c["indices"][: c["indptr"][r]], heavy, c["indices"][c["indptr"][r + 1]:]])
    data = np.concatenate([c["data"][: c["indptr"][r]], rng.standard_normal(heavy.size).astype(np.float32),
                           c["data"][c["indptr"][r + 1]:]])
    counts[r] = heavy.size
    indptr = np.zeros(n + 1, np.int32)
    np.cumsum(counts, out=indptr[1:])
    c = dict(data=data, indices=indices, indptr=indptr)
    xs = bd.P2PSpikeExchange.emulate(n, world)
    bounds = xs[0].bounds
    projs = []
    for x, (ip, ix, dw, (lo, hi)) in zip(xs, _csr_shards(bd, c, n, world)):
        p = bd.PostShardedProjection("csr", n, n, x, data=t(dw), indices=t(ix), indptr=t(ip))
        p.set_option("fused", fused)
        projs.append(p)
    for step, rate in enumerate([0.02, 0.0, 0.3, 0.005, 1.0, 0.02, 0.02]):     # odd count: parity lives on the device
        spikes = cases.spikes(100 + step, n, rate)
        spikes[4321] = step % 2 == 0
        ref = O.binary_csrmv(c["data"], c["indices"], c["indptr"], spikes, shape=(n, n), transpose=True)
        outs = _run_emulated(projs, bounds, spikes)
        for o, (lo, hi) in zip(outs, bounds):
            np.testing.assert_allclose(o, ref[lo:hi], rtol=1e-4, atol=2e-3 if rate > 0.2 else 1e-4)
    for x in xs:
        assert not x.timed_out()


@pytest.mark.parametrize("world,homo", [(1, False), (4, False), (2, True)])
def test_sharded_fcn(G, world, homo):
    from brainevent_b200 import dist as bd
    n, conn = 8192 + 64, 40
    xs = bd.P2PSpikeExchange.emulate(n, world)
    bounds = xs[0].bounds
    projs, mats = [], []
    for r, x in enumerate(xs):
        lo, hi = bounds[r]
        f = cases.fcn_case(30 + r, n, hi - lo, conn, homo=homo)
        mats.append(f)
        projs.append(bd.PostShardedProjection("fcn", n, n, x, weights=t(f["weights"]), indices=t(f["indices"])))
    for step, rate in enumerate([0.05, 0.0, 0.5]):
        spikes = cases.spikes(200 + step, n, rate)
        outs = _run_emulated(projs, bounds, spikes)
        for o, f, (lo, hi) in zip(outs, mats, bounds):
            ref = O.binary_fcnmv(f["weights"], f["indices"], spikes, shape=(n, hi - lo), transpose=True)
            np.testing.assert_allclose(o, ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("world", [1, 3])
def test_sharded_float_spikes_and_f64(G, world):
    """float32 spikes (> 0 is active) and float64 weights through the same object"""
    from brainevent_b200 import dist as bd
    n = 5000
    c = cases.csr_case(12, n, n, 10, 0.0, ragged=True, dtype=np.float64)
    xs = bd.P2PSpikeExchange.emulate(n, world)
    bounds = xs[0].bounds
    projs = [bd.PostShardedProjection("csr", n, n, x, data=t(dw), indices=t(ix), indptr=t(ip))
             for x, (ip, ix, dw, _) in zip(xs, _csr_shards(bd, c, n, world))]
    spikes = cases.spikes(7, n, 0.1, float_spikes=True)
    ref = O.binary_csrmv(c["data"], c["indices"], c["indptr"], spikes, shape=(n, n), transpose=True)
    outs = [p.publish(t(spikes[lo:hi])) for p, (lo, hi) in zip(projs, bounds)]
    for p, o, (lo, hi) in zip(projs, outs, bounds):
        got = p.consume(o)
        assert got.dtype == torch.float64
        np.testing.assert_allclose(got.cpu().numpy(), ref[lo:hi], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("world", [1, 2, 8])
@pytest.mark.parametrize("corder", [False, True])
def test_sharded_jitc(G, world, corder):
    """JITC post-sharding: the union of the ranks' windows is index-for-index the single-GPU matrix (P = 8
    splits a chunk of the scatter order: the upper half replays the chunk prefix)"""
    from brainevent_b200 import dist as bd
    n = 4096 + 96
    xs = bd.P2PSpikeExchange.emulate(n, world)
    bounds = xs[0].bounds
    projs = [bd.PostShardedProjection("jitc", n, n, x, family="normal", params=(1.0, 0.1), prob=0.01, seed=5,
                                      corder=corder) for x in xs]
    for step in range(2):
        spikes = cases.spikes(300 + step, n, 0.05)
        ref = O.binary_jitnmv(1.0, 0.1, 0.01, spikes, 5, shape=(n, n), transpose=True, corder=corder)
        outs = _run_emulated(projs, bounds, spikes)
        for o, (lo, hi) in zip(outs, bounds):
            np.testing.assert_allclose(o, ref[lo:hi], rtol=RTOL, atol=3e-4)     # fastmath normal01 of the fixture side


@pytest.mark.parametrize("world", [1, 4])
def test_sharded_dense(G, world):
    """SURVEY 8(e) row "dense event matmul: column (post) split of W + all-gather spikes" """
    from brainevent_b200 import dist as bd
    n_pre, n_post = 3000, 3000
    rng = np.random.default_rng(9)
    W = rng.standard_normal((n_pre, n_post)).astype(np.float32)
    xs = bd.P2PSpikeExchange.emulate(n_pre, world)
    bounds = xs[0].bounds
    projs = [bd.PostShardedProjection("dense", n_pre, n_post, x, W=t(W[:, lo:hi])) for x, (lo, hi) in zip(xs, bounds)]
    for step, rate in enumerate([0.05, 0.0, 0.6]):
        spikes = cases.spikes(400 + step, n_pre, rate)
        ref = O.binary_densemv(W, spikes, transpose=True)
        outs = _run_emulated(projs, bounds, spikes)
        for o, (lo, hi) in zip(outs, bounds):
            np.testing.assert_allclose(o, ref[lo:hi], rtol=1e-4, atol=1e-3)


def test_host_pipeline_world1(G):
    """be_b200_sharded_submit_host / _wait / _run_host: host spikes in, host results out, 4 in flight"""
    from brainevent_b200 import dist as bd
    n, conn = 50_000, 32
    f = cases.fcn_case(3, n, n, conn)
    x = bd.P2PSpikeExchange(n, world=1, rank=0)
    proj = bd.PostShardedProjection("fcn", n, n, x, weights=t(f["weights"]), indices=t(f["indices"]))
    pipe = bd.HostStepPipeline(proj, n, slots=4)
    assert pipe.native
    T = 11
    h_in = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(T)]
    h_out = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(T)]
    refs = []
    for q in range(T):
        sp = cases.spikes(500 + q, n, 0.03)
        h_in[q].copy_(torch.from_numpy(sp.astype(np.uint8)))
        refs.append(O.binary_fcnmv(f["weights"], f["indices"], sp, shape=(n, n), transpose=True))
    for q in range(T):
        pipe.submit(h_in[q], h_out[q])
    pipe.wait()
    for q in range(T):
        np.testing.assert_allclose(h_out[q].numpy(), refs[q], rtol=1e-4, atol=1e-4)
    for o in h_out:
        o.zero_()
    pipe.run(T, h_in, h_out)                    # one C call for the whole block of timesteps
    for q in range(T):
        np.testing.assert_allclose(h_out[q].numpy(), refs[q], rtol=1e-4, atol=1e-4)
    pipe.close()


def test_exchange_timeout_is_fatal_and_consumes_nothing(G):
    """a flag wait that gives up must not consume stale words and must surface on the host.
    Two emulated ranks, only rank 0 publishes; its consume waits 30 ms for rank 1 and gives up."""
    from brainevent_b200 import dist as bd
    n = 4096
    c = cases.csr_case(2, n, n, 8, 0.0)
    xs = bd.P2PSpikeExchange.emulate(n, 2)
    bounds = xs[0].bounds
    shards = _csr_shards(bd, c, n, 2)
    for fused in (1, 0):
        x = bd.P2PSpikeExchange.emulate(n, 2)[0]
        ip, ix, dw, (lo, hi) = shards[0]
        p = bd.PostShardedProjection("csr", n, n, x, data=t(dw), indices=t(ix), indptr=t(ip))
        p.set_option("fused", fused)
        p.set_option("spin_ms", 30)
        spikes = cases.spikes(1, n, 0.2)
        out = p.consume(p.publish(t(spikes[lo:hi])))
        torch.cuda.synchronize()
        assert x.timed_out()
        # rank 1 never published: nothing of ITS block may have been scattered.  The fused kernel has already
        # propagated the rank's own rows (they need no exchange); the unfused path scatters nothing at all.
        own_only = spikes.copy()
        own_only[bounds[1][0]:] = False
        ref = O.binary_csrmv(c["data"], c["indices"], c["indptr"], own_only, shape=(n, n), transpose=True)[lo:hi]
        if fused:
            np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-4, atol=1e-4)
        else:
            assert float(out.abs().max().item()) == 0.0
        with pytest.raises(bd.ExchangeTimeout):
            x.check()
        # sticky: later steps do no exchange work either
        out2 = p.consume(p.publish(t(spikes[lo:hi])))
        torch.cuda.synchronize()
        assert x.timed_out()
        assert float(out2.abs().max().item()) == 0.0 or fused


def test_layout_assertions(G):
    """mismatched lengths are refused instead of shifting ids / reading n_active from the wrong slot"""
    from brainevent_b200 import dist as bd
    n = 2048
    f = cases.fcn_case(1, n, n, 4)
    x = bd.P2PSpikeExchange(n, world=1, rank=0)
    with pytest.raises(AssertionError):
        bd.PostShardedProjection("fcn", n + 32, n, x, weights=t(f["weights"]), indices=t(f["indices"]))
    p = bd.PostShardedProjection("fcn", n, n, x, weights=t(f["weights"]), indices=t(f["indices"]))
    with pytest.raises(AssertionError):
        p.step(torch.zeros(n - 1, dtype=torch.bool, device="cuda"))


def test_operator_on_non_current_device_stream(G):
    """the binding runs every call with the tensor's device current (single-GPU box: the guard is a
    no-op but the path is exercised; on a multi-GPU box cuda:1 tensors work while cuda:0 is current)."""
    import brainevent_b200 as be
    dev = torch.device("cuda", torch.cuda.device_count() - 1)
    c = cases.csr_case(4, 300, 200, 6, 0.0)
    v = cases.spikes(2, 300, 0.2)
    ref = O.binary_csrmv(c["data"], c["indices"], c["indptr"], v, shape=(300, 200), transpose=True)
    with torch.cuda.device(0):
        y = be.binary_csrmv(torch.as_tensor(c["data"], device=dev), torch.as_tensor(c["indices"], device=dev),
                            torch.as_tensor(c["indptr"], device=dev), torch.as_tensor(v, device=dev),
                            shape=(300, 200), transpose=True)
    assert y.device == dev
    np.testing.assert_allclose(y.cpu().numpy(), ref, rtol=1e-5, atol=1e-5)
    # B[64, k] @ W needs > 48 KB of dynamic shared memory: the per-device function attribute must be set
    W = np.random.default_rng(0).standard_normal((256, 512)).astype(np.float64)
    B = cases.spikes(3, 256, 0.1, batch=8)
    Y = be.binary_densemm(torch.as_tensor(W, device=dev), torch.as_tensor(B, device=dev), transpose=True)
    np.testing.assert_allclose(Y.cpu().numpy(), O.binary_densemm(W, B, transpose=True), rtol=1e-10, atol=1e-10)
