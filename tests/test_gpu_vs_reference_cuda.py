"""GPU parity against the reference's OWN CUDA kernels (unmodified sources compiled into
oracle/_ref by `make -C oracle/ref_cuda`) — a second oracle beside the CPU
restatement.  Connectivity (JITC) is compared exactly through homogeneous unit weights; float sums
within rtol 1e-5 / atol 1e-6 scaled by the row length (atomic order differs on both sides)."""
import math

import pytest
import torch

from oracle import ref_cuda as R

pytestmark = pytest.mark.gpu

needs = lambda name: pytest.mark.skipif(not R.available(name), reason=f"oracle/_ref/libref_{name}.so not built")


@pytest.fixture(scope="module")
def be():
    import brainevent_b200
    return brainevent_b200


def _close(a, b, terms, rtol=1e-5, atol=4e-6):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol * max(1.0, math.sqrt(terms)))


def _g(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


@needs("fcnmv")
@pytest.mark.parametrize("homo", [False, True])
@pytest.mark.parametrize("n_pre,n_post,conn,rate", [(5000, 7000, 64, 0.05), (20000, 20000, 1000, 0.01), (33, 17, 5, 0.5)])
def test_fcn_scatter(be, homo, n_pre, n_post, conn, rate):
    g = _g(1)
    idx = torch.randint(0, n_post, (n_pre, conn), device="cuda", dtype=torch.int32, generator=g)
    w = torch.tensor([0.37], device="cuda") if homo else torch.rand((n_pre, conn), device="cuda", generator=g)
    spk = torch.rand(n_pre, device="cuda", generator=g) < rate
    ours = be.binary_fcnmv(w, idx, spk, shape=(n_pre, n_post), transpose=True)
    ref = R.binary_fcnmv_scatter(w, idx, spk, n_post)
    _close(ours, ref, n_pre * conn * rate / n_post)


@needs("csrmv_hybrid")
@needs("csrmv")
@pytest.mark.parametrize("homo", [False, True])
@pytest.mark.parametrize("ragged", [False, True])
def test_csr_scatter_and_gather(be, homo, ragged):
    g = _g(2)
    n_pre, n_post = 6000, 5000
    if ragged:   # includes empty rows and rows above the hybrid's 128-nnz task threshold
        lens = torch.randint(0, 40, (n_pre,), device="cuda", generator=g)
        lens[::97] = 700
        lens[5] = 9000
    else:
        lens = torch.full((n_pre,), 80, device="cuda")
    indptr = torch.zeros(n_pre + 1, dtype=torch.int32, device="cuda")
    indptr[1:] = torch.cumsum(lens, 0)
    nse = int(indptr[-1])
    indices = torch.randint(0, n_post, (nse,), device="cuda", dtype=torch.int32, generator=g)
    w = torch.tensor([0.6], device="cuda") if homo else torch.randn(nse, device="cuda", generator=g)
    spk = torch.rand(n_pre, device="cuda", generator=g) < 0.1
    spk[5] = True
    ws = R.HybridWorkspace(indptr)
    ours = be.binary_csrmv(w, indices, indptr, spk, shape=(n_pre, n_post), transpose=True)
    ref = R.binary_csrmv_t_hybrid(w, indices, indptr, spk, n_post, ws)
    _close(ours, ref, 64)
    spk_post = torch.rand(n_post, device="cuda", generator=g) < 0.1
    ours = be.binary_csrmv(w, indices, indptr, spk_post, shape=(n_pre, n_post), transpose=False)
    ref = R.binary_csrmv_nt(w, indices, indptr, spk_post, n_pre)
    _close(ours, ref, 900)


@needs("csrmm_hybrid")
def test_csr_spmm_scatter(be):
    g = _g(3)
    n_pre, n_post, nnz, batch = 3000, 2500, 50, 64
    indptr = torch.arange(0, (n_pre + 1) * nnz, nnz, device="cuda", dtype=torch.int32)
    indices = torch.randint(0, n_post, (n_pre * nnz,), device="cuda", dtype=torch.int32, generator=g)
    w = torch.randn(n_pre * nnz, device="cuda", generator=g)
    B = torch.rand((n_pre, batch), device="cuda", generator=g) < 0.05
    ws = R.HybridWorkspace(indptr)
    ours = be.binary_csrmm(w, indices, indptr, B, shape=(n_pre, n_post), transpose=True)     # [n_post, batch]
    ref = R.binary_csrmm_t_hybrid(w, indices, indptr, B, n_post, ws).t()
    _close(ours, ref, 16)


@pytest.mark.parametrize("corder", [False, True])
@pytest.mark.parametrize("shape,prob", [((3000, 4100), 0.01), ((2047, 999), 0.05), ((40000, 30000), 2e-4)])
class TestJitc:
    """spk @ M for the three JITC families; unit homogeneous weights make the output an exact count
    of (active source, target) connections, so equality is index-for-index connectivity."""

    def _spk(self, n):
        return torch.rand(n, device="cuda", generator=_g(4)) < 0.1

    @needs("jitsmv")
    def test_scalar_connectivity_exact(self, be, corder, shape, prob):
        spk = self._spk(shape[0])
        w = torch.tensor([1.0], device="cuda")
        ours = be.binary_jitsmv(w, prob, spk, 42, shape=shape, transpose=True, corder=corder)
        ref = R.binary_jitsmv(w, prob, spk, 42, shape=shape, transpose=True, corder=corder)
        assert torch.equal(ours, ref)
        assert ours.sum().item() > 0

    @needs("jitnmv")
    def test_normal(self, be, corder, shape, prob):
        spk = self._spk(shape[0])
        loc, scale = torch.tensor([1.0], device="cuda"), torch.tensor([0.1], device="cuda")
        ours = be.binary_jitnmv(loc, scale, prob, spk, 123, shape=shape, transpose=True, corder=corder)
        ref = R.binary_jitnmv(loc, scale, prob, spk, 123, shape=shape, transpose=True, corder=corder)
        # Acklam's central rational cancels to ~3e-3 near |u-0.5| ~ 0.47, so one rounding of its
        # denominator moves N(0,1) by up to ~3e-4.  The reference's CUDA file is compiled with nvcc's
        # default FMA contraction and so differs from the reference's OWN strict `light_rng_normal01`
        # (which this repo matches bit-for-bit, test_jitc_normal_weight_bit_exact) by that much per
        # edge: 3e-4 * w_scale per edge here, exactly like the fastmath numba fixtures.
        _close(ours, ref, shape[0] * 0.1 * prob, rtol=1e-4, atol=3e-4 * 0.1)

    @needs("jitumv")
    def test_uniform(self, be, corder, shape, prob):
        spk = self._spk(shape[0])
        lo, hi = torch.tensor([-0.5], device="cuda"), torch.tensor([1.5], device="cuda")
        ours = be.binary_jitumv(lo, hi, prob, spk, 7, shape=shape, transpose=True, corder=corder)
        ref = R.binary_jitumv(lo, hi, prob, spk, 7, shape=shape, transpose=True, corder=corder)
        _close(ours, ref, shape[0] * 0.1 * prob)


@needs("densemm")
def test_dense_event_matmul(be):
    g = _g(5)
    k, m, batch = 1024, 768, 64
    W = torch.randn((k, m), device="cuda", generator=g)
    B = torch.rand((k, batch), device="cuda", generator=g) < 0.1
    ours = be.binary_densemm(W, B, transpose=True)             # [m, batch]
    ref = R.binary_densemm_transpose(W, B).t()
    _close(ours, ref, k * 0.1)


@needs("compact")
@pytest.mark.parametrize("n", [1, 31, 4097, 1_000_003])
def test_compact_same_id_set(be, n):
    spk = torch.rand(n, device="cuda", generator=_g(6)) < 0.03
    ids, cnt = be.binary_1d_array_index_p_call(spk)
    rids, rcnt = R.compact_1d(spk)
    k = int(rcnt.item())
    assert int(cnt.reshape(-1)[0].item()) == k
    assert torch.equal(ids[:k], torch.sort(rids[:k]).values)     # the reference's GPU order is unspecified (F7)


# ---- the BASELINE configs at FULL size, element by element against the reference's own CUDA kernels (VERDICT r1:
# test_gpu_parity checks C3 / C5 at full size by checksum and linearity only; here every output element is compared) ----
@needs("csrmm_hybrid")
def test_baseline_c3_spmm_full_size_vs_reference_cuda(be):
    n, nnz, batch = 1_000_000, 100, 64
    g = _g(31)
    indptr = torch.arange(0, (n + 1) * nnz, nnz, device="cuda", dtype=torch.int32)
    indices = torch.randint(0, n, (n * nnz,), device="cuda", dtype=torch.int32, generator=g)
    w = torch.rand(n * nnz, device="cuda", generator=g)
    B = torch.rand((n, batch), device="cuda", generator=g) < 0.01
    ours = be.binary_csrmm(w, indices, indptr, B, shape=(n, n), transpose=True)            # logical [n, batch]
    ref = R.binary_csrmm_t_hybrid(w, indices, indptr, B, n, R.HybridWorkspace(indptr))      # physical [batch, n]
    _close(ours, ref.t(), terms=nnz * 0.01 * 64)
    assert float(ref.abs().max()) > 1.0


@needs("fcnmv")
def test_baseline_c5_shard_full_size_vs_reference_cuda(be):
    """one GPU's ELL shard of the 10 M x 1000 network over 8 GPUs (10 M pre rows x 125 local targets)"""
    n_total, n_local, conn = 10_000_128, 1_250_016, 125
    g = _g(32)
    idx = torch.randint(0, n_local, (n_total, conn), device="cuda", dtype=torch.int32, generator=g)
    w = torch.rand((n_total, conn), device="cuda", generator=g)
    spk = torch.rand(n_total, device="cuda", generator=g) < 0.01
    ours = be.binary_fcnmv(w, idx, spk, shape=(n_total, n_local), transpose=True)
    ref = R.binary_fcnmv_scatter(w, idx, spk, n_local)
    _close(ours, ref, terms=10.0)
    # and the same shard through the sharded object (world 1: pack + fused exchange/scatter kernels)
    from brainevent_b200 import dist as bd
    x = bd.P2PSpikeExchange(n_total, world=1, rank=0)
    proj = bd.PostShardedProjection("fcn", n_total, n_local, x, weights=w, indices=idx)     # feed-forward: posts != pres
    _close(proj.step(spk), ref, terms=10.0)
