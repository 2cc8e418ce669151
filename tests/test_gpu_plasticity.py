"""Parity of the spike-triggered plasticity kernels with the oracle and the
reference-generated fixture.  One add per slot: results are bit-exact."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import oracle as O  # noqa: E402
from tests import cases, golden_cases  # noqa: E402


@pytest.fixture(scope="module")
def G():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from tests import gpu_api
    return gpu_api


@pytest.mark.parametrize("tag,fn", list(golden_cases.plasticity_cases()), ids=lambda x: x if isinstance(x, str) else "")
def test_plasticity_fixture(G, tag, fn):
    np.testing.assert_array_equal(fn(G), golden_cases.load("float_ops")[tag])


@pytest.mark.parametrize("m,k,nnz", [(1, 1, 1), (300, 200, 9), (5000, 7000, 60), (2000, 1500, 700)])
@pytest.mark.parametrize("float_spikes", [False, True])
@pytest.mark.parametrize("bounds", [(None, None), (-0.5, 0.8), (None, 0.3)])
def test_csr_plasticity(G, m, k, nnz, float_spikes, bounds):
    rng = np.random.default_rng(100)
    c = cases.csr_case(101, m, k, nnz, 0.1, homo=False, ragged=True)
    pre, post = cases.spikes(102, m, 0.2, float_spikes), cases.spikes(103, k, 0.2, float_spikes)
    tr_pre, tr_post = rng.standard_normal(m).astype(np.float32), rng.standard_normal(k).astype(np.float32)
    lo, hi = bounds
    np.testing.assert_array_equal(
        G.update_csr_on_binary_pre(c["data"], c["indices"], c["indptr"], pre, tr_post, lo, hi, shape=c["shape"]),
        O.update_csr_on_binary_pre(c["data"], c["indices"], c["indptr"], pre, tr_post, lo, hi, shape=c["shape"]))
    cp, ci, perm = O.csr_to_csc_index(c["indptr"], c["indices"], shape=c["shape"])
    for pd in (np.int64, np.int32):
        np.testing.assert_array_equal(
            G.update_csr_on_binary_post(c["data"], ci, cp.astype(pd), perm.astype(pd), tr_pre, post, lo, hi, shape=c["shape"]),
            O.update_csr_on_binary_post(c["data"], ci, cp, perm, tr_pre, post, lo, hi, shape=c["shape"]))


@pytest.mark.parametrize("n_rows,n_post,n_conn", [(250, 180, 11), (4000, 4000, 80), (3000, 5000, 1000)])
def test_fcn_and_dense_plasticity(G, n_rows, n_post, n_conn):
    rng = np.random.default_rng(104)
    f = cases.fcn_case(105, n_rows, n_post, n_conn)
    spk = cases.spikes(106, n_rows, 0.1, True)
    tr = rng.standard_normal(n_post).astype(np.float32)
    np.testing.assert_array_equal(G.fcn_plasticity_row(f["weights"], f["indices"], spk, tr, -1.0, 1.0),
                                  O.fcn_plasticity_row(f["weights"], f["indices"], spk, tr, -1.0, 1.0))
    W = rng.standard_normal((min(n_rows, 700), min(n_post, 900))).astype(np.float32)
    sp, so = cases.spikes(107, W.shape[0], 0.1), cases.spikes(108, W.shape[1], 0.1, True)
    tp, to = rng.standard_normal(W.shape[1]).astype(np.float32), rng.standard_normal(W.shape[0]).astype(np.float32)
    np.testing.assert_array_equal(G.update_dense_on_binary_pre(W, sp, tp, None, 0.5), O.update_dense_on_binary_pre(W, sp, tp, None, 0.5))
    np.testing.assert_array_equal(G.update_dense_on_binary_post(W, to, so, -0.5, None), O.update_dense_on_binary_post(W, to, so, -0.5, None))


def test_csc_routes_inplace_and_errors(G):
    import brainevent_b200 as be
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda")
    rng = np.random.default_rng(109)
    c = cases.csr_case(110, 400, 300, 12, 0.1, homo=False, ragged=True)
    m, k = c["shape"]
    csr = be.CSR((t(c["data"]), t(c["indices"]), t(c["indptr"])), shape=c["shape"])
    dense = csr.todense()
    pre, post = t(cases.spikes(111, m, 0.2)), t(cases.spikes(112, k, 0.2))
    tr_pre, tr_post = t(rng.standard_normal(m).astype(np.float32)), t(rng.standard_normal(k).astype(np.float32))
    # multiplicity of every (pre, post) pair (the ragged test matrix may hold duplicate synapses)
    mask = be.CSR((torch.ones_like(csr.data), csr.indices, csr.indptr), shape=c["shape"]).todense()
    # CSC-ordered weights of the same matrix: both directions against dense algebra
    cp, ci, perm = be.csr_to_csc_index(csr.indptr, csr.indices, shape=csr.shape)
    w_csc = csr.data[perm.long()]
    new = be.update_csc_on_binary_pre(w_csc, ci, cp, pre, tr_post, shape=(m, k))
    want = dense + mask * (pre.float()[:, None] * tr_post[None, :])
    got = be.CSC((new, ci, cp), shape=(m, k)).todense()
    torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-6)
    new = be.update_csc_on_binary_post(w_csc, ci, cp, tr_pre, post, shape=(m, k))
    want = dense + mask * (tr_pre[:, None] * post.float()[None, :])
    torch.testing.assert_close(be.CSC((new, ci, cp), shape=(m, k)).todense(), want, rtol=1e-6, atol=1e-6)
    # FixedNumPerPost rows
    f = cases.fcn_case(113, 300, 400, 7)
    new = be.update_fixed_pre_conn_on_binary_post(t(f["weights"]), t(f["indices"]), tr_pre, post, shape=(400, 300))
    np.testing.assert_array_equal(new.cpu().numpy(), O.fcn_plasticity_row(f["weights"], f["indices"], post.cpu().numpy(), tr_pre.cpu().numpy()))
    # functional by default, in place on request
    w = t(c["data"]).clone()
    out = be.update_csr_on_binary_pre(w, csr.indices, csr.indptr, pre, tr_post, shape=c["shape"])
    assert out.data_ptr() != w.data_ptr() and torch.equal(w, t(c["data"]))
    out2 = be.update_csr_on_binary_pre(w, csr.indices, csr.indptr, pre, tr_post, shape=c["shape"], inplace=True)
    assert out2.data_ptr() == w.data_ptr() and torch.equal(out, w)
    with pytest.raises(ValueError, match="heterogeneous"):
        be.update_csr_on_binary_pre(torch.ones(1, device="cuda"), csr.indices, csr.indptr, pre, tr_post, shape=c["shape"])
