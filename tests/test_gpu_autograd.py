"""torch.autograd rules of the event matvec operators (brainevent_b200/_autograd.py; ref: the primitives' jvp / transpose
rules, _csr/binary.py:656-722).  The operators are linear in the weights and -- for float events -- treated as linear in
the event vector, so every cotangent has a closed form that numpy evaluates directly."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def be():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import brainevent_b200 as be
    return be


def _csr(rng, m, k, nnz):
    counts = rng.integers(0, 2 * nnz, m)
    indptr = np.zeros(m + 1, np.int32)
    np.cumsum(counts, out=indptr[1:])
    indices = rng.integers(0, k, int(indptr[-1]), dtype=np.int32)
    data = rng.standard_normal(int(indptr[-1])).astype(np.float32)
    rows = np.repeat(np.arange(m), counts)
    return data, indices, indptr, rows


@pytest.mark.parametrize("transpose", [True, False])
@pytest.mark.parametrize("homo", [False, True])
@pytest.mark.parametrize("float_events", [False, True])
def test_csrmv_grad(be, transpose, homo, float_events):
    rng = np.random.default_rng(3)
    m, k = 700, 900
    data, indices, indptr, rows = _csr(rng, m, k, 12)
    n_src, n_out = (m, k) if transpose else (k, m)
    act = rng.random(n_src) < 0.2
    ev_np = (act * rng.uniform(0.5, 2.0, n_src) - (~act) * rng.uniform(0.0, 1.0, n_src)).astype(np.float32) if float_events else act
    g = rng.standard_normal(n_out).astype(np.float32)
    t = lambda a: torch.as_tensor(a, device="cuda")
    w = t(np.float32([0.7]) if homo else data).requires_grad_(True)
    ev = t(ev_np)
    if float_events:
        ev.requires_grad_(True)
    y = be.binary_csrmv(w, t(indices), t(indptr), ev, shape=(m, k), transpose=transpose)
    ref_w = np.float32([0.7]) if homo else data
    np.testing.assert_allclose(y.detach().cpu().numpy(),
                               O.binary_csrmv(ref_w, indices, indptr, ev_np, shape=(m, k), transpose=transpose),
                               rtol=1e-5, atol=1e-5)
    (y * t(g)).sum().backward()
    src_of, dst_of = (rows, indices) if transpose else (indices, rows)
    per_slot = act[src_of] * g[dst_of]                                   # d loss / d w[slot]
    expect = np.float32([per_slot.sum()]) if homo else per_slot.astype(np.float32)
    np.testing.assert_allclose(w.grad.cpu().numpy(), expect, rtol=1e-5, atol=1e-4 if homo else 1e-6)
    if float_events:
        # the product taken as linear in the event vector: ct_v = A g (or A^T g) with the weights as stored
        full_w = np.full(indices.shape, 0.7, np.float32) if homo else data
        gv = np.zeros(n_src, np.float64)
        np.add.at(gv, src_of, full_w.astype(np.float64) * g[dst_of])
        np.testing.assert_allclose(ev.grad.cpu().numpy(), gv, rtol=1e-4, atol=1e-4)
    else:
        assert ev.grad is None


@pytest.mark.parametrize("transpose", [True, False])
@pytest.mark.parametrize("homo", [False, True])
def test_fcnmv_grad(be, transpose, homo):
    rng = np.random.default_rng(5)
    n_rows, n_other, n_conn = 600, 800, 9
    idx = rng.integers(0, n_other, (n_rows, n_conn)).astype(np.int32)
    wv = rng.standard_normal((n_rows, n_conn)).astype(np.float32)
    n_src, n_out = (n_rows, n_other) if transpose else (n_other, n_rows)
    act = rng.random(n_src) < 0.15
    g = rng.standard_normal(n_out).astype(np.float32)
    t = lambda a: torch.as_tensor(a, device="cuda")
    w = t(np.float32([1.3]) if homo else wv).requires_grad_(True)
    y = be.binary_fcnmv(w, t(idx), t(act), shape=(n_rows, n_other), transpose=transpose)
    (y * t(g)).sum().backward()
    per_slot = (act[:, None] * g[idx]) if transpose else (act[idx] * g[:, None])
    expect = np.float32([per_slot.sum()]) if homo else per_slot.astype(np.float32)
    np.testing.assert_allclose(w.grad.cpu().numpy(), expect, rtol=1e-5, atol=1e-4 if homo else 1e-6)


@pytest.mark.parametrize("transpose", [True, False])
def test_densemv_grad(be, transpose):
    rng = np.random.default_rng(7)
    k, m = 300, 500
    W = rng.standard_normal((k, m)).astype(np.float32)
    n_src, n_out = (k, m) if transpose else (m, k)
    act = rng.random(n_src) < 0.1
    g = rng.standard_normal(n_out).astype(np.float32)
    t = lambda a: torch.as_tensor(a, device="cuda")
    w = t(W).requires_grad_(True)
    y = be.binary_densemv(w, t(act), transpose=transpose)
    (y * t(g)).sum().backward()
    expect = np.outer(act, g) if transpose else np.outer(g, act)
    np.testing.assert_allclose(w.grad.cpu().numpy(), expect.astype(np.float32), rtol=1e-6, atol=1e-6)


def test_no_grad_path_is_unchanged(be):
    """without requires_grad the operators bypass the autograd wrappers (same kernels, no graph)"""
    rng = np.random.default_rng(9)
    data, indices, indptr, _ = _csr(rng, 100, 120, 5)
    t = lambda a: torch.as_tensor(a, device="cuda")
    y = be.binary_csrmv(t(data), t(indices), t(indptr), t(rng.random(100) < 0.3), shape=(100, 120), transpose=True)
    assert not y.requires_grad and y.grad_fn is None


def test_grad_through_the_matrix_classes(be):
    """`event_array @ matrix` of the mirrored classes reaches the same rules"""
    rng = np.random.default_rng(11)
    m, k = 400, 300
    data, indices, indptr, rows = _csr(rng, m, k, 6)
    act = rng.random(m) < 0.25
    g = rng.standard_normal(k).astype(np.float32)
    t = lambda a: torch.as_tensor(a, device="cuda")
    w = t(data).requires_grad_(True)
    y = be.BinaryArray(t(act)) @ be.CSR((w, t(indices), t(indptr)), shape=(m, k))
    (y * t(g)).sum().backward()
    np.testing.assert_allclose(w.grad.cpu().numpy(), (act[rows] * g[indices]).astype(np.float32), rtol=1e-5, atol=1e-6)
    idx = rng.integers(0, k, (m, 7)).astype(np.int32)
    wf = t(rng.standard_normal((m, 7)).astype(np.float32)).requires_grad_(True)
    y = be.BinaryArray(t(act)) @ be.FixedNumPerPre((wf, t(idx)), shape=(m, k))
    (y * t(g)).sum().backward()
    np.testing.assert_allclose(wf.grad.cpu().numpy(), (act[:, None] * g[idx]).astype(np.float32), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("family", ["scalar", "normal", "uniform"])
@pytest.mark.parametrize("corder", [True, False])
def test_jitc_parameter_grads(be, family, corder):
    """JITC products are linear in their weight parameters: d<g, y>/dp_i = <g, y(e_i)>, checked against the oracle's
    product with unit parameters."""
    rng = np.random.default_rng(13)
    shape, prob, seed = (900, 1100), 0.02, 77
    act = rng.random(shape[0]) < 0.2
    g = rng.standard_normal(shape[1]).astype(np.float32)
    t = lambda a: torch.as_tensor(a, device="cuda")
    if family == "scalar":
        ps = [t(np.float32([0.8])).requires_grad_(True)]
        y = be.binary_jitsmv(ps[0], prob, t(act), seed, shape=shape, transpose=True, corder=corder)
        units = [O.binary_jitsmv(1.0, prob, act, seed, shape=shape, transpose=True, corder=corder)]
    elif family == "normal":
        ps = [t(np.float32([0.5])).requires_grad_(True), t(np.float32([0.2])).requires_grad_(True)]
        y = be.binary_jitnmv(ps[0], ps[1], prob, t(act), seed, shape=shape, transpose=True, corder=corder)
        units = [O.binary_jitnmv(1.0, 0.0, prob, act, seed, shape=shape, transpose=True, corder=corder),
                 O.binary_jitnmv(0.0, 1.0, prob, act, seed, shape=shape, transpose=True, corder=corder)]
    else:
        ps = [t(np.float32([-0.3])).requires_grad_(True), t(np.float32([0.9])).requires_grad_(True)]
        y = be.binary_jitumv(ps[0], ps[1], prob, t(act), seed, shape=shape, transpose=True, corder=corder)
        units = [O.binary_jitumv(1.0, 0.0, prob, act, seed, shape=shape, transpose=True, corder=corder),
                 O.binary_jitumv(0.0, 1.0, prob, act, seed, shape=shape, transpose=True, corder=corder)]
    (y * t(g)).sum().backward()
    for p, u in zip(ps, units):
        expect = float(np.dot(g.astype(np.float64), np.asarray(u, np.float64)))
        np.testing.assert_allclose(float(p.grad.item()), expect, rtol=2e-4, atol=2e-3)
