"""The oracle against (a) the reference's own known-answer vectors and (b) committed fixtures
generated from the unmodified reference kernels (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

import oracle as O
from tests import cases, golden_cases

RTOL = 1e-5


# ---- (a) known-answer vectors quoted from the reference's own tests ---------------------------
def test_uniform01_kat():
    # /root/reference/brainevent/_numba_random_test.py:59-69 (exact float32)
    kat = [(42, 0, 0, 0.2929498553276062), (42, 3, 7, 0.548724353313446),
           (123, 19, 29, 0.5329357385635376), (0, 1, 2, 0.31099069118499756),
           (0xFFFFFFFF, 65535, 123456, 0.8090267777442932)]
    for seed, row, col, exp in kat:
        assert O.light_rng_uniform01(seed, row, col) == np.float32(exp)


def test_normal01_kat():
    # /root/reference/brainevent/_numba_random_test.py:80-90 (rtol=atol=1e-6)
    kat = [(42, 0, 0, -0.5447874069213867), (42, 3, 7, 0.12243907153606415),
           (123, 19, 29, 0.08265165984630585), (0, 1, 2, -0.4930441081523895),
           (0xFFFFFFFF, 65535, 123456, 0.8743151426315308)]
    for seed, row, col, exp in kat:
        assert np.allclose(O.light_rng_normal01(seed, row, col), np.float32(exp), rtol=1e-6, atol=1e-6)


def test_survey_probe_values():
    # values probed from the unmodified reference _numba_random.py
    assert int(O.light_rng_init(42, 3, 0, 1)) == 0x87d859b7
    q, st = O.light_rng_initial_q(O.light_rng_init(42, 3, 0, 1), 1000)
    assert (int(q), int(st)) == (67, 0x08eafba4)


def test_packed_bits_and_active_rows_kat():
    # /root/reference/brainevent/_event/compact_test.py:182-197
    s = np.array([[1, 0, 1, 0], [0, 0, 0, 0], [0, 1, 1, 0]], bool)
    packed, ids, na = O.binary_2d_array_index(s)
    np.testing.assert_array_equal(packed[:, 0], np.array([5, 0, 6], np.uint32))
    np.testing.assert_array_equal(ids[:int(na[0])], [0, 2])


def test_small_csr_and_docstring_examples():
    # /root/reference/brainevent/_csr/_test_util.py:257-279: [[1,0,2],[0,3,0]]
    data = np.array([1, 2, 3], np.float32)
    idx = np.array([0, 2, 1], np.int32)
    ptr = np.array([0, 2, 3], np.int32)
    np.testing.assert_array_equal(
        O.binary_csrmv(data, idx, ptr, np.array([True, False, True]), shape=(2, 3)), [3, 0])
    np.testing.assert_array_equal(
        O.binary_csrmv(data, idx, ptr, np.array([True, True]), shape=(2, 3), transpose=True), [1, 3, 2])
    # /root/reference/brainevent/_event/binary.py:186-191: BinaryArray([T,F,T]) @ [[1,2],[3,4],[5,6]] = [6,8]
    W = np.array([[1, 2], [3, 4], [5, 6]], np.float32)
    np.testing.assert_array_equal(O.binary_densemv(W, np.array([True, False, True]), transpose=True), [6, 8])
    # /root/reference/brainevent/_fcn/binary.py docstring: homogeneous ones, indices [[0,1],[1,2]], spikes [T,F,T] -> [1,1]
    np.testing.assert_array_equal(
        O.binary_fcnmv(np.ones(1, np.float32), np.array([[0, 1], [1, 2]], np.int32),
                       np.array([True, False, True]), shape=(2, 3)), [1, 1])
    # /root/reference/brainevent/_dense/binary.py docstring: ones(3,4) @ [T,F,T,F] = [2,2,2]
    np.testing.assert_array_equal(
        O.binary_densemv(np.ones((3, 4), np.float32), np.array([True, False, True, False]), transpose=False), [2, 2, 2])


def test_bitpack_matches_definition():
    rng = np.random.default_rng(0)
    a = rng.random((5, 77)) < 0.3
    for axis in (0, 1):
        p = O.bitpack(a, axis=axis)
        n = a.shape[axis]
        pad = np.zeros_like(a, shape=tuple((((n + 31) // 32) * 32 if i == axis else s) for i, s in enumerate(a.shape)))
        sl = [slice(None)] * 2
        sl[axis] = slice(0, n)
        pad[tuple(sl)] = a
        pm = np.moveaxis(pad, axis, -1).reshape(-1, (n + 31) // 32, 32)
        exp = (pm.astype(np.uint64) << np.arange(32, dtype=np.uint64)).sum(-1).astype(np.uint32)
        np.testing.assert_array_equal(np.moveaxis(p, axis, -1).reshape(exp.shape), exp)


# ---- (b) fixtures generated from the reference -------------------------------------------------
def test_rng_fixture():
    g = golden_cases.load("rng")
    for i, key in enumerate(g["keys"]):
        s = O.light_rng_init(*[int(x) for x in key])
        assert int(s) == int(g["init"][i])
        for j, cl in enumerate(g["cls"]):
            q, st = O.light_rng_initial_q(s, int(cl))
            assert (int(q), int(st)) == (int(g["q"][i, j]), int(g["st"][i, j]))
        for t in range(64):
            s = O.light_rng_next(s)
            assert int(s) == int(g["stream"][i, t])
    for key, u, n in zip(g["hash_keys"], g["u01"], g["n01"]):
        key = [int(x) for x in key]
        assert O.light_rng_uniform01(*key) == u
        assert abs(float(O.light_rng_normal01(*key)) - float(n)) <= 1e-6


def test_connectivity_fixture_index_for_index():
    g = golden_cases.load("jitc_connectivity")
    for tag in ("a", "b"):
        meta = g[f"{tag}_meta"]
        m, k, clen, seed, chunk = (int(x) for x in meta[:5])
        rows = [int(x) for x in meta[5:]]
        for mode, stride in (("mv", O.MV_STRIDE), ("mm", O.MM_STRIDE)):
            for r, counts in zip(rows, g[f"{tag}_{mode}"]):
                cols = O.jitc_row_connections(seed, clen, r, k, chunk, stride)
                mine = np.bincount(cols, minlength=k)
                np.testing.assert_array_equal(mine, counts)
                assert mine.sum() > 0


def test_compaction_fixture():
    g = golden_cases.load("ops")
    ids, na = O.binary_1d_array_index(cases.spikes(1, 1000, 0.1))
    assert int(na[0]) == int(g["compact1d_n"][0])
    np.testing.assert_array_equal(ids[:na[0]], g["compact1d_ids"])
    assert not ids[na[0]:].any()
    p, ids, na = O.binary_2d_array_index(cases.spikes(2, 300, 0.02, batch=70))
    np.testing.assert_array_equal(p, g["compact2d_packed"])
    np.testing.assert_array_equal(ids[:na[0]], g["compact2d_ids"])


@pytest.mark.parametrize("tag,fn,atol", list(golden_cases.op_cases()), ids=lambda x: x if isinstance(x, str) else "")
def test_ops_fixture(tag, fn, atol):
    g = golden_cases.load("ops")
    np.testing.assert_allclose(fn(O), g[tag], rtol=RTOL, atol=atol)


@pytest.mark.parametrize("tag,fn,atol", list(golden_cases.float_op_cases()), ids=lambda x: x if isinstance(x, str) else "")
def test_float_ops_fixture(tag, fn, atol):
    g = golden_cases.load("float_ops")
    np.testing.assert_allclose(fn(O), g[tag], rtol=1e-5, atol=atol)


def test_jitc_to_csr_fixture():
    g = golden_cases.load("float_ops")
    for tag, fn, atol in golden_cases.to_csr_cases():
        data, indices, indptr = fn(O)
        np.testing.assert_array_equal(indptr, g[tag + "_indptr"])
        np.testing.assert_array_equal(indices, g[tag + "_indices"])
        np.testing.assert_allclose(data, g[tag + "_data"], rtol=1e-6, atol=atol)


@pytest.mark.parametrize("tag,fn", list(golden_cases.plasticity_cases()), ids=lambda x: x if isinstance(x, str) else "")
def test_plasticity_fixture(tag, fn):
    np.testing.assert_allclose(fn(O), golden_cases.load("float_ops")[tag], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("tag,fn", list(golden_cases.encoder_cases()), ids=lambda x: x if isinstance(x, str) else "")
def test_encoder_fixture(tag, fn):
    g = golden_cases.load("float_ops")
    for i, arr in enumerate(fn(O)):
        np.testing.assert_array_equal(arr, g[f"{tag}_{i}"])
