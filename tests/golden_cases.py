"""Re-creates the inputs of tests/golden/ops.npz (same seeds as make_golden.py) and yields
(tag, callable(impl) -> result, golden, atol) so both the oracle and the CUDA path are checked
against the reference's own outputs through one table."""
import os

import numpy as np

from tests import cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def op_cases():
    """Yield (tag, fn(api) -> ndarray, atol) for every entry of ops.npz that is a float result."""
    for homo in (False, True):
        c = cases.csr_case(4, 300, 200, 9, 0.1, homo=homo, ragged=True)
        m, k = c["shape"]
        for tr in (False, True):
            for fs in (False, True):
                tag = f"csr_h{int(homo)}_t{int(tr)}_f{int(fs)}"
                v = cases.spikes(5, m if tr else k, 0.15, fs)
                B = cases.spikes(6, m if tr else k, 0.1, fs, batch=13)
                yield (tag + "_mv", (lambda api, c=c, v=v, tr=tr: api.binary_csrmv(
                    c["data"], c["indices"], c["indptr"], v, shape=c["shape"], transpose=tr)), 1e-6)
                yield (tag + "_mm", (lambda api, c=c, B=B, tr=tr: api.binary_csrmm(
                    c["data"], c["indices"], c["indptr"], B, shape=c["shape"], transpose=tr)), 1e-6)
    for homo in (False, True):
        c = cases.fcn_case(9, 250, 180, 11, homo=homo)
        for tr in (False, True):
            tag = f"fcn_h{int(homo)}_t{int(tr)}"
            v = cases.spikes(10, 250 if tr else 180, 0.15)
            B = cases.spikes(11, 250 if tr else 180, 0.1, batch=9)
            yield (tag + "_mv", (lambda api, c=c, v=v, tr=tr: api.binary_fcnmv(
                c["weights"], c["indices"], v, shape=(250, 180), transpose=tr)), 1e-6)
            yield (tag + "_mm", (lambda api, c=c, B=B, tr=tr: api.binary_fcnmm(
                c["weights"], c["indices"], B, shape=(250, 180), transpose=tr)), 1e-6)
    shape, prob, seed = (230, 310), 0.05, 1234
    for fam, params in [("s", (1.5,)), ("n", (1.0, 0.1)), ("u", (-0.5, 1.0))]:
        # normal: the golden values come from the reference's fastmath numba build, which
        # deviates from its own strict light_rng_normal01 by up to 3e-4 per edge
        # (see tests/test_oracle_vs_reference.py::test_jitc).
        atol = 3e-4 if fam == "n" else 2e-6
        for tr in (False, True):
            for co in (False, True):
                tag = f"jit{fam}_t{int(tr)}_c{int(co)}"
                vec_len = shape[0] if tr else shape[1]
                v = cases.spikes(12, vec_len, 0.2)
                B = cases.spikes(13, vec_len, 0.2, batch=6)
                yield (tag + "_mv", (lambda api, fam=fam, params=params, v=v, tr=tr, co=co: getattr(
                    api, f"binary_jit{fam}mv")(*params, prob, v, seed, shape=shape, transpose=tr,
                                               corder=co)), atol)
                yield (tag + "_mm", (lambda api, fam=fam, params=params, B=B, tr=tr, co=co: getattr(
                    api, f"binary_jit{fam}mm")(*params, prob, B, seed, shape=shape, transpose=tr,
                                               corder=co)), atol)
    W = np.random.default_rng(14).standard_normal((90, 130)).astype(np.float32)
    for tr in (False, True):
        k = W.shape[0] if tr else W.shape[1]
        v = cases.spikes(15, k, 0.2)
        B = cases.spikes(16, k, 0.2, batch=7)
        yield (f"dense_t{int(tr)}_mv", (lambda api, v=v, tr=tr: api.binary_densemv(W, v, transpose=tr)), 1e-5)
        yield (f"dense_t{int(tr)}_mm", (lambda api, B=B, tr=tr: api.binary_densemm(W, B, transpose=tr)), 1e-5)


def _fvals(seed, n, batch=None, density=0.7):
    rng = np.random.default_rng(seed)
    shape = (n,) if batch is None else (n, batch)
    return (rng.standard_normal(shape) * (rng.random(shape) < density)).astype(np.float32)


def float_op_cases():
    """(tag, fn(api) -> ndarray, atol) for tests/golden/float_ops.npz: the float (non-event)
    operators, generated from the reference's numba kernels (make_golden.py::golden_float_ops)."""
    for homo in (False, True):
        c = cases.csr_case(20, 300, 200, 9, 0.1, homo=homo, ragged=True)
        m, k = c["shape"]
        for tr in (False, True):
            tag = f"csr_h{int(homo)}_t{int(tr)}"
            v = _fvals(21, m if tr else k)
            B = _fvals(22, m if tr else k, batch=13)
            yield (tag + "_mv", (lambda api, c=c, v=v, tr=tr: api.csrmv(
                c["data"], c["indices"], c["indptr"], v, shape=c["shape"], transpose=tr)), 2e-5)
            yield (tag + "_mm", (lambda api, c=c, B=B, tr=tr: api.csrmm(
                c["data"], c["indices"], c["indptr"], B, shape=c["shape"], transpose=tr)), 2e-5)
    shape, prob, seed = (230, 310), 0.05, 4321
    for fam, params in [("s", (1.5,)), ("n", (1.0, 0.1)), ("u", (-0.5, 1.0))]:
        atol = 3e-4 if fam == "n" else 2e-5   # normal: fastmath reference build, see op_cases()
        for tr in (False, True):
            for co in (False, True):
                tag = f"jit{fam}_t{int(tr)}_c{int(co)}"
                vec_len = shape[0] if tr else shape[1]
                v = _fvals(23, vec_len, density=0.6)
                B = _fvals(24, vec_len, batch=6, density=0.6)
                yield (tag + "_mv", (lambda api, fam=fam, params=params, v=v, tr=tr, co=co: getattr(
                    api, f"jit{fam}mv")(*params, prob, v, seed, shape=shape, transpose=tr, corder=co)), atol)
                yield (tag + "_mm", (lambda api, fam=fam, params=params, B=B, tr=tr, co=co: getattr(
                    api, f"jit{fam}mm")(*params, prob, B, seed, shape=shape, transpose=tr, corder=co)), atol)


def to_csr_cases():
    """(tag, fn(api) -> (data, indices, indptr), atol): JITC -> CSR materialisation."""
    shape, prob, seed = (230, 310), 0.05, 2468
    for fam, params in [("s", (1.5,)), ("n", (1.0, 0.1)), ("u", (-0.5, 1.0))]:
        for co in (False, True):
            for mode in ("mv", "mm"):
                yield (f"tocsr_{fam}_c{int(co)}_{mode}", (lambda api, fam=fam, params=params, co=co, mode=mode: getattr(
                    api, f"jit{fam}_to_csr")(*params, prob, seed, shape=shape, corder=co, matrix_mode=mode)),
                    3e-4 if fam == "n" else 1e-6)


def plasticity_cases():
    """(tag, fn(api) -> ndarray): spike-triggered weight updates without clipping (the reference
    kernels do not clip; clipping is a separate whole-array pass)."""
    rng = np.random.default_rng(30)
    c = cases.csr_case(31, 300, 200, 9, 0.1, homo=False, ragged=True)
    m, k = c["shape"]
    tr_pre, tr_post = rng.standard_normal(m).astype(np.float32), rng.standard_normal(k).astype(np.float32)
    from oracle import csr_to_csc_index
    cp, ci, perm = csr_to_csc_index(c["indptr"], c["indices"], shape=c["shape"])
    f = cases.fcn_case(32, 250, 180, 11)
    tr_f = rng.standard_normal(180).astype(np.float32)
    W = rng.standard_normal((90, 130)).astype(np.float32)
    for fs in (False, True):
        pre, post = cases.spikes(33, m, 0.2, fs), cases.spikes(34, k, 0.2, fs)
        yield (f"plast_csr_pre_f{int(fs)}", lambda api, pre=pre: api.update_csr_on_binary_pre(
            c["data"], c["indices"], c["indptr"], pre, tr_post, shape=c["shape"]))
        yield (f"plast_csr_post_f{int(fs)}", lambda api, post=post: api.update_csr_on_binary_post(
            c["data"], ci, cp, perm, tr_pre, post, shape=c["shape"]))
        spk = cases.spikes(35, 250, 0.2, fs)
        yield (f"plast_fcn_f{int(fs)}", lambda api, spk=spk: api.fcn_plasticity_row(f["weights"], f["indices"], spk, tr_f))
        sp, so = cases.spikes(36, 90, 0.2, fs), cases.spikes(37, 130, 0.2, fs)
        yield (f"plast_dense_pre_f{int(fs)}", lambda api, sp=sp: api.update_dense_on_binary_pre(W, sp, tr_post[:130]))
        yield (f"plast_dense_post_f{int(fs)}", lambda api, so=so: api.update_dense_on_binary_post(W, tr_pre[:90], so))


def encoder_cases():
    """(tag, fn(api) -> tuple of ndarrays): dense 2-D spikes -> sparse layouts."""
    for fs in (False, True):
        spk = cases.spikes(40, 300, 0.05, fs, batch=70)
        yield (f"enc_pair_f{int(fs)}", lambda api, spk=spk: tuple(api.binary_2d_pair_stream_encode(spk)))
        yield (f"enc_rowsparse_f{int(fs)}", lambda api, spk=spk: (api.binary_2d_row_sparse_encode(spk, 40),))
        yield (f"enc_csr_f{int(fs)}", lambda api, spk=spk: tuple(api.binary_2d_csr_encode(spk)))
