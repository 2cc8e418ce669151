"""Generate tests/golden/*.npz from the UNMODIFIED reference numba kernels.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Inputs come from tests/cases.py (seeded); outputs are what the reference's own kernels produce
through oracle/ref_shim.py.  The fixtures travel to the GPU box, the reference does not.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_kernels as RK  # noqa: E402
from oracle import conn_length, normalize_chunk_size  # noqa: E402
from tests import cases  # noqa: E402


def save(name, **arrs):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in arrs.items()})


def golden_rng():
    r = RK.rng()
    keys = np.array([(42, 3, 0, 1), (0, 0, 0, 0), (123, 999999, 3, 31), (2**31 - 1, 7, 2, 5),
                     (7, 400000, 1, 17)], np.int64)
    init = np.array([int(r.light_rng_init(*k)) for k in keys], np.uint32)
    cls = np.array([2, 3, 1000, 2000, 10000], np.int64)
    q = np.zeros((len(keys), len(cls)), np.uint32)
    st = np.zeros_like(q)
    for i, s in enumerate(init):
        for j, cl in enumerate(cls):
            a, b = r.light_rng_initial_q(s, int(cl))
            q[i, j], st[i, j] = int(a), int(b)
    stream = np.zeros((len(keys), 64), np.uint32)
    for i, s in enumerate(init):
        for t in range(64):
            s = r.light_rng_next(s)
            stream[i, t] = int(s)
    rng = np.random.default_rng(0)
    hk = rng.integers(0, 2**31 - 1, (512, 3)).astype(np.int64)
    u01 = np.array([r.light_rng_uniform01(*k) for k in hk], np.float32)
    n01 = np.array([r.light_rng_normal01(*k) for k in hk], np.float32)
    save("rng", keys=keys, init=init, cls=cls, q=q, st=st, stream=stream, hash_keys=hk, u01=u01,
         n01=n01)


def golden_connectivity():
    """Index lists of generated rows, recovered from the reference scatter kernel with one-hot
    input vectors (row r active -> nonzeros of the output are exactly row r's connections)."""
    out = {}
    for tag, shape, prob, seed in [("a", (64, 1000), 0.05, 42), ("b", (40, 4097), 0.01, 123)]:
        m, k = shape
        rows = [0, 1, 7, m - 1]
        # mv walk (stride 32): transpose=True, corder=False -> rows = vector entries, walk = shape[1]
        conn = []
        for r_ in rows:
            v = np.zeros(m, bool)
            v[r_] = True
            y = RK.binary_jitsmv(1.0, prob, v, seed, shape=shape, transpose=True, corder=False)
            conn.append(y.astype(np.int32))  # multiplicity per column (duplicates impossible, but keep counts)
        out[f"{tag}_mv"] = np.stack(conn)
        conn = []
        for r_ in rows:
            B = np.zeros((m, 1), bool)
            B[r_, 0] = True
            y = RK.binary_jitsmm(1.0, prob, B, seed, shape=shape, transpose=True, corder=False)
            conn.append(y[:, 0].astype(np.int32))
        out[f"{tag}_mm"] = np.stack(conn)
        out[f"{tag}_meta"] = np.array([m, k, conn_length(prob), seed, normalize_chunk_size(k)] + rows,
                                      np.int64)
    save("jitc_connectivity", **out)


def golden_ops():
    out = {}
    # CSR (ragged, hetero + homo, both directions, bool + float spikes)
    for homo in (False, True):
        c = cases.csr_case(4, 300, 200, 9, 0.1, homo=homo, ragged=True)
        m, k = c["shape"]
        for tr in (False, True):
            for fs in (False, True):
                tag = f"csr_h{int(homo)}_t{int(tr)}_f{int(fs)}"
                v = cases.spikes(5, m if tr else k, 0.15, fs)
                out[tag + "_mv"] = RK.binary_csrmv(c["data"], c["indices"], c["indptr"], v,
                                                  shape=c["shape"], transpose=tr)
                B = cases.spikes(6, m if tr else k, 0.1, fs, batch=13)
                out[tag + "_mm"] = RK.binary_csrmm(c["data"], c["indices"], c["indptr"], B,
                                                  shape=c["shape"], transpose=tr)
    # FCN
    for homo in (False, True):
        c = cases.fcn_case(9, 250, 180, 11, homo=homo)
        for tr in (False, True):
            tag = f"fcn_h{int(homo)}_t{int(tr)}"
            v = cases.spikes(10, 250 if tr else 180, 0.15)
            out[tag + "_mv"] = RK.binary_fcnmv(c["weights"], c["indices"], v, shape=(250, 180),
                                              transpose=tr)
            B = cases.spikes(11, 250 if tr else 180, 0.1, batch=9)
            out[tag + "_mm"] = RK.binary_fcnmm(c["weights"], c["indices"], B, shape=(250, 180),
                                              transpose=tr)
    # JITC
    shape, prob, seed = (230, 310), 0.05, 1234
    for fam, params in [("s", (1.5,)), ("n", (1.0, 0.1)), ("u", (-0.5, 1.0))]:
        for tr in (False, True):
            for co in (False, True):
                tag = f"jit{fam}_t{int(tr)}_c{int(co)}"
                vec_len = shape[0] if tr else shape[1]
                v = cases.spikes(12, vec_len, 0.2)
                out[tag + "_mv"] = getattr(RK, f"binary_jit{fam}mv")(
                    *params, prob, v, seed, shape=shape, transpose=tr, corder=co)
                B = cases.spikes(13, vec_len, 0.2, batch=6)
                out[tag + "_mm"] = getattr(RK, f"binary_jit{fam}mm")(
                    *params, prob, B, seed, shape=shape, transpose=tr, corder=co)
    # dense
    W = np.random.default_rng(14).standard_normal((90, 130)).astype(np.float32)
    for tr in (False, True):
        k = W.shape[0] if tr else W.shape[1]
        out[f"dense_t{int(tr)}_mv"] = RK.binary_densemv(W, cases.spikes(15, k, 0.2), transpose=tr)
        out[f"dense_t{int(tr)}_mm"] = RK.binary_densemm(W, cases.spikes(16, k, 0.2, batch=7),
                                                       transpose=tr)
    # compaction
    s = cases.spikes(1, 1000, 0.1)
    ids, na = RK.binary_1d_array_index(s)
    out["compact1d_ids"], out["compact1d_n"] = ids[:int(na[0])], na
    s2 = cases.spikes(2, 300, 0.02, batch=70)
    p, ids, na = RK.binary_2d_array_index(s2)
    out["compact2d_packed"], out["compact2d_ids"], out["compact2d_n"] = p, ids[:int(na[0])], na
    save("ops", **out)


def golden_float_ops():
    """float (non-event) operators: every case of golden_cases.float_op_cases() through the
    reference's numba kernels."""
    from tests import golden_cases
    out = {tag: fn(RK) for tag, fn, _ in golden_cases.float_op_cases()}
    for tag, fn, _ in golden_cases.to_csr_cases():
        data, indices, indptr = fn(RK)
        out[tag + "_data"], out[tag + "_indices"], out[tag + "_indptr"] = data, indices, indptr
    for tag, fn in golden_cases.plasticity_cases():
        out[tag] = fn(RK)
    for tag, fn in golden_cases.encoder_cases():
        for i, arr in enumerate(fn(RK)):
            out[f"{tag}_{i}"] = arr
    save("float_ops", **out)


if __name__ == "__main__":
    golden_rng()
    golden_connectivity()
    golden_ops()
    golden_float_ops()
