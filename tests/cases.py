"""Seeded synthetic inputs shared by the oracle, golden and GPU parity tests."""
import numpy as np


def csr_case(seed, m, k, nnz_per_row, rate, homo=False, dtype=np.float32, ragged=False,
             float_spikes=False, vec_len=None):
    """Random CSR [m, k]; returns dict(data, indices, indptr, shape)."""
    rng = np.random.default_rng(seed)
    if ragged:
        counts = rng.integers(0, 2 * nnz_per_row + 1, m)
        # a few heavy rows and a few empty ones
        if m > 4:
            counts[rng.integers(0, m)] = min(k * 2, 40 * nnz_per_row)
            counts[rng.integers(0, m)] = 0
    else:
        counts = np.full(m, nnz_per_row)
    indptr = np.zeros(m + 1, np.int32)
    np.cumsum(counts, out=indptr[1:])
    nse = int(indptr[-1])
    indices = rng.integers(0, k, nse).astype(np.int32)
    data = (np.asarray([rng.standard_normal()], dtype) if homo
            else rng.standard_normal(nse).astype(dtype))
    return dict(data=data, indices=indices, indptr=indptr, shape=(m, k))


def spikes(seed, n, rate, float_spikes=False, batch=None):
    rng = np.random.default_rng(seed + 7919)
    shape = (n,) if batch is None else (n, batch)
    s = rng.random(shape) < rate
    if float_spikes:
        # positive, zero and negative entries: only > 0 is active for the operators
        vals = rng.standard_normal(shape).astype(np.float32)
        return np.where(s, np.abs(vals) + 0.1, np.where(rng.random(shape) < 0.2, -1.0, 0.0)
                        ).astype(np.float32)
    return s


def fcn_case(seed, n_rows, n_post, n_conn, homo=False, dtype=np.float32):
    rng = np.random.default_rng(seed)
    indices = rng.integers(0, n_post, (n_rows, n_conn)).astype(np.int32)
    w = (np.asarray([rng.standard_normal()], dtype) if homo
         else rng.standard_normal((n_rows, n_conn)).astype(dtype))
    return dict(weights=w, indices=indices)
