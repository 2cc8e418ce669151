"""CPU oracle for the event-driven synaptic propagation path.  TEST INFRASTRUCTURE ONLY.

numpy-in / numpy-out callables with the reference's operator names and argument meaning
(`binary_csrmv`, `binary_fcnmv`, `binary_jitsmv`, ... -- brainevent/__init__.py) on top of the
C restatement in be_oracle.c.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` leg may import this package; brainevent_b200/ never does.

Parity status: PINNED.  The restatement is checked (a) against the reference's own known-answer
vectors (tests/test_oracle_golden.py: _numba_random_test.py:59-69,80-90, _event/compact_test.py:
182-197, _csr/_test_util.py:257-279, docstring examples) and committed golden fixtures generated
from the unmodified reference numba kernels (tests/golden/make_golden.py), and (b) in the build
container directly against those kernels through oracle/ref_shim.py
(tests/test_oracle_vs_reference.py).
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MV_STRIDE = 32  # _misc.py:37
MM_STRIDE = 4   # _misc.py:38


def build(force: bool = False) -> str:
    """Compile be_oracle.c with gcc (oracle/Makefile).  Returns the .so path."""
    so = os.path.join(_HERE, "libbe_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("be_oracle.c", "be_oracle_ops.inc")]
    stale = (not os.path.exists(so)) or any(
        os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libbe_oracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        L = _LIB
        u32, i64, i32, f32 = ctypes.c_uint32, ctypes.c_int64, ctypes.c_int32, ctypes.c_float
        L.ora_mix32.restype = u32; L.ora_mix32.argtypes = [u32]
        L.ora_rng_bounded.restype = u32; L.ora_rng_bounded.argtypes = [u32, u32]
        L.ora_rng_next.restype = u32; L.ora_rng_next.argtypes = [u32]
        L.ora_rng_init.restype = u32; L.ora_rng_init.argtypes = [u32, u32, u32, u32]
        L.ora_rng_initial_q.restype = u32
        L.ora_rng_initial_q.argtypes = [ctypes.POINTER(u32), u32]
        L.ora_uniform01.restype = f32; L.ora_uniform01.argtypes = [u32, u32, u32]
        L.ora_normal01.restype = f32; L.ora_normal01.argtypes = [u32, u32, u32]
        L.ora_jitc_row_conn.restype = i64
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dtype=None):
    return np.ascontiguousarray(a, dtype=dtype)


def _i64(x):
    return ctypes.c_int64(int(x))


def _u32(x):
    return ctypes.c_uint32(int(x) & 0xFFFFFFFF)


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "_f32"
    if dtype == np.float64:
        return "_f64"
    raise TypeError(f"oracle computes in float32/float64 only, got {dtype}")


def _scalar(dtype, x):
    return (ctypes.c_float if np.dtype(dtype) == np.float32 else ctypes.c_double)(float(x))


# ----------------------------------------------------------------------------------------------
# activity predicates
# ----------------------------------------------------------------------------------------------
def active_mask(spikes, for_compaction: bool = False) -> np.ndarray:
    """bool -> as is; float -> `> 0` (operators) or `!= 0` (compaction)."""
    s = np.asarray(spikes)
    if s.dtype == np.bool_:
        return _c(s.view(np.uint8))
    if s.dtype in (np.int8, np.uint8):
        return _c((s != 0).astype(np.uint8))
    if for_compaction:
        return _c((s != 0).astype(np.uint8))
    return _c((s > 0).astype(np.uint8))


# ----------------------------------------------------------------------------------------------
# light-RNG
# ----------------------------------------------------------------------------------------------
def light_rng_mix32(x):
    return np.uint32(lib().ora_mix32(_u32(x)))


def light_rng_bounded(r, bound):
    return np.uint32(lib().ora_rng_bounded(_u32(r), _u32(bound)))


def light_rng_next(state):
    return np.uint32(lib().ora_rng_next(_u32(state)))


def light_rng_init(seed, row, chunk_id, lane):
    return np.uint32(lib().ora_rng_init(_u32(seed), _u32(row), _u32(chunk_id), _u32(lane)))


def light_rng_initial_q(state, cl):
    st = ctypes.c_uint32(int(state) & 0xFFFFFFFF)
    q = lib().ora_rng_initial_q(ctypes.byref(st), _u32(cl))
    return np.uint32(q), np.uint32(st.value)


def light_rng_uniform01(seed, row, col):
    return np.float32(lib().ora_uniform01(_u32(seed), _u32(row), _u32(col)))


def light_rng_normal01(seed, row, col):
    return np.float32(lib().ora_normal01(_u32(seed), _u32(row), _u32(col)))


def normalize_chunk_size(n_cols, chunk_size=None, target_chunks=4):
    """_misc.py:74-122."""
    if chunk_size is None:
        chunk_size = max(1, (int(n_cols) + target_chunks - 1) // target_chunks)
    chunk_size = int(chunk_size)
    if chunk_size <= 0:
        raise ValueError("chunk_size must be positive")
    return chunk_size


def conn_length(prob) -> int:
    """_data.py:1212-1245: ceil(2/prob) as int32."""
    return int(np.asarray(np.ceil(2.0 / prob), dtype=np.int32))


def jitc_row_connections(seed, clen, row, k, chunk_size, stride=MV_STRIDE):
    """Columns of generated row `row` in walk order (index-for-index parity target)."""
    L = lib()
    cap = 1024
    while True:
        cols = np.empty(cap, np.int64)
        n = L.ora_jitc_row_conn(_u32(seed), _u32(clen), _i64(row), _i64(k), _i64(chunk_size),
                                ctypes.c_int(stride), _p(cols), _i64(cap))
        if n <= cap:
            return cols[:n].copy()
        cap = int(n)


def jitc_to_dense(kind, p0, p1, prob, seed, shape_rows, walk_len, chunk_size, stride=MV_STRIDE,
                  dtype=np.float32):
    """Dense [shape_rows, walk_len] matrix of the generated connectivity in the kernel frame."""
    clen = conn_length(prob)
    M = np.zeros((shape_rows, walk_len), dtype)
    for r in range(shape_rows):
        cols = jitc_row_connections(seed, clen, r, walk_len, chunk_size, stride)
        for j in cols:
            if kind == 0:
                w = p0
            elif kind == 1:
                w = np.float32(p0) + light_rng_normal01(seed, r, j) * np.float32(p1)
            else:
                w = np.float32(p0) + light_rng_uniform01(seed, r, j) * (np.float32(p1) - np.float32(p0))
            M[r, j] += w
    return M


# ----------------------------------------------------------------------------------------------
# bit-pack / compaction
# ----------------------------------------------------------------------------------------------
def bitpack(arr, axis=0):
    """_event/bitpack_binary.py:32-76 (LSB-first uint32 words along `axis`)."""
    a = np.asarray(arr).astype(bool)
    a = np.moveaxis(a, axis, -1)
    n = a.shape[-1]
    nw = (n + 31) // 32
    flat = _c(a.reshape(-1, n).view(np.uint8))
    out = np.empty((flat.shape[0], nw), np.uint32)
    for i in range(flat.shape[0]):
        lib().ora_bitpack_1d(_p(flat[i]), _i64(n), _p(out[i]))
    out = out.reshape(a.shape[:-1] + (nw,))
    return np.moveaxis(out, -1, axis)


def binary_1d_array_index(spikes):
    """_event/compact.py:305-335 -> (active_ids int32[n] ascending + zero tail, n_active int32[1])."""
    act = active_mask(spikes, for_compaction=True)
    n = act.shape[0]
    ids = np.zeros(n, np.int32)
    na = np.zeros(1, np.int32)
    lib().ora_compact_1d(_p(act), _i64(n), _p(ids), _p(na))
    return ids, na


def binary_2d_array_index(spikes):
    """_event/compact.py:451-503 -> (packed uint32[n_pre, ceil(b/32)], active_ids, n_active)."""
    act = active_mask(spikes, for_compaction=True)
    n_pre, nb = act.shape
    packed = np.zeros((n_pre, (nb + 31) // 32), np.uint32)
    ids = np.zeros(n_pre, np.int32)
    na = np.zeros(1, np.int32)
    lib().ora_compact_2d(_p(act), _i64(n_pre), _i64(nb), _p(packed), _p(ids), _p(na))
    return packed, ids, na


def binary_2d_compact_only(spikes):
    """_event/compact.py:158-199."""
    act = active_mask(spikes, for_compaction=True)
    n_pre, nb = act.shape
    ids = np.zeros(n_pre, np.int32)
    na = np.zeros(1, np.int32)
    lib().ora_compact_2d(_p(act), _i64(n_pre), _i64(nb), None, _p(ids), _p(na))
    return ids, na


def csr_to_csc_index(indptr, indices, shape):
    """_misc.py:1341-1384 -> (csc_indptr int64[n_post+1], csc_indices int32[nse], perm int64[nse])."""
    m, n_post = shape
    indptr = _c(indptr, np.int64)
    indices = _c(indices, np.int32)
    nse = int(indptr[m])
    ci = np.zeros(n_post + 1, np.int64)
    cx = np.zeros(nse, np.int32)
    pm = np.zeros(nse, np.int64)
    lib().ora_csr_to_csc(_p(indptr), _p(indices), _i64(m), _i64(n_post), _p(ci), _p(cx), _p(pm))
    return ci, cx, pm


# ----------------------------------------------------------------------------------------------
# CSR
# ----------------------------------------------------------------------------------------------
def _csr_common(data, indices, indptr, perm):
    data = _c(np.atleast_1d(data))
    homo = int(data.size == 1)
    return (data, homo, _c(indices, np.int32), _c(indptr, np.int64),
            None if perm is None else _c(perm, np.int64))


def binary_csrmv(data, indices, indptr, v, *, shape, transpose=False, perm=None):
    """_csr/binary.py:128 (and binary_indexed.py:70 when perm is given)."""
    data, homo, indices, indptr, perm = _csr_common(data, indices, indptr, perm)
    m, k = shape
    act = active_mask(v)
    assert act.shape[0] == (m if transpose else k)
    out = np.zeros(k if transpose else m, data.dtype)
    getattr(lib(), "ora_csrmv" + _sfx(data.dtype))(
        _p(data), homo, _p(indices), _p(indptr), _p(perm), _p(act), _i64(m), _i64(k),
        int(bool(transpose)), _p(out))
    return out


def binary_csrmm(data, indices, indptr, B, *, shape, transpose=False, perm=None):
    """_csr/binary.py:264."""
    data, homo, indices, indptr, perm = _csr_common(data, indices, indptr, perm)
    m, k = shape
    act = active_mask(B)
    n = act.shape[1]
    assert act.shape[0] == (m if transpose else k)
    out = np.zeros((k if transpose else m, n), data.dtype)
    getattr(lib(), "ora_csrmm" + _sfx(data.dtype))(
        _p(data), homo, _p(indices), _p(indptr), _p(perm), _p(act), _i64(m), _i64(k), _i64(n),
        int(bool(transpose)), _p(out))
    return out


def binary_csrmv_indexed(data, indices, indptr, perm, v, *, shape, transpose=False):
    return binary_csrmv(data, indices, indptr, v, shape=shape, transpose=transpose, perm=perm)


def binary_csrmm_indexed(data, indices, indptr, perm, B, *, shape, transpose=False):
    return binary_csrmm(data, indices, indptr, B, shape=shape, transpose=transpose, perm=perm)


# ----------------------------------------------------------------------------------------------
# FCN
# ----------------------------------------------------------------------------------------------
def binary_fcnmv(weights, indices, spikes, *, shape, transpose=False):
    """_fcn/binary.py:43."""
    weights = _c(np.atleast_1d(weights))
    homo = int(weights.size == 1)
    indices = _c(indices, np.int32)
    n_rows, n_conn = indices.shape
    act = active_mask(spikes)
    n_out = shape[1] if transpose else shape[0]
    out = np.zeros(n_out, weights.dtype)
    getattr(lib(), "ora_fcnmv" + _sfx(weights.dtype))(
        _p(weights), homo, _p(indices), _i64(n_rows), _i64(n_conn), _p(act), _i64(n_out),
        int(bool(transpose)), _p(out))
    return out


def binary_fcnmm(weights, indices, matrix, *, shape, transpose=False):
    """_fcn/binary.py:512."""
    weights = _c(np.atleast_1d(weights))
    homo = int(weights.size == 1)
    indices = _c(indices, np.int32)
    n_rows, n_conn = indices.shape
    act = active_mask(matrix)
    n = act.shape[1]
    n_out = shape[1] if transpose else shape[0]
    out = np.zeros((n_out, n), weights.dtype)
    getattr(lib(), "ora_fcnmm" + _sfx(weights.dtype))(
        _p(weights), homo, _p(indices), _i64(n_rows), _i64(n_conn), _p(act), _i64(n_out), _i64(n),
        int(bool(transpose)), _p(out))
    return out


# ----------------------------------------------------------------------------------------------
# JITC
# ----------------------------------------------------------------------------------------------
def _jit_mv(kind, p0, p1, prob, vector, seed, shape, transpose, corder, dtype, rows=None):
    clen = conn_length(prob)
    act = active_mask(vector)
    out_len = shape[1] if transpose else shape[0]
    vec_len = shape[0] if transpose else shape[1]
    assert act.shape[0] == vec_len
    chunk = normalize_chunk_size(shape[1])
    r0, rn = (0, out_len) if rows is None else rows
    assert rows is None or corder, "row subsampling is defined for the gather orientation only"
    out = np.zeros(rn if corder else out_len, dtype)
    getattr(lib(), "ora_jitmv" + _sfx(dtype))(
        kind, _scalar(dtype, p0), _scalar(dtype, p1), _u32(clen), _u32(seed), _p(act),
        _i64(vec_len), _i64(out_len), _i64(chunk), int(bool(corder)), _i64(r0), _i64(rn), _p(out))
    return out


def _jit_mm(kind, p0, p1, prob, B, seed, shape, transpose, corder, dtype):
    clen = conn_length(prob)
    act = active_mask(B)
    n = act.shape[1]
    out_len = shape[1] if transpose else shape[0]
    vec_len = shape[0] if transpose else shape[1]
    assert act.shape[0] == vec_len
    chunk = normalize_chunk_size(shape[1])
    out = np.zeros((out_len, n), dtype)
    getattr(lib(), "ora_jitmm" + _sfx(dtype))(
        kind, _scalar(dtype, p0), _scalar(dtype, p1), _u32(clen), _u32(seed), _p(act),
        _i64(vec_len), _i64(out_len), _i64(n), _i64(chunk), int(bool(corder)), _p(out))
    return out


def binary_jitsmv(weight, prob, vector, seed, *, shape, transpose=False, corder=True,
                  dtype=np.float32, rows=None):
    """_jit_scalar/binary.py:44.  rows=(offset, count): gather orientation, row subsample."""
    return _jit_mv(0, weight, 0.0, prob, vector, seed, shape, transpose, corder, dtype, rows)


def binary_jitsmm(weight, prob, B, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    """_jit_scalar/binary.py:171."""
    return _jit_mm(0, weight, 0.0, prob, B, seed, shape, transpose, corder, dtype)


def binary_jitnmv(w_loc, w_scale, prob, vector, seed, *, shape, transpose=False, corder=True,
                  dtype=np.float32, rows=None):
    """_jit_normal/binary.py:44."""
    return _jit_mv(1, w_loc, w_scale, prob, vector, seed, shape, transpose, corder, dtype, rows)


def binary_jitnmm(w_loc, w_scale, prob, B, seed, *, shape, transpose=False, corder=True,
                  dtype=np.float32):
    """_jit_normal/binary.py:177."""
    return _jit_mm(1, w_loc, w_scale, prob, B, seed, shape, transpose, corder, dtype)


def binary_jitumv(w_low, w_high, prob, vector, seed, *, shape, transpose=False, corder=True,
                  dtype=np.float32, rows=None):
    """_jit_uniform/binary.py:44."""
    return _jit_mv(2, w_low, w_high, prob, vector, seed, shape, transpose, corder, dtype, rows)


def binary_jitumm(w_low, w_high, prob, B, seed, *, shape, transpose=False, corder=True,
                  dtype=np.float32):
    """_jit_uniform/binary.py:177."""
    return _jit_mm(2, w_low, w_high, prob, B, seed, shape, transpose, corder, dtype)


# ----------------------------------------------------------------------------------------------
# dense
# ----------------------------------------------------------------------------------------------
def binary_densemv(weights, spikes, *, transpose):
    """_dense/binary.py:79."""
    W = _c(weights)
    rows, cols = W.shape
    act = active_mask(spikes)
    assert act.shape[0] == (rows if transpose else cols)
    out = np.zeros(cols if transpose else rows, W.dtype)
    getattr(lib(), "ora_densemv" + _sfx(W.dtype))(
        _p(W), _i64(rows), _i64(cols), _p(act), int(bool(transpose)), _p(out))
    return out


def binary_densemm(weights, spikes, *, transpose):
    """_dense/binary.py:487: weights[k,m].T @ spikes[k,n] (transpose) or weights[m,k] @ spikes[k,n]."""
    W = _c(weights)
    rows, cols = W.shape
    act = active_mask(spikes)
    n = act.shape[1]
    assert act.shape[0] == (rows if transpose else cols)
    out = np.zeros((cols if transpose else rows, n), W.dtype)
    getattr(lib(), "ora_densemm" + _sfx(W.dtype))(
        _p(W), _i64(rows), _i64(cols), _p(act), _i64(n), int(bool(transpose)), _p(out))
    return out


# ----------------------------------------------------------------------------------------------
# float (non-event) operators -- SURVEY 8(f)-1
# ----------------------------------------------------------------------------------------------
def csrmv(data, indices, indptr, v, *, shape, transpose=False):
    """_csr/float.py:49."""
    data, homo, indices, indptr, _ = _csr_common(data, indices, indptr, None)
    m, k = shape
    x = _c(v, data.dtype)
    assert x.shape[0] == (m if transpose else k)
    out = np.zeros(k if transpose else m, data.dtype)
    getattr(lib(), "ora_csrmv_val" + _sfx(data.dtype))(
        _p(data), homo, _p(indices), _p(indptr), _p(x), _i64(m), _i64(k), int(bool(transpose)), _p(out))
    return out


def csrmm(data, indices, indptr, B, *, shape, transpose=False):
    """_csr/float.py:559."""
    data, homo, indices, indptr, _ = _csr_common(data, indices, indptr, None)
    m, k = shape
    X = _c(B, data.dtype)
    n = X.shape[1]
    assert X.shape[0] == (m if transpose else k)
    out = np.zeros((k if transpose else m, n), data.dtype)
    getattr(lib(), "ora_csrmm_val" + _sfx(data.dtype))(
        _p(data), homo, _p(indices), _p(indptr), _p(X), _i64(m), _i64(k), _i64(n), int(bool(transpose)), _p(out))
    return out


def _fcn_as_csr(weights, indices):
    weights = _c(np.atleast_1d(weights))
    indices = _c(indices, np.int32)
    n_rows, n_conn = indices.shape
    indptr = np.arange(n_rows + 1, dtype=np.int64) * n_conn
    return (weights if weights.size == 1 else weights.reshape(-1)), indices.reshape(-1), indptr


def fcnmv(weights, indices, vector, *, shape, transpose=False):
    """_fcn/float.py:33 (ELL == CSR with a regular indptr)."""
    w, ix, ip = _fcn_as_csr(weights, indices)
    return csrmv(w, ix, ip, vector, shape=shape, transpose=transpose)


def fcnmm(weights, indices, matrix, *, shape, transpose=False):
    """_fcn/float.py:136."""
    w, ix, ip = _fcn_as_csr(weights, indices)
    return csrmm(w, ix, ip, matrix, shape=shape, transpose=transpose)


def _jit_mv_val(kind, p0, p1, prob, vector, seed, shape, transpose, corder, dtype):
    clen = conn_length(prob)
    x = _c(vector, dtype)
    out_len = shape[1] if transpose else shape[0]
    vec_len = shape[0] if transpose else shape[1]
    assert x.shape[0] == vec_len
    chunk = normalize_chunk_size(shape[1])
    out = np.zeros(out_len, dtype)
    getattr(lib(), "ora_jitmv_val" + _sfx(dtype))(
        kind, _scalar(dtype, p0), _scalar(dtype, p1), _u32(clen), _u32(seed), _p(x),
        _i64(vec_len), _i64(out_len), _i64(chunk), int(bool(corder)), _p(out))
    return out


def jitsmv(weight, prob, vector, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    """_jit_scalar/float.py:159."""
    return _jit_mv_val(0, weight, 0.0, prob, vector, seed, shape, transpose, corder, dtype)


def jitnmv(w_loc, w_scale, prob, vector, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    """_jit_normal/float.py (jitnmv)."""
    return _jit_mv_val(1, w_loc, w_scale, prob, vector, seed, shape, transpose, corder, dtype)


def jitumv(w_low, w_high, prob, vector, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    """_jit_uniform/float.py (jitumv)."""
    return _jit_mv_val(2, w_low, w_high, prob, vector, seed, shape, transpose, corder, dtype)


def _jit_mm_val(kind, p0, p1, prob, B, seed, shape, transpose, corder, dtype):
    clen = conn_length(prob)
    X = _c(B, dtype)
    n = X.shape[1]
    out_len = shape[1] if transpose else shape[0]
    vec_len = shape[0] if transpose else shape[1]
    assert X.shape[0] == vec_len
    chunk = normalize_chunk_size(shape[1])
    out = np.zeros((out_len, n), dtype)
    getattr(lib(), "ora_jitmm_val" + _sfx(dtype))(
        kind, _scalar(dtype, p0), _scalar(dtype, p1), _u32(clen), _u32(seed), _p(X),
        _i64(vec_len), _i64(out_len), _i64(n), _i64(chunk), int(bool(corder)), _p(out))
    return out


def jitsmm(weight, prob, B, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    """_jit_scalar/float.py:1298."""
    return _jit_mm_val(0, weight, 0.0, prob, B, seed, shape, transpose, corder, dtype)


def jitnmm(w_loc, w_scale, prob, B, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    return _jit_mm_val(1, w_loc, w_scale, prob, B, seed, shape, transpose, corder, dtype)


def jitumm(w_low, w_high, prob, B, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    return _jit_mm_val(2, w_low, w_high, prob, B, seed, shape, transpose, corder, dtype)


# ----------------------------------------------------------------------------------------------
# JITC -> CSR materialisation -- SURVEY 8(f)-2
# ----------------------------------------------------------------------------------------------
def _jit_to_csr(kind, p0, p1, prob, seed, shape, corder, matrix_mode, dtype):
    """(data, indices, indptr) column-sorted, like jits_to_csr (_jit_scalar/csr.py:488-570)."""
    n_rows, n_cols = int(shape[0]), int(shape[1])
    stride = {"mv": 32, "mm": 4}[matrix_mode]
    clen = conn_length(prob)
    chunk = normalize_chunk_size(n_cols)
    counts = np.zeros(n_rows, np.int64)
    if n_rows and n_cols:
        lib().ora_jit_csr_count(_u32(clen), _u32(seed), _i64(n_rows), _i64(n_cols), _i64(chunk), stride,
                                int(bool(corder)), _p(counts))
    indptr = np.zeros(n_rows + 1, np.int64)
    np.cumsum(counts, out=indptr[1:])
    nnz = int(indptr[-1])
    indices = np.zeros(nnz, np.int32)
    data = np.zeros(nnz, dtype)
    if nnz:
        cursor = np.zeros(n_rows, np.int64)
        getattr(lib(), "ora_jit_csr_fill" + _sfx(dtype))(
            kind, _scalar(dtype, p0), _scalar(dtype, p1), _u32(clen), _u32(seed), _i64(n_rows), _i64(n_cols),
            _i64(chunk), stride, int(bool(corder)), _p(indptr), _p(cursor), _p(indices), _p(data))
        seg = np.repeat(np.arange(n_rows, dtype=np.int64), counts)
        order = np.lexsort((indices, seg))
        indices, data = indices[order], data[order]
    return data, indices, indptr


def jits_to_csr(weight, prob, seed, *, shape, corder, matrix_mode, dtype=np.float32):
    """_jit_scalar/csr.py:488."""
    return _jit_to_csr(0, weight, 0.0, prob, seed, shape, corder, matrix_mode, dtype)


def jitn_to_csr(w_loc, w_scale, prob, seed, *, shape, corder, matrix_mode, dtype=np.float32):
    """_jit_normal/csr.py:465."""
    return _jit_to_csr(1, w_loc, w_scale, prob, seed, shape, corder, matrix_mode, dtype)


def jitu_to_csr(w_low, w_high, prob, seed, *, shape, corder, matrix_mode, dtype=np.float32):
    """_jit_uniform/csr.py."""
    return _jit_to_csr(2, w_low, w_high, prob, seed, shape, corder, matrix_mode, dtype)


# ----------------------------------------------------------------------------------------------
# spike-triggered additive plasticity -- SURVEY 8(f)-4 (plain numpy restatements)
# ----------------------------------------------------------------------------------------------
def _clip(w, w_min, w_max):
    if w_min is None and w_max is None:
        return w
    return np.clip(w, w_min, w_max).astype(w.dtype)


def update_csr_on_binary_pre(weight, indices, indptr, pre_spike, post_trace, w_min=None, w_max=None, *, shape):
    """_csr/plasticity_binary.py:176-204: out_w[i_start:i_end] += post_trace[indices[i_start:i_end]]."""
    w = np.array(weight, copy=True)
    act = active_mask(pre_spike, True)
    for i in np.nonzero(act)[0]:
        s, e = int(indptr[i]), int(indptr[i + 1])
        w[s:e] += np.asarray(post_trace, w.dtype)[indices[s:e]]
    return _clip(w, w_min, w_max)


def update_csr_on_binary_post(weight, indices, indptr, weight_indices, pre_trace, post_spike, w_min=None, w_max=None,
                              *, shape):
    """_csr/plasticity_binary.py:621-654: out_w[weight_indices[seg]] += pre_trace[indices[seg]]."""
    w = np.array(weight, copy=True)
    act = active_mask(post_spike, True)
    for i in np.nonzero(act)[0]:
        s, e = int(indptr[i]), int(indptr[i + 1])
        w[weight_indices[s:e]] += np.asarray(pre_trace, w.dtype)[indices[s:e]]
    return _clip(w, w_min, w_max)


def fcn_plasticity_row(data, indices, spike, trace, w_min=None, w_max=None):
    """_fcn/plasticity_binary.py:81-109: out[r, k] += trace[indices[r, k]] for firing rows r."""
    w = np.array(data, copy=True)
    act = active_mask(spike, True).astype(bool)
    w[act] += np.asarray(trace, w.dtype)[indices[act]]
    return _clip(w, w_min, w_max)


def update_dense_on_binary_pre(weight, pre_spike, post_trace, w_min=None, w_max=None):
    """_dense/plasticity_binary.py:144-164."""
    w = np.array(weight, copy=True)
    w[active_mask(pre_spike, True).astype(bool)] += np.asarray(post_trace, w.dtype)[None, :]
    return _clip(w, w_min, w_max)


def update_dense_on_binary_post(weight, pre_trace, post_spike, w_min=None, w_max=None):
    """_dense/plasticity_binary.py:462-482."""
    w = np.array(weight, copy=True)
    w[:, active_mask(post_spike, True).astype(bool)] += np.asarray(pre_trace, w.dtype)[:, None]
    return _clip(w, w_min, w_max)


# ----------------------------------------------------------------------------------------------
# dense 2-D spikes -> sparse event layouts -- SURVEY 8(f)-3 (plain numpy restatements)
# ----------------------------------------------------------------------------------------------
def binary_2d_pair_stream_encode(spikes):
    """_event/compact.py:634-670 -> (pair_stream int32 [n*b, 2], n_pairs int32 [1]); row-major order."""
    act = active_mask(spikes, True).astype(bool)
    r, c = np.nonzero(act)
    out = np.zeros((act.size, 2), np.int32)
    out[:r.size, 0], out[:r.size, 1] = r, c
    return out, np.asarray([r.size], np.int32)


def binary_2d_row_sparse_encode(spikes, row_size):
    """_event/compact.py:792-826 -> int32 [n_src, row_size], 1-based column ids, zero padded."""
    act = active_mask(spikes, True).astype(bool)
    out = np.zeros((act.shape[0], row_size), np.int32)
    for r in range(act.shape[0]):
        cols = np.nonzero(act[r])[0]
        out[r, :cols.size] = cols + 1
    return out


def binary_2d_csr_encode(spikes):
    """_event/compact.py:962-1217 -> (indices int32 [n*b] zero-filled tail, indptr int32 [n_src+1])."""
    act = active_mask(spikes, True).astype(bool)
    r, c = np.nonzero(act)
    indices = np.zeros(act.size, np.int32)
    indices[:c.size] = c
    indptr = np.zeros(act.shape[0] + 1, np.int32)
    np.cumsum(act.sum(axis=1), out=indptr[1:])
    return indices, indptr


def binary_2d_csc_encode(spikes):
    """_event/compact.py:1220-1262 -> (row ids int32 [n*b] zero-filled tail, indptr int32 [n_batch+1])."""
    act = active_mask(spikes, True).astype(bool)
    c, r = np.nonzero(act.T)
    indices = np.zeros(act.size, np.int32)
    indices[:r.size] = r
    indptr = np.zeros(act.shape[1] + 1, np.int32)
    np.cumsum(act.sum(axis=0), out=indptr[1:])
    return indices, indptr
