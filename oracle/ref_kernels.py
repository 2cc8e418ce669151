"""numpy-in / numpy-out adapters over the UNMODIFIED reference numba kernel generators.

TEST INFRASTRUCTURE ONLY; needs /root/reference (build container).  Each adapter passes the same
static kwargs the reference primitive would pass and has the same
signature as its counterpart in oracle/__init__.py, so tests can compare them call-for-call.
"""
import numpy as np

from . import ref_shim as R
from . import conn_length, normalize_chunk_size

SDS = R.SDS
available = R.available


def _first(o):
    return o[0] if isinstance(o, (list, tuple)) else o


def _spk(v):
    v = np.ascontiguousarray(v)
    return v


# -- RNG ----------------------------------------------------------------------------------------
def rng():
    return R.ref_module("_numba_random")


# -- compaction / bitpack -----------------------------------------------------------------------
def binary_1d_array_index(spikes):
    m = R.ref_module("_event.compact")
    s = _spk(spikes)
    n = s.shape[0]
    k = m._binary_1d_array_index_numba_kernel(
        spikes_info=SDS(s.shape, s.dtype), outs=[SDS((n,), np.int32), SDS((1,), np.int32)])
    ids, na = k(s)
    return ids, na


def binary_2d_array_index(spikes):
    m = R.ref_module("_event.compact")
    s = _spk(spikes)
    n, b = s.shape
    k = m._binary_2d_array_index_numba_kernel(
        spikes_info=SDS(s.shape, s.dtype),
        outs=[SDS((n, (b + 31) // 32), np.uint32), SDS((n,), np.int32), SDS((1,), np.int32)])
    return tuple(k(s))


def binary_2d_compact_only(spikes):
    m = R.ref_module("_event.compact")
    s = _spk(spikes)
    n, b = s.shape
    k = m._binary_2d_compact_only_numba_kernel(
        spikes_info=SDS(s.shape, s.dtype), outs=[SDS((n,), np.int32), SDS((1,), np.int32)])
    return tuple(k(s))


def csr_to_csc_index(indptr, indices, shape):
    m = R.ref_module("_misc")
    return m._csr_to_csc_index_numpy(np.asarray(indptr), np.asarray(indices), shape=shape)


# -- CSR ------------------------------------------------------------------------------------------
def binary_csrmv(data, indices, indptr, v, *, shape, transpose=False):
    m = R.ref_module("_csr.binary")
    data = np.atleast_1d(data)
    v = _spk(v)
    out_n = shape[1] if transpose else shape[0]
    k = m._csrmv_numba_kernel(weight_info=SDS(data.shape, data.dtype),
                              vector_info=SDS(v.shape, v.dtype), transpose=transpose,
                              outs=[SDS((out_n,), data.dtype)])
    return _first(k(data, indices, indptr, v, None, None, None)[0])


def binary_csrmm(data, indices, indptr, B, *, shape, transpose=False):
    m = R.ref_module("_csr.binary")
    data = np.atleast_1d(data)
    B = _spk(B)
    out_n = shape[1] if transpose else shape[0]
    k = m._csrmm_numba_kernel(weight_info=SDS(data.shape, data.dtype),
                              vector_info=SDS(B.shape, B.dtype), transpose=transpose,
                              outs=[SDS((out_n, B.shape[1]), data.dtype)])
    return _first(k(data, indices, indptr, B, None, None, None)[0])


def binary_csrmv_indexed(data, indices, indptr, perm, v, *, shape, transpose=False):
    m = R.ref_module("_csr.binary_indexed")
    data = np.atleast_1d(data)
    v = _spk(v)
    out_n = shape[1] if transpose else shape[0]
    k = m._csrmv_indexed_numba_kernel(weight_info=SDS(data.shape, data.dtype),
                                      vector_info=SDS(v.shape, v.dtype), transpose=transpose,
                                      outs=[SDS((out_n,), data.dtype)])
    return _first(k(data, indices, indptr, perm, v, None, None, None)[0])


def binary_csrmm_indexed(data, indices, indptr, perm, B, *, shape, transpose=False):
    m = R.ref_module("_csr.binary_indexed")
    data = np.atleast_1d(data)
    B = _spk(B)
    out_n = shape[1] if transpose else shape[0]
    k = m._csrmm_indexed_numba_kernel(weight_info=SDS(data.shape, data.dtype),
                                      vector_info=SDS(B.shape, B.dtype), transpose=transpose,
                                      outs=[SDS((out_n, B.shape[1]), data.dtype)])
    return _first(k(data, indices, indptr, perm, B, None, None, None)[0])


# -- FCN ------------------------------------------------------------------------------------------
def binary_fcnmv(weights, indices, spikes, *, shape, transpose=False):
    m = R.ref_module("_fcn.binary")
    weights = np.atleast_1d(weights)
    s = _spk(spikes)
    out_n = shape[1] if transpose else shape[0]
    k = m._ell_binary_matvec_numba_kernel(weight_info=SDS(weights.shape, weights.dtype),
                                          spike_info=SDS(s.shape, s.dtype), transpose=transpose,
                                          outs=[SDS((out_n,), weights.dtype)])
    return _first(k(weights, indices, s))


def binary_fcnmm(weights, indices, matrix, *, shape, transpose=False):
    m = R.ref_module("_fcn.binary")
    weights = np.atleast_1d(weights)
    s = _spk(matrix)
    out_n = shape[1] if transpose else shape[0]
    k = m._binary_fcnmm_numba_kernel(weight_info=SDS(weights.shape, weights.dtype),
                                     matrix_info=SDS(s.shape, s.dtype), transpose=transpose,
                                     outs=[SDS((out_n, s.shape[1]), weights.dtype)])
    return _first(k(weights, indices, s))


# -- JITC -----------------------------------------------------------------------------------------
def _jit(modname, gen, is_mm, params, prob, x, seed, shape, transpose, corder, dtype):
    m = R.ref_module(modname)
    x = _spk(x)
    out_n = shape[1] if transpose else shape[0]
    oshape = (out_n, x.shape[1]) if is_mm else (out_n,)
    info_kw = {"B_info" if is_mm else "vector_info": SDS(x.shape, x.dtype)}
    k = getattr(m, gen)(corder=corder, shape=tuple(shape), outs=[SDS(oshape, dtype)],
                        transpose=transpose, **info_kw)
    clen = np.asarray([conn_length(prob)], np.int32)
    seed = np.asarray([seed], np.int32)
    ps = [np.asarray([p], dtype) for p in params]
    if modname.startswith("_jit_scalar"):
        return _first(k(*ps, clen, x, seed, None))
    return _first(k(*ps, clen, x, seed))


def binary_jitsmv(weight, prob, vector, seed, *, shape, transpose=False, corder=True,
                  dtype=np.float32):
    return _jit("_jit_scalar.binary", "_jitsmv_numba_kernel", False, [weight], prob, vector, seed,
                shape, transpose, corder, dtype)


def binary_jitsmm(weight, prob, B, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    return _jit("_jit_scalar.binary", "_jitsmm_numba_kernel", True, [weight], prob, B, seed,
                shape, transpose, corder, dtype)


def binary_jitnmv(w_loc, w_scale, prob, vector, seed, *, shape, transpose=False, corder=True,
                  dtype=np.float32):
    return _jit("_jit_normal.binary", "_jitc_mv_normal_numba_kernel_generator", False,
                [w_loc, w_scale], prob, vector, seed, shape, transpose, corder, dtype)


def binary_jitnmm(w_loc, w_scale, prob, B, seed, *, shape, transpose=False, corder=True,
                  dtype=np.float32):
    return _jit("_jit_normal.binary", "_jitc_mm_normal_numba_kernel_generator", True,
                [w_loc, w_scale], prob, B, seed, shape, transpose, corder, dtype)


def binary_jitumv(w_low, w_high, prob, vector, seed, *, shape, transpose=False, corder=True,
                  dtype=np.float32):
    return _jit("_jit_uniform.binary", "_jitumv_numba_kernel_generator", False,
                [w_low, w_high], prob, vector, seed, shape, transpose, corder, dtype)


def binary_jitumm(w_low, w_high, prob, B, seed, *, shape, transpose=False, corder=True,
                  dtype=np.float32):
    return _jit("_jit_uniform.binary", "_jitumm_numba_kernel_generator", True,
                [w_low, w_high], prob, B, seed, shape, transpose, corder, dtype)


# -- dense ----------------------------------------------------------------------------------------
def binary_densemv(weights, spikes, *, transpose):
    m = R.ref_module("_dense.binary")
    W = np.ascontiguousarray(weights)
    s = _spk(spikes)
    out_n = W.shape[1] if transpose else W.shape[0]
    k = m._binary_densemv_numba_kernel(spk_info=SDS(s.shape, s.dtype), transpose=transpose,
                                       outs=[SDS((out_n,), W.dtype)])
    return _first(k(W, s))


def binary_densemm(weights, spikes, *, transpose):
    m = R.ref_module("_dense.binary")
    W = np.ascontiguousarray(weights)
    s = _spk(spikes)
    out_m = W.shape[1] if transpose else W.shape[0]
    k = m._binary_densemm_numba_kernel(spk_info=SDS(s.shape, s.dtype),
                                       weight_info=SDS(W.shape, W.dtype), transpose=transpose,
                                       outs=[SDS((out_m, s.shape[1]), W.dtype)])
    return _first(k(W, s))


# -- float (non-event) operators ------------------------------------------------------------------
def csrmv(data, indices, indptr, v, *, shape, transpose=False):
    m = R.ref_module("_csr.float")
    data = np.atleast_1d(data)
    v = np.ascontiguousarray(v)
    out_n = shape[1] if transpose else shape[0]
    k = m._csrmv_numba_kernel_generator(weight_info=SDS(data.shape, data.dtype), transpose=transpose,
                                        outs=[SDS((out_n,), data.dtype)])
    return _first(k(data, indices, indptr, v))


def csrmm(data, indices, indptr, B, *, shape, transpose=False):
    m = R.ref_module("_csr.float")
    data = np.atleast_1d(data)
    B = np.ascontiguousarray(B)
    out_n = shape[1] if transpose else shape[0]
    k = m._csrmm_numba_kernel_generator(weight_info=SDS(data.shape, data.dtype), transpose=transpose,
                                        outs=[SDS((out_n, B.shape[1]), data.dtype)])
    return _first(k(data, indices, indptr, B))


def _jit_float(modname, gen, params, prob, x, seed, shape, transpose, corder, dtype):
    m = R.ref_module(modname)
    x = np.ascontiguousarray(x)
    out_n = shape[1] if transpose else shape[0]
    is_mm = x.ndim == 2
    oshape = (out_n, x.shape[1]) if is_mm else (out_n,)
    k = getattr(m, gen)(corder=corder, shape=tuple(shape), outs=[SDS(oshape, dtype)], transpose=transpose,
                        **{"B_info" if is_mm else "vector_info": SDS(x.shape, x.dtype)})
    clen = np.asarray([conn_length(prob)], np.int32)
    seed = np.asarray([seed], np.int32)
    ps = [np.asarray([p], dtype) for p in params]
    if modname.startswith("_jit_scalar"):
        return _first(k(*ps, clen, x, seed, None))
    return _first(k(*ps, clen, x, seed))


def jitsmv(weight, prob, vector, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    return _jit_float("_jit_scalar.float", "_jitsmv_numba_kernel", [weight], prob, vector, seed, shape, transpose, corder, dtype)


def jitnmv(w_loc, w_scale, prob, vector, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    return _jit_float("_jit_normal.float", "_jitnmv_numba_kernel_generator", [w_loc, w_scale], prob, vector, seed, shape,
                      transpose, corder, dtype)


def jitumv(w_low, w_high, prob, vector, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    return _jit_float("_jit_uniform.float", "_jitumv_numba_kernel_generator", [w_low, w_high], prob, vector, seed, shape,
                      transpose, corder, dtype)


def jitsmm(weight, prob, B, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    return _jit_float("_jit_scalar.float", "_jitsmm_numba_kernel", [weight], prob, B, seed, shape, transpose, corder, dtype)


def jitnmm(w_loc, w_scale, prob, B, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    return _jit_float("_jit_normal.float", "_jitnmm_numba_kernel_generator", [w_loc, w_scale], prob, B, seed, shape,
                      transpose, corder, dtype)


def jitumm(w_low, w_high, prob, B, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    return _jit_float("_jit_uniform.float", "_jitumm_numba_kernel_generator", [w_low, w_high], prob, B, seed, shape,
                      transpose, corder, dtype)


# -- JITC -> CSR materialisation ---------------------------------------------------------------------
def _jit_to_csr(modname, fam, params, prob, seed, shape, corder, matrix_mode, dtype):
    """count + fill through the reference's numba kernels, with the orchestration of
    jits_to_csr (_jit_scalar/csr.py:488-570: row sums, exclusive offsets, lexsort) in numpy."""
    m = R.ref_module(modname)
    n_rows, n_cols = int(shape[0]), int(shape[1])
    chunk = normalize_chunk_size(n_cols)
    n_chunks = -(-n_cols // chunk)
    clen = np.asarray([conn_length(prob)], np.int32)
    seed = np.asarray([seed], np.int32)
    ps = [np.asarray([p], dtype) for p in params]
    kw = dict(shape=(n_rows, n_cols), corder=corder, matrix_mode=matrix_mode)
    cshape = (n_rows, n_chunks) if corder else (n_rows,)
    cc = _first(getattr(m, f"_jit{fam}_csr_count_numba_kernel")(outs=[SDS(cshape, np.int32)], **kw)(*ps, clen, seed))
    row_counts = cc.sum(axis=1).astype(np.int64) if corder else cc.astype(np.int64)
    indptr = np.concatenate([[0], np.cumsum(row_counts)]).astype(np.int64)
    nnz = int(indptr[-1])
    if corder:
        offsets = (indptr[:-1, None] + np.cumsum(cc, axis=1) - cc).astype(np.int32)
    else:
        offsets = indptr.astype(np.int32)
    outs = [SDS((nnz,), np.int32), SDS((nnz,), dtype)]
    if not corder:
        outs.append(SDS((n_rows,), np.int32))
    res = getattr(m, f"_jit{fam}_csr_fill_numba_kernel")(outs=outs, **kw)(*ps, clen, seed, offsets)
    indices, data = res[0], res[1]
    order = np.lexsort((indices, np.repeat(np.arange(n_rows), row_counts)))
    return data[order], indices[order], indptr


def jits_to_csr(weight, prob, seed, *, shape, corder, matrix_mode, dtype=np.float32):
    return _jit_to_csr("_jit_scalar.csr", "s", [weight], prob, seed, shape, corder, matrix_mode, dtype)


def jitn_to_csr(w_loc, w_scale, prob, seed, *, shape, corder, matrix_mode, dtype=np.float32):
    return _jit_to_csr("_jit_normal.csr", "n", [w_loc, w_scale], prob, seed, shape, corder, matrix_mode, dtype)


def jitu_to_csr(w_low, w_high, prob, seed, *, shape, corder, matrix_mode, dtype=np.float32):
    return _jit_to_csr("_jit_uniform.csr", "u", [w_low, w_high], prob, seed, shape, corder, matrix_mode, dtype)


# -- plasticity: the numba kernels add into an aliased copy of the weights; the shim hands them a
#    zero buffer, so updated = weight + returned delta (the update is purely additive) -----------------
def update_csr_on_binary_pre(weight, indices, indptr, pre_spike, post_trace, *, shape):
    m = R.ref_module("_csr.plasticity_binary")
    spk = _spk(pre_spike)
    k = m._csr_on_pre_numba_kernel_generator(spike_info=SDS(spk.shape, spk.dtype), outs=[SDS(weight.shape, weight.dtype)])
    return weight + _first(k(weight, indices, indptr, spk, post_trace))


def update_csr_on_binary_post(weight, indices, indptr, weight_indices, pre_trace, post_spike, *, shape):
    m = R.ref_module("_csr.plasticity_binary")
    spk = _spk(post_spike)
    k = m._csr2csc_on_post_numba_kernel_generator(spike_info=SDS(spk.shape, spk.dtype), outs=[SDS(weight.shape, weight.dtype)])
    return weight + _first(k(weight, indices, indptr, weight_indices, pre_trace, spk))


def fcn_plasticity_row(data, indices, spike, trace):
    m = R.ref_module("_fcn.plasticity_binary")
    spk = _spk(spike)
    k = m._fcn_plasticity_row_numba_kernel(spike_info=SDS(spk.shape, spk.dtype), outs=[SDS(data.shape, data.dtype)])
    return data + _first(k(data, indices, spk, trace))


def update_dense_on_binary_pre(weight, pre_spike, post_trace):
    m = R.ref_module("_dense.plasticity_binary")
    spk = _spk(pre_spike)
    k = m._dense_on_pre_numba_kernel(spike_info=SDS(spk.shape, spk.dtype), outs=[SDS(weight.shape, weight.dtype)])
    return weight + _first(k(weight, spk, post_trace))


def update_dense_on_binary_post(weight, pre_trace, post_spike):
    m = R.ref_module("_dense.plasticity_binary")
    spk = _spk(post_spike)
    k = m._dense_on_post_numba_kernel(spike_info=SDS(spk.shape, spk.dtype), outs=[SDS(weight.shape, weight.dtype)])
    return weight + _first(k(weight, pre_trace, spk))


# -- 2-D spike encoders ----------------------------------------------------------------------------------
def binary_2d_pair_stream_encode(spikes):
    m = R.ref_module("_event.compact")
    s = _spk(spikes)
    n = s.size
    k = m._binary_2d_pair_stream_encode_numba_kernel(spikes_info=SDS(s.shape, s.dtype),
                                                     outs=[SDS((n, 2), np.int32), SDS((1,), np.int32)])
    o = k(s)
    return o[0], o[1]


def binary_2d_row_sparse_encode(spikes, row_size):
    m = R.ref_module("_event.compact")
    s = _spk(spikes)
    k = m._binary_2d_row_sparse_encode_numba_kernel(spikes_info=SDS(s.shape, s.dtype), row_size=row_size,
                                                    outs=[SDS((s.shape[0], row_size), np.int32)])
    return _first(k(s))


def binary_2d_csr_encode(spikes):
    m = R.ref_module("_event.compact")
    s = _spk(spikes)
    counts = _first(m._binary_2d_csr_row_count_numba_kernel(spikes_info=SDS(s.shape, s.dtype),
                                                            outs=[SDS((s.shape[0],), np.int32)])(s))
    indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    indices = _first(m._binary_2d_csr_fill_numba_kernel(spikes_info=SDS(s.shape, s.dtype), indptr_info=SDS(indptr.shape, indptr.dtype),
                                                        outs=[SDS((s.size,), np.int32)])(s, indptr))
    return indices, indptr
