"""numpy-in / numpy-out adapter over brainevent_b200 (the CUDA path through the C ABI), with the
same call signatures as `oracle`, so one test table drives both."""
import numpy as np
import torch

import brainevent_b200 as be

DEV = "cuda"


def _t(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x), device=DEV)
    return t if dtype is None else t.to(dtype)


def _np(t):
    return t.detach().cpu().numpy()


def binary_csrmv(data, indices, indptr, v, *, shape, transpose=False, perm=None):
    if perm is not None:
        return _np(be.binary_csrmv_indexed(_t(np.atleast_1d(data)), _t(indices), _t(indptr), _t(perm), _t(v),
                                           shape=shape, transpose=transpose))
    return _np(be.binary_csrmv(_t(np.atleast_1d(data)), _t(indices), _t(indptr), _t(v), shape=shape,
                               transpose=transpose))


def binary_csrmm(data, indices, indptr, B, *, shape, transpose=False, perm=None):
    if perm is not None:
        return _np(be.binary_csrmm_indexed(_t(np.atleast_1d(data)), _t(indices), _t(indptr), _t(perm), _t(B),
                                           shape=shape, transpose=transpose))
    return _np(be.binary_csrmm(_t(np.atleast_1d(data)), _t(indices), _t(indptr), _t(B), shape=shape,
                               transpose=transpose))


def binary_csrmv_indexed(data, indices, indptr, perm, v, *, shape, transpose=False):
    return binary_csrmv(data, indices, indptr, v, shape=shape, transpose=transpose, perm=perm)


def binary_csrmm_indexed(data, indices, indptr, perm, B, *, shape, transpose=False):
    return binary_csrmm(data, indices, indptr, B, shape=shape, transpose=transpose, perm=perm)


def binary_fcnmv(weights, indices, spikes, *, shape, transpose=False):
    return _np(be.binary_fcnmv(_t(np.atleast_1d(weights)), _t(indices), _t(spikes), shape=shape,
                               transpose=transpose))


def binary_fcnmm(weights, indices, matrix, *, shape, transpose=False):
    return _np(be.binary_fcnmm(_t(np.atleast_1d(weights)), _t(indices), _t(matrix), shape=shape,
                               transpose=transpose))


def _jit(fn, params, prob, x, seed, shape, transpose, corder, dtype):
    td = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
    ps = [torch.tensor([p], dtype=td, device=DEV) for p in params]
    return _np(fn(*ps, prob, _t(x), seed, shape=shape, transpose=transpose, corder=corder))


def binary_jitsmv(w, prob, v, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    return _jit(be.binary_jitsmv, [w], prob, v, seed, shape, transpose, corder, dtype)


def binary_jitsmm(w, prob, B, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    return _jit(be.binary_jitsmm, [w], prob, B, seed, shape, transpose, corder, dtype)


def binary_jitnmv(a, b, prob, v, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    return _jit(be.binary_jitnmv, [a, b], prob, v, seed, shape, transpose, corder, dtype)


def binary_jitnmm(a, b, prob, B, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    return _jit(be.binary_jitnmm, [a, b], prob, B, seed, shape, transpose, corder, dtype)


def binary_jitumv(a, b, prob, v, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    return _jit(be.binary_jitumv, [a, b], prob, v, seed, shape, transpose, corder, dtype)


def binary_jitumm(a, b, prob, B, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    return _jit(be.binary_jitumm, [a, b], prob, B, seed, shape, transpose, corder, dtype)


def binary_densemv(W, spikes, *, transpose):
    return _np(be.binary_densemv(_t(W), _t(spikes), transpose=transpose))


def binary_densemm(W, spikes, *, transpose):
    return _np(be.binary_densemm(_t(W), _t(spikes), transpose=transpose))


def binary_1d_array_index(spikes):
    ids, na = be.binary_1d_array_index_p_call(_t(spikes))
    return _np(ids), _np(na)


def binary_2d_array_index(spikes):
    p, ids, na = be.binary_2d_array_index_p_call(_t(spikes))
    return _np(p).view(np.uint32), _np(ids), _np(na)


def binary_2d_compact_only(spikes):
    ids, na = be.binary_2d_compact_only_p_call(_t(spikes))
    return _np(ids), _np(na)


def bitpack(arr, axis=0):
    return _np(be.bitpack(_t(arr), axis=axis).contiguous()).view(np.uint32)


def csr_to_csc_index(indptr, indices, shape):
    a, b, c = be.csr_to_csc_index(_t(indptr), _t(indices), shape=shape)
    return _np(a).astype(np.int64), _np(b), _np(c).astype(np.int64)


# ---- float (non-event) operators ----------------------------------------------------------------
def csrmv(data, indices, indptr, v, *, shape, transpose=False):
    return _np(be.csrmv(_t(np.atleast_1d(data)), _t(indices), _t(indptr), _t(v), shape=shape, transpose=transpose))


def csrmm(data, indices, indptr, B, *, shape, transpose=False):
    return _np(be.csrmm(_t(np.atleast_1d(data)), _t(indices), _t(indptr), _t(B), shape=shape, transpose=transpose))


def fcnmv(weights, indices, v, *, shape, transpose=False):
    return _np(be.fcnmv(_t(np.atleast_1d(weights)), _t(indices), _t(v), shape=shape, transpose=transpose))


def fcnmm(weights, indices, B, *, shape, transpose=False):
    return _np(be.fcnmm(_t(np.atleast_1d(weights)), _t(indices), _t(B), shape=shape, transpose=transpose))


def jitsmv(w, prob, v, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    return _jit(be.jitsmv, [w], prob, v, seed, shape, transpose, corder, dtype)


def jitsmm(w, prob, B, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    return _jit(be.jitsmm, [w], prob, B, seed, shape, transpose, corder, dtype)


def jitnmv(loc, scale, prob, v, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    return _jit(be.jitnmv, [loc, scale], prob, v, seed, shape, transpose, corder, dtype)


def jitnmm(loc, scale, prob, B, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    return _jit(be.jitnmm, [loc, scale], prob, B, seed, shape, transpose, corder, dtype)


def jitumv(lo, hi, prob, v, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    return _jit(be.jitumv, [lo, hi], prob, v, seed, shape, transpose, corder, dtype)


def jitumm(lo, hi, prob, B, seed, *, shape, transpose=False, corder=True, dtype=np.float32):
    return _jit(be.jitumm, [lo, hi], prob, B, seed, shape, transpose, corder, dtype)


def _to_csr(fn, params, prob, seed, shape, corder, matrix_mode, dtype):
    td = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
    ps = [torch.tensor([p], dtype=td, device=DEV) for p in params]
    c = fn(*ps, prob, seed, shape=shape, corder=corder, matrix_mode=matrix_mode)
    return _np(c.data), _np(c.indices), _np(c.indptr).astype(np.int64)


def jits_to_csr(w, prob, seed, *, shape, corder, matrix_mode, dtype=np.float32):
    return _to_csr(be.jits_to_csr, [w], prob, seed, shape, corder, matrix_mode, dtype)


def jitn_to_csr(loc, scale, prob, seed, *, shape, corder, matrix_mode, dtype=np.float32):
    return _to_csr(be.jitn_to_csr, [loc, scale], prob, seed, shape, corder, matrix_mode, dtype)


def jitu_to_csr(lo, hi, prob, seed, *, shape, corder, matrix_mode, dtype=np.float32):
    return _to_csr(be.jitu_to_csr, [lo, hi], prob, seed, shape, corder, matrix_mode, dtype)


# ---- plasticity -----------------------------------------------------------------------------------
def update_csr_on_binary_pre(weight, indices, indptr, pre_spike, post_trace, w_min=None, w_max=None, *, shape):
    return _np(be.update_csr_on_binary_pre(_t(weight), _t(indices), _t(indptr), _t(pre_spike), _t(post_trace), w_min,
                                           w_max, shape=shape))


def update_csr_on_binary_post(weight, indices, indptr, weight_indices, pre_trace, post_spike, w_min=None, w_max=None,
                              *, shape):
    return _np(be.update_csr_on_binary_post(_t(weight), _t(indices), _t(indptr), _t(weight_indices), _t(pre_trace),
                                            _t(post_spike), w_min, w_max, shape=shape))


def fcn_plasticity_row(data, indices, spike, trace, w_min=None, w_max=None):
    n_rows = indices.shape[0]
    return _np(be.update_fixed_post_conn_on_binary_pre(_t(data), _t(indices), _t(spike), _t(trace), w_min, w_max,
                                                       shape=(n_rows, trace.shape[0])))


def update_dense_on_binary_pre(weight, pre_spike, post_trace, w_min=None, w_max=None):
    return _np(be.update_dense_on_binary_pre(_t(weight), _t(pre_spike), _t(post_trace), w_min, w_max))


def update_dense_on_binary_post(weight, pre_trace, post_spike, w_min=None, w_max=None):
    return _np(be.update_dense_on_binary_post(_t(weight), _t(pre_trace), _t(post_spike), w_min, w_max))


# ---- 2-D spike encoders ------------------------------------------------------------------------------
def binary_2d_pair_stream_encode(spikes):
    p, n = be.binary_2d_pair_stream_encode_p_call(_t(spikes))
    return _np(p), _np(n)


def binary_2d_row_sparse_encode(spikes, row_size):
    return _np(be.binary_2d_row_sparse_encode_p_call(_t(spikes), row_size=row_size)[0])


def binary_2d_csr_encode(spikes):
    i, p = be.binary_2d_csr_encode_p_call(_t(spikes))
    return _np(i), _np(p)


def binary_2d_csc_encode(spikes):
    i, p = be.binary_2d_csc_encode_p_call(_t(spikes))
    return _np(i), _np(p)
