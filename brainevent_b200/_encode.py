"""Dense 2-D spike matrix -> sparse event layouts.

Mirrors ref: _event/compact.py binary_2d_pair_stream_encode_p_call (:706), binary_2d_row_sparse_encode_p_call
(:875), binary_2d_csr_row_count_p_call (:1014), binary_2d_csr_fill_p_call (:1136), binary_2d_csr_encode_p_call
(:1180), binary_2d_csc_encode_p_call (:1265).  Same names, output shapes (static capacity, zero-filled
tails) and error behaviour; the order of valid entries is the numba kernels' row-major order.
"""
import torch

from . import _lib
from ._event import _spk_tensor
from ._op import XLACustomKernel
from ._runtime import ptr, stream_ptr, workspace


def _check_2d(spikes):
    if spikes.ndim != 2:
        raise ValueError(f"`spikes` must be 2D, got {spikes.ndim}D.")


def _pair_stream_kernel(outs=None, **kw):
    def kernel(spikes):
        s, kind = _spk_tensor(spikes)
        n_src, n_batch = s.shape
        n = n_src * n_batch
        pairs = torch.empty((n, 2), dtype=torch.int32, device=s.device)
        n_pairs = torch.empty(1, dtype=torch.int32, device=s.device)
        ws = workspace(s.device, n, 1)
        _lib.check(_lib.lib().be_b200_binary_2d_pair_stream_encode(ptr(s), kind, n_src, n_batch, ptr(pairs), ptr(n_pairs),
                                                                   ptr(ws), ws.numel(), stream_ptr(s.device)),
                   "binary_2d_pair_stream_encode")
        return pairs, n_pairs
    return kernel


def _csr_encode_kernel(outs=None, **kw):
    def kernel(spikes):
        s, kind = _spk_tensor(spikes)
        n_src, n_batch = s.shape
        n = n_src * n_batch
        indices = torch.empty(n, dtype=torch.int32, device=s.device)
        indptr = torch.empty(n_src + 1, dtype=torch.int32, device=s.device)
        ws = workspace(s.device, n, 1)
        _lib.check(_lib.lib().be_b200_binary_2d_csr_encode(ptr(s), kind, n_src, n_batch, ptr(indices), ptr(indptr),
                                                           ptr(ws), ws.numel(), stream_ptr(s.device)),
                   "binary_2d_csr_encode")
        return indices, indptr
    return kernel


def _csc_encode_kernel(outs=None, **kw):
    def kernel(spikes):
        s, kind = _spk_tensor(spikes)
        n_src, n_batch = s.shape
        n = n_src * n_batch
        indices = torch.empty(n, dtype=torch.int32, device=s.device)
        indptr = torch.empty(n_batch + 1, dtype=torch.int32, device=s.device)
        scratch = torch.empty(max(n, 1), dtype=torch.uint8, device=s.device)
        ws = workspace(s.device, n, 1)
        _lib.check(_lib.lib().be_b200_binary_2d_csc_encode(ptr(s), kind, n_src, n_batch, ptr(indices), ptr(indptr),
                                                           ptr(scratch), ptr(ws), ws.numel(), stream_ptr(s.device)),
                   "binary_2d_csc_encode")
        return indices, indptr
    return kernel


def _row_sparse_kernel(outs=None, row_size=None, **kw):
    def kernel(spikes):
        s, kind = _spk_tensor(spikes)
        n_src, n_batch = s.shape
        out = torch.empty((n_src, row_size), dtype=torch.int32, device=s.device)
        max_nnz = torch.empty(1, dtype=torch.int32, device=s.device)
        _lib.check(_lib.lib().be_b200_binary_2d_row_sparse_encode(ptr(s), kind, n_src, n_batch, row_size, ptr(out),
                                                                  ptr(max_nnz), stream_ptr(s.device)),
                   "binary_2d_row_sparse_encode")
        return out, max_nnz
    return kernel


def _reg(name, kg, *tags):
    p = XLACustomKernel(name)
    p.def_kernel("b200", "gpu", kg, asdefault=True)
    p.def_tags(*tags)
    return p


binary_2d_pair_stream_encode_p = _reg("binary_2d_pair_stream_encode", _pair_stream_kernel, "event", "binary")
binary_2d_csr_encode_p = _reg("binary_2d_csr_encode", _csr_encode_kernel, "event", "binary", "csr")
binary_2d_csc_encode_p = _reg("binary_2d_csc_encode", _csc_encode_kernel, "event", "binary", "csc")
binary_2d_row_sparse_encode_p = _reg("binary_2d_row_sparse_encode", _row_sparse_kernel, "event", "binary")


def binary_2d_pair_stream_encode_p_call(spikes, *, backend=None):
    """-> (pair_stream int32 [n_src*n_batch, 2], n_pairs int32 [1]); ref: _event/compact.py:706."""
    _check_2d(spikes)
    return binary_2d_pair_stream_encode_p(spikes, outs=None, backend=backend)


def binary_2d_csr_encode_p_call(spikes, *, backend=None):
    """-> (indices int32 [n_src*n_batch], indptr int32 [n_src+1]); ref: _event/compact.py:1180."""
    _check_2d(spikes)
    return binary_2d_csr_encode_p(spikes, outs=None, backend=backend)


def binary_2d_csr_row_count_p_call(spikes, *, backend=None):
    """-> (row_counts int32 [n_src],); ref: _event/compact.py:1014 (derived from the same encode)."""
    _check_2d(spikes)
    indptr = binary_2d_csr_encode_p(spikes, outs=None, backend=backend)[1]
    return ((indptr[1:] - indptr[:-1]).to(torch.int32),)


def binary_2d_csr_fill_p_call(spikes, indptr, *, backend=None):
    """-> (indices,); the fill pass of the reference's two-pass encode (:1136); `indptr` is implied
    by the spikes and only checked for shape."""
    _check_2d(spikes)
    assert indptr.shape[0] == spikes.shape[0] + 1, "indptr must have shape (n_src + 1,)"
    return (binary_2d_csr_encode_p(spikes, outs=None, backend=backend)[0],)


def binary_2d_csc_encode_p_call(spikes, *, backend=None):
    """-> (indices int32 [n_src*n_batch] of row ids, indptr int32 [n_batch+1]); ref: _event/compact.py:1265."""
    _check_2d(spikes)
    return binary_2d_csc_encode_p(spikes, outs=None, backend=backend)


def binary_2d_row_sparse_encode_p_call(spikes, *, row_size: int, backend=None):
    """-> (spike_indices int32 [n_src, row_size],) of 1-based active column ids, zero padded;
    ref: _event/compact.py:875."""
    _check_2d(spikes)
    if row_size <= 0:
        raise ValueError(f"`row_size` must be positive, got {row_size}.")
    n_src, n_batch = spikes.shape
    if row_size > n_batch:
        raise ValueError(f"`row_size` must be <= n_batch={n_batch}, got {row_size}.")
    out, max_nnz = binary_2d_row_sparse_encode_p(spikes, outs=None, row_size=int(row_size), backend=backend)
    worst = int(max_nnz.item())       # the reference validates eagerly too (compact.py:853-872)
    if worst > row_size:
        raise ValueError(f"`row_size={row_size}` is too small for the input spikes; max row NNZ is {worst}.")
    return (out,)
