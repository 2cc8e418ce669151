"""Fixed-number connectivity (ELL) event operators and classes.

Mirrors ref: brainevent/_fcn/binary.py (binary_fcnmv :43, binary_fcnmm :512) and _fcn/main.py
(FixedNumPerPre :781, FixedNumPerPost :1042, routing :308-462).
"""
import torch

from . import _lib
from ._csr import _EVENT_TYPES, binary_csrmm_indexed, binary_csrmv_indexed, csr_to_csc_index
from ._event import BinaryArray
from ._op import XLACustomKernel, maybe_decimal, split_mantissa_unit
from ._runtime import as_tensor, event_args, ptr, require_cuda, stream_ptr, workspace, wt_code


def _fcn_kernel(outs=None, shape=None, transpose=False, matmat=False, **kw):
    def kernel(weights, indices, ev):
        n_rows, n_conn = indices.shape
        n_out = shape[1] if transpose else shape[0]
        dev = weights.device
        homo = int(weights.numel() == 1)
        L = _lib.lib()
        if matmat:
            n = ev.batch
            out = (torch.empty((n, n_out), dtype=weights.dtype, device=dev) if transpose
                   else torch.empty((n_out, n), dtype=weights.dtype, device=dev))   # see _csr.py
            ws = workspace(dev, ev.n, n)
            rc = L.be_b200_binary_fcnmm(ptr(weights), wt_code(weights.dtype), homo, ptr(indices), n_rows,
                                        n_conn, ptr(ev.spikes), ptr(ev.aux), ev.kind, ev.n, n_out, n,
                                        int(transpose), int(transpose), ptr(out), ptr(ws), ws.numel(),
                                        stream_ptr(dev))
            _lib.check(rc, "binary_fcnmm")
            if transpose:
                out = out.t()
        else:
            out = torch.empty(n_out, dtype=weights.dtype, device=dev)
            ws = workspace(dev, ev.n, 1)
            rc = L.be_b200_binary_fcnmv(ptr(weights), wt_code(weights.dtype), homo, ptr(indices), n_rows,
                                        n_conn, ptr(ev.spikes), ptr(ev.aux), ev.kind, ev.n, n_out,
                                        int(transpose), ptr(out), ptr(ws), ws.numel(), stream_ptr(dev))
            _lib.check(rc, "binary_fcnmv")
        return [out]
    return kernel


binary_fcnmv_p = XLACustomKernel("binary_fcnmv")
binary_fcnmv_p.def_kernel("b200", "gpu", lambda **kw: _fcn_kernel(matmat=False, **kw), asdefault=True)
binary_fcnmv_p.def_tags("fcn", "binary")
binary_fcnmm_p = XLACustomKernel("binary_fcnmm")
binary_fcnmm_p.def_kernel("b200", "gpu", lambda **kw: _fcn_kernel(matmat=True, **kw), asdefault=True)
binary_fcnmm_p.def_tags("fcn", "binary")


def _fcn_p_call(prim, weights, indices, spikes, shape, transpose, backend, matmat):
    weights, indices = as_tensor(weights), as_tensor(indices)
    if weights.ndim == 0:
        weights = weights.reshape(1)
    require_cuda(weights, indices)
    assert weights.dtype.is_floating_point, "Weights must be a floating-point type."
    assert indices.ndim == 2 and indices.dtype == torch.int32, "indices must be int32 [n_rows, n_conn]"
    assert weights.numel() == 1 or tuple(weights.shape) == tuple(indices.shape), (
        f"weights must have shape (1,) or {tuple(indices.shape)}, got {tuple(weights.shape)}")
    n_pre, n_post = int(shape[0]), int(shape[1])
    ev = event_args(spikes)
    assert (ev.batch is not None) == matmat
    if transpose:
        assert indices.shape[0] == n_pre, f"indices rows {indices.shape[0]} != shape[0] {n_pre}"
        assert ev.n == n_pre, f"events have {ev.n} rows, expected shape[0]={n_pre}"
    else:
        assert indices.shape[0] == n_pre, f"indices rows {indices.shape[0]} != shape[0] {n_pre}"
        assert ev.n == n_post, f"events have {ev.n} rows, expected shape[1]={n_post}"
    return prim(weights.contiguous(), indices.contiguous(), ev, outs=None, shape=(n_pre, n_post),
                transpose=bool(transpose), backend=backend)


def binary_fcnmv_p_call(weights, indices, spikes, *, shape, transpose=False, backend=None):
    return _fcn_p_call(binary_fcnmv_p, weights, indices, spikes, shape, transpose, backend, False)


def binary_fcnmm_p_call(weights, indices, matrix, *, shape, transpose=False, backend=None):
    return _fcn_p_call(binary_fcnmm_p, weights, indices, matrix, shape, transpose, backend, True)


def binary_fcnmv(weights, indices, spikes, *, shape, transpose=False, backend=None):
    """y = W @ s (gather) or W^T @ s (scatter, transpose=True); ref: _fcn/binary.py:43."""
    weights, w_unit = split_mantissa_unit(weights)
    spikes, v_unit = split_mantissa_unit(spikes)
    from . import _autograd
    if _autograd.needs_grad(weights, spikes):
        r = _autograd.fcnmv(weights, torch.as_tensor(indices), spikes, shape, transpose, backend)
    else:
        r = binary_fcnmv_p_call(weights, indices, spikes, shape=shape, transpose=transpose, backend=backend)[0]
    return maybe_decimal(r, v_unit, w_unit)


def binary_fcnmm(weights, indices, matrix, *, shape, transpose=False, backend=None):
    """ref: _fcn/binary.py:512"""
    weights, w_unit = split_mantissa_unit(weights)
    matrix, m_unit = split_mantissa_unit(matrix)
    r = binary_fcnmm_p_call(weights, indices, matrix, shape=shape, transpose=transpose, backend=backend)[0]
    return maybe_decimal(r, m_unit, w_unit)


class FixedNumConn:
    """Logical matrix W [num_pre, num_post] in fixed-connection ELL storage (ref: _fcn/main.py:204)."""

    def __init__(self, args, *, shape, backend=None):
        data, indices = args
        self.data = as_tensor(data)
        if self.data.ndim == 0:
            self.data = self.data.reshape(1)
        self.indices = as_tensor(indices)
        require_cuda(self.data, self.indices)
        if self.data.numel() != 1 and tuple(self.data.shape) != tuple(self.indices.shape):
            raise ValueError(f"Data shape {tuple(self.data.shape)} must match indices shape "
                             f"{tuple(self.indices.shape)}.")
        self.shape = (int(shape[0]), int(shape[1]))
        self.backend = backend
        self._mirror = None

    # subclasses: stored CSR-structure shape and the favourable direction
    @property
    def _a_shape(self):
        raise NotImplementedError

    def _ell_transpose(self, transpose_W: bool) -> bool:
        raise NotImplementedError

    def _weight_indices(self):
        if self._mirror is None:
            n_rows, n_conn = self.indices.shape
            indptr = torch.arange(0, (n_rows + 1) * n_conn, n_conn, dtype=torch.int64 if n_rows * n_conn >= 2**31 else torch.int32,
                                  device=self.indices.device)
            self._mirror = csr_to_csc_index(indptr, self.indices.reshape(-1), shape=self._a_shape)
        return self._mirror

    build_weight_indices = _weight_indices

    def _binary_matvec(self, s, transpose_W: bool):
        a_shape, ell_transpose = self._a_shape, self._ell_transpose(transpose_W)
        if ell_transpose:
            return binary_fcnmv(self.data, self.indices, s, shape=a_shape, transpose=True, backend=self.backend)
        csc_indptr, csc_indices, perm = self._weight_indices()
        return binary_csrmv_indexed(self.data.reshape(-1), csc_indices, csc_indptr, perm, s,
                                    shape=a_shape[::-1], transpose=True, backend=self.backend)

    def _binary_matmat(self, matrix, transpose_W: bool):
        a_shape, ell_transpose = self._a_shape, self._ell_transpose(transpose_W)
        if ell_transpose:
            return binary_fcnmm(self.data, self.indices, matrix, shape=a_shape, transpose=True, backend=self.backend)
        csc_indptr, csc_indices, perm = self._weight_indices()
        return binary_csrmm_indexed(self.data.reshape(-1), csc_indices, csc_indptr, perm, matrix,
                                    shape=a_shape[::-1], transpose=True, backend=self.backend)

    def _dispatch(self, other, transpose_W: bool):
        if isinstance(other, _EVENT_TYPES):
            if other.ndim == 1:
                return self._binary_matvec(other, transpose_W)
            if other.ndim == 2:
                if transpose_W:
                    if not isinstance(other, BinaryArray):
                        raise NotImplementedError("2-D packed events must be packed along the batch axis")
                    return self._binary_matmat(other.value.t(), transpose_W).t()
                return self._binary_matmat(other, transpose_W)
            raise NotImplementedError(f"matmul with object of shape {other.shape}")
        # plain (non-event) operand (ref: _fcn/main.py FixedPreNumConn/FixedPostNumConn __matmul__)
        from ._float import fcnmm, fcnmv
        other = as_tensor(other, device=self.data.device)
        a_shape, ell_transpose = self._a_shape, self._ell_transpose(transpose_W)
        if other.ndim == 1:
            return fcnmv(self.data, self.indices, other, shape=a_shape, transpose=ell_transpose, backend=self.backend)
        if other.ndim == 2:
            if transpose_W:
                return fcnmm(self.data, self.indices, other.t(), shape=a_shape, transpose=ell_transpose,
                             backend=self.backend).t()
            return fcnmm(self.data, self.indices, other, shape=a_shape, transpose=ell_transpose, backend=self.backend)
        raise NotImplementedError(f"matmul with object of shape {tuple(other.shape)}")

    def __matmul__(self, other):
        return self._dispatch(other, transpose_W=False)

    def __rmatmul__(self, other):
        return self._dispatch(other, transpose_W=True)

    def todense(self):
        n_rows, n_conn = self.indices.shape
        a = torch.zeros(self._a_shape, dtype=self.data.dtype, device=self.data.device)
        rows = torch.arange(n_rows, device=a.device).repeat_interleave(n_conn)
        vals = self.data.expand(n_rows, n_conn).reshape(-1) if self.data.numel() == 1 else self.data.reshape(-1)
        a.index_put_((rows, self.indices.reshape(-1).long()), vals, accumulate=True)
        return a if self._a_shape == self.shape else a.t()


class FixedNumPerPre(FixedNumConn):
    """data/indices [num_pre, num_conn] hold post ids: stored structure == W (ref: _fcn/main.py:781)."""
    num_pre = property(lambda self: self.indices.shape[0])
    num_post = property(lambda self: self.shape[1])

    @property
    def _a_shape(self):
        return self.shape

    def _ell_transpose(self, transpose_W: bool) -> bool:
        return bool(transpose_W)


class FixedNumPerPost(FixedNumConn):
    """data/indices [num_post, num_conn] hold pre ids: stored structure == W^T (ref: _fcn/main.py:1042)."""
    num_post = property(lambda self: self.indices.shape[0])
    num_pre = property(lambda self: self.shape[0])

    @property
    def _a_shape(self):
        return self.shape[::-1]

    def _ell_transpose(self, transpose_W: bool) -> bool:
        return not bool(transpose_W)
