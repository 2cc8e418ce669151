"""CSR / CSC event operators and matrix classes.

Mirrors ref: brainevent/_csr/binary.py (binary_csrmv :128, binary_csrmm :264, *_p_call :827,
:1452), _csr/binary_indexed.py (:70, :615), _csr/main.py (CSR.__matmul__/__rmatmul__ :1595-1776,
CSC :2509-2696) and _misc.py:1341-1384 (csr_to_csc_index).  Same names, argument meaning and
assertion behaviour; arrays are torch CUDA tensors and the work happens in libbrainevent_b200.so.
"""
import torch

from . import _lib
from ._event import BinaryArray, BitPackedBinary, CompactBinary
from ._op import XLACustomKernel, maybe_decimal, split_mantissa_unit
from ._runtime import as_tensor, csc_workspace, event_args, ptr, require_cuda, stream_ptr, workspace, wt_code

_EVENT_TYPES = (BinaryArray, BitPackedBinary, CompactBinary)


def _prep_structure(data, indices, indptr, perm=None):
    data = as_tensor(data)
    if data.ndim == 0:
        data = data.reshape(1)
    indices, indptr = as_tensor(indices), as_tensor(indptr)
    require_cuda(data, indices, indptr, perm)
    assert data.dtype.is_floating_point, "Weights must be a floating-point type."
    assert indices.dtype == torch.int32, f"indices must be int32, got {indices.dtype}"
    assert indptr.dtype in (torch.int32, torch.int64), f"indptr must be int32/int64, got {indptr.dtype}"
    if perm is not None:
        perm = as_tensor(perm)
        assert perm.dtype in (torch.int32, torch.int64), f"perm must be int32/int64, got {perm.dtype}"
        perm = perm.contiguous()
    return data.contiguous(), indices.contiguous(), indptr.contiguous(), perm


def _csr_kernel(outs=None, shape=None, transpose=False, matmat=False, **kw):
    m, k = shape

    def kernel(data, indices, indptr, perm, ev):
        n_out = k if transpose else m
        dev = data.device
        homo = int(data.numel() == 1)
        nse = indices.numel()
        L = _lib.lib()
        if matmat:
            n = ev.batch
            # scatter: physical [n, n_out] (each batch column a contiguous output vector, the
            # reference's own physical layout, _csr/binary.py:1263-1286), returned as the logical
            # [n_out, n] view; gather: [n_out, n] row-major
            out = (torch.empty((n, n_out), dtype=data.dtype, device=dev) if transpose
                   else torch.empty((n_out, n), dtype=data.dtype, device=dev))
            ws = workspace(dev, ev.n, n)
            rc = L.be_b200_binary_csrmm(ptr(data), wt_code(data.dtype), homo, ptr(indices), ptr(indptr),
                                        int(indptr.dtype == torch.int64), ptr(perm),
                                        int(perm is not None and perm.dtype == torch.int64),
                                        ptr(ev.spikes), ptr(ev.aux), ev.kind, m, k, nse, n, int(transpose),
                                        int(transpose), ptr(out), ptr(ws), ws.numel(), stream_ptr(dev))
            _lib.check(rc, "binary_csrmm")
            if transpose:
                out = out.t()
        else:
            out = torch.empty(n_out, dtype=data.dtype, device=dev)
            ws = workspace(dev, ev.n, 1)
            rc = L.be_b200_binary_csrmv(ptr(data), wt_code(data.dtype), homo, ptr(indices), ptr(indptr),
                                        int(indptr.dtype == torch.int64), ptr(perm),
                                        int(perm is not None and perm.dtype == torch.int64),
                                        ptr(ev.spikes), ptr(ev.aux), ev.kind, m, k, nse, int(transpose),
                                        ptr(out), ptr(ws), ws.numel(), stream_ptr(dev))
            _lib.check(rc, "binary_csrmv")
        return [out]
    return kernel


binary_csrmv_p = XLACustomKernel("binary_csrmv")
binary_csrmv_p.def_kernel("b200", "gpu", lambda **kw: _csr_kernel(matmat=False, **kw), asdefault=True)
binary_csrmv_p.def_tags("csr", "binary")
binary_csrmm_p = XLACustomKernel("binary_csrmm")
binary_csrmm_p.def_kernel("b200", "gpu", lambda **kw: _csr_kernel(matmat=True, **kw), asdefault=True)
binary_csrmm_p.def_tags("csr", "binary")
binary_csrmv_indexed_p = XLACustomKernel("binary_csrmv_indexed")
binary_csrmv_indexed_p.def_kernel("b200", "gpu", lambda **kw: _csr_kernel(matmat=False, **kw), asdefault=True)
binary_csrmv_indexed_p.def_tags("csr", "binary", "indexed")
binary_csrmm_indexed_p = XLACustomKernel("binary_csrmm_indexed")
binary_csrmm_indexed_p.def_kernel("b200", "gpu", lambda **kw: _csr_kernel(matmat=True, **kw), asdefault=True)
binary_csrmm_indexed_p.def_tags("csr", "binary", "indexed")


def _p_call(prim, data, indices, indptr, perm, v, shape, transpose, backend, matmat):
    data, indices, indptr, perm = _prep_structure(data, indices, indptr, perm)
    assert len(shape) == 2, "shape must be (rows, cols)"
    m, k = int(shape[0]), int(shape[1])
    assert indptr.ndim == 1 and indptr.shape[0] == m + 1, (
        f"indptr must have shape ({m + 1},), got {tuple(indptr.shape)}")
    assert indices.ndim == 1
    assert data.shape == (1,) or data.shape == indices.shape, (
        f"data must have shape (1,) or {tuple(indices.shape)}, got {tuple(data.shape)}")
    ev = event_args(v)
    if matmat:
        assert ev.batch is not None, "binary_csrmm needs a 2-D event matrix"
    else:
        assert ev.batch is None, "binary_csrmv needs a 1-D event vector"
    n_src = m if transpose else k
    assert ev.n == n_src, (f"shape mismatch: events have {ev.n} rows but the "
                           f"{'transposed ' if transpose else ''}matrix expects {n_src}")
    return prim(data, indices, indptr, perm, ev, outs=None, shape=(m, k), transpose=bool(transpose),
                backend=backend)


def binary_csrmv_p_call(data, indices, indptr, v, *, shape, transpose=False, backend=None, workspace=None):
    return _p_call(binary_csrmv_p, data, indices, indptr, None, v, shape, transpose, backend, False)


def binary_csrmm_p_call(data, indices, indptr, B, *, shape, transpose=False, backend=None, workspace=None):
    return _p_call(binary_csrmm_p, data, indices, indptr, None, B, shape, transpose, backend, True)


def binary_csrmv(data, indices, indptr, v, *, shape, transpose=False, backend=None, workspace=None):
    """y = A @ e(v) (or A^T @ e(v)); only active entries of v contribute (ref: _csr/binary.py:128)."""
    data, unitd = split_mantissa_unit(data)
    v, unitv = split_mantissa_unit(v)
    from . import _autograd
    if _autograd.needs_grad(data, v):      # torch.autograd rule (ref: the primitive's jvp / transpose rules, :656-722)
        res = _autograd.csrmv(data, as_tensor(indices), as_tensor(indptr), v, shape, transpose, backend)
    else:
        res = binary_csrmv_p_call(data, indices, indptr, v, shape=shape, transpose=transpose, backend=backend)[0]
    return maybe_decimal(res, unitd, unitv)


def binary_csrmm(data, indices, indptr, B, *, shape, transpose=False, backend=None, workspace=None):
    """Y = A @ e(B) (or A^T @ e(B)), B [k or m, n] (ref: _csr/binary.py:264)."""
    data, unitd = split_mantissa_unit(data)
    B, unitb = split_mantissa_unit(B)
    res = binary_csrmm_p_call(data, indices, indptr, B, shape=shape, transpose=transpose, backend=backend)[0]
    return maybe_decimal(res, unitd, unitb)


def binary_csrmv_indexed(data, indices, indptr, perm, v, *, shape, workspace=None, transpose=False,
                         backend=None):
    """binary_csrmv(data[perm], ...) with the gather fused (ref: _csr/binary_indexed.py:70)."""
    data, unitd = split_mantissa_unit(data)
    v, unitv = split_mantissa_unit(v)
    res = _p_call(binary_csrmv_indexed_p, data, indices, indptr, perm, v, shape, transpose, backend, False)[0]
    return maybe_decimal(res, unitd, unitv)


def binary_csrmm_indexed(data, indices, indptr, perm, B, *, shape, workspace=None, transpose=False,
                         backend=None):
    """ref: _csr/binary_indexed.py:615"""
    data, unitd = split_mantissa_unit(data)
    B, unitb = split_mantissa_unit(B)
    res = _p_call(binary_csrmm_indexed_p, data, indices, indptr, perm, B, shape, transpose, backend, True)[0]
    return maybe_decimal(res, unitd, unitb)


SCATTER_GATHER_RATE_THRESHOLD = 0.25   # above this fraction of active sources the gather wins


def binary_csrmv_auto(data, indices, indptr, mirror, v, *, shape, rate_threshold=SCATTER_GATHER_RATE_THRESHOLD):
    """`e(v) @ A` for a CSR `A` [m, k] whose CSC mirror `(csc_indptr, csc_indices, perm)` is
    available: atomic scatter over active rows or deterministic post-major gather, chosen on the
    device from the spike count (no host sync)."""
    data, unitd = split_mantissa_unit(data)
    v, unitv = split_mantissa_unit(v)
    data, indices, indptr, _ = _prep_structure(data, indices, indptr)
    csc_indptr, csc_indices, perm = mirror
    m, k = int(shape[0]), int(shape[1])
    ev = event_args(v, n_expected=m)
    assert ev.batch is None, "binary_csrmv_auto needs a 1-D event vector"
    assert csc_indptr.dtype == indptr.dtype == perm.dtype or data.numel() == 1
    dev = data.device
    out = torch.empty(k, dtype=data.dtype, device=dev)
    ws = workspace(dev, m, 1)
    rc = _lib.lib().be_b200_binary_csrmv_auto(ptr(data), wt_code(data.dtype), int(data.numel() == 1), ptr(indices),
                                              ptr(indptr), int(indptr.dtype == torch.int64), ptr(csc_indptr),
                                              ptr(csc_indices), ptr(perm), ptr(ev.spikes), ptr(ev.aux), ev.kind,
                                              m, k, indices.numel(), float(rate_threshold), ptr(out), ptr(ws),
                                              ws.numel(), stream_ptr(dev))
    _lib.check(rc, "binary_csrmv_auto")
    return maybe_decimal(out, unitd, unitv)


def csr_to_csc_index(csr_indptr, csr_indices, *, shape):
    """Stable by-column mirror -> (csc_indptr, csc_indices = row ids, perm); ref: _misc.py:1341-1384."""
    indptr, indices = as_tensor(csr_indptr).contiguous(), as_tensor(csr_indices).contiguous()
    require_cuda(indptr, indices)
    m, n_post = int(shape[0]), int(shape[1])
    nse = indices.numel()
    is64 = indptr.dtype == torch.int64
    odt = torch.int64 if is64 else torch.int32
    dev = indptr.device
    csc_indptr = torch.empty(n_post + 1, dtype=odt, device=dev)
    csc_indices = torch.empty(nse, dtype=torch.int32, device=dev)
    perm = torch.empty(nse, dtype=odt, device=dev)
    ws = csc_workspace(dev, m, n_post, nse)
    _lib.check(_lib.lib().be_b200_csr_to_csc(ptr(indptr), int(is64), ptr(indices), m, n_post, nse,
                                             ptr(csc_indptr), ptr(csc_indices), ptr(perm), ptr(ws),
                                             ws.numel(), stream_ptr(dev)), "csr_to_csc")
    return csc_indptr, csc_indices, perm


# ---------------------------------------------------------------------------------------------
# matrix classes (routing only)
# ---------------------------------------------------------------------------------------------
def _float_matmul(data, indices, indptr, other, a_shape, transpose, backend, left=None):
    """Plain (non-event) operand: route to csrmv / csrmm (ref: _csr/main.py:1678-1700, :1756-1776).
    `left`: whether the matrix is the left factor (defaults to `not transpose`, the CSR case); a
    right-hand 2-D operand [b, rows] is transposed around the kernel call."""
    from ._float import csrmm, csrmv
    other = as_tensor(other, device=data.device)
    if left is None:
        left = not transpose
    if other.ndim == 1:
        return csrmv(data, indices, indptr, other, shape=a_shape, transpose=transpose, backend=backend)
    if other.ndim == 2:
        if left:
            return csrmm(data, indices, indptr, other, shape=a_shape, transpose=transpose, backend=backend)
        return csrmm(data, indices, indptr, other.t(), shape=a_shape, transpose=transpose, backend=backend).t()
    raise NotImplementedError(f"matmul with object of shape {tuple(other.shape)}")


class _Compressed:
    def __init__(self, args, *, shape, backend=None):
        data, indices, indptr = args
        self.data, self.indices, self.indptr, _ = _prep_structure(data, indices, indptr)
        self.shape = (int(shape[0]), int(shape[1]))
        self.backend = backend
        self._mirror = None

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def nse(self):
        return self.indices.numel()


class CSR(_Compressed):
    """Row-compressed matrix [m, k] (ref: _csr/main.py:977).  `spk @ csr` is the SNN direction
    (pre-spikes x (pre, post) weights): active-row scatter.  `csr @ spk` gathers."""

    def _weight_indices(self):
        if self._mirror is None:
            self._mirror = csr_to_csc_index(self.indptr, self.indices, shape=self.shape)
        return self._mirror

    build_weight_indices = _weight_indices

    @property
    def T(self):
        return CSC((self.data, self.indices, self.indptr), shape=self.shape[::-1], backend=self.backend)

    def todense(self):
        m, k = self.shape
        rows = torch.repeat_interleave(torch.arange(m, device=self.data.device),
                                       (self.indptr[1:] - self.indptr[:-1]).long())
        out = torch.zeros(m, k, dtype=self.data.dtype, device=self.data.device)
        vals = self.data.expand(self.nse) if self.data.numel() == 1 else self.data
        out.index_put_((rows, self.indices.long()), vals, accumulate=True)
        return out

    def __matmul__(self, other):
        # csr @ other (ref: _csr/main.py:1633-1700)
        if isinstance(other, _EVENT_TYPES):
            if other.ndim == 1:
                return binary_csrmv(self.data, self.indices, self.indptr, other, shape=self.shape,
                                    transpose=False, backend=self.backend)
            if other.ndim == 2:
                return binary_csrmm(self.data, self.indices, self.indptr, other, shape=self.shape,
                                    transpose=False, backend=self.backend)
            raise NotImplementedError(f"matmul with object of shape {other.shape}")
        return _float_matmul(self.data, self.indices, self.indptr, other, self.shape, False, self.backend)

    def __rmatmul__(self, other):
        # other @ csr (ref: _csr/main.py:1731-1776)
        if isinstance(other, _EVENT_TYPES):
            if other.ndim == 1:
                if self._mirror is not None:
                    # mirror built (build_weight_indices): scatter vs gather picked by spike rate
                    return binary_csrmv_auto(self.data, self.indices, self.indptr, self._mirror, other,
                                             shape=self.shape)
                return binary_csrmv(self.data, self.indices, self.indptr, other, shape=self.shape,
                                    transpose=True, backend=self.backend)
            if other.ndim == 2:
                if not isinstance(other, BinaryArray):
                    raise NotImplementedError("2-D packed events must be packed along the batch axis; "
                                              "use binary_csrmm(..., transpose=True) directly")
                r = binary_csrmm(self.data, self.indices, self.indptr, other.value.t(), shape=self.shape,
                                 transpose=True, backend=self.backend)
                return r.t()
            raise NotImplementedError(f"matmul with object of shape {other.shape}")
        return _float_matmul(self.data, self.indices, self.indptr, other, self.shape, True, self.backend)


class CSC(_Compressed):
    """Column-compressed matrix of logical shape [m, k]: (data, indices, indptr) compress the
    columns, i.e. they are the CSR of the transpose (ref: _csr/main.py:2509-2696)."""

    def _weight_indices(self):
        if self._mirror is None:
            self._mirror = csr_to_csc_index(self.indptr, self.indices, shape=self.shape[::-1])
        return self._mirror

    build_weight_indices = _weight_indices

    @property
    def T(self):
        return CSR((self.data, self.indices, self.indptr), shape=self.shape[::-1], backend=self.backend)

    def todense(self):
        return self.T.todense().t()

    def __matmul__(self, other):
        # csc @ other == (csr of the transpose)^T @ other: favourable scatter
        if isinstance(other, _EVENT_TYPES):
            if other.ndim == 1:
                return binary_csrmv(self.data, self.indices, self.indptr, other, shape=self.shape[::-1],
                                    transpose=True, backend=self.backend)
            if other.ndim == 2:
                return binary_csrmm(self.data, self.indices, self.indptr, other, shape=self.shape[::-1],
                                    transpose=True, backend=self.backend)
            raise NotImplementedError(f"matmul with object of shape {other.shape}")
        return _float_matmul(self.data, self.indices, self.indptr, other, self.shape[::-1], True, self.backend, left=True)

    def __rmatmul__(self, other):
        # other @ csc: gather over the stored (transposed) rows
        if isinstance(other, _EVENT_TYPES):
            if other.ndim == 1:
                return binary_csrmv(self.data, self.indices, self.indptr, other, shape=self.shape[::-1],
                                    transpose=False, backend=self.backend)
            if other.ndim == 2:
                if not isinstance(other, BinaryArray):
                    raise NotImplementedError("2-D packed events must be packed along the batch axis")
                r = binary_csrmm(self.data, self.indices, self.indptr, other.value.t(),
                                 shape=self.shape[::-1], transpose=False, backend=self.backend)
                return r.t()
            raise NotImplementedError(f"matmul with object of shape {other.shape}")
        return _float_matmul(self.data, self.indices, self.indptr, other, self.shape[::-1], False, self.backend, left=False)
