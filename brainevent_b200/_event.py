"""Event representations and the compaction / bit-pack primitives.

Mirrors ref: brainevent/_event/{base,binary,bitpack_binary,compact_binary,compact}.py.
In the reference BitPackedBinary / CompactBinary are containers that no sparse operator consumes
(SURVEY F6); here every operator accepts them directly (event kinds PACKED / COMPACT of the C
ABI), which is what makes the multi-GPU spike all-gather 8x smaller than a byte mask.
"""
import ctypes

import torch

from . import _lib
from ._error import MathError
from ._op import XLACustomKernel
from ._runtime import as_tensor, ptr, require_cuda, stream_ptr, workspace


# ---------------------------------------------------------------------------------------------
# primitives
# ---------------------------------------------------------------------------------------------
def _spk_tensor(spikes, for_compaction=True):
    t = as_tensor(spikes)
    require_cuda(t)
    if t.dtype in (torch.bool, torch.uint8, torch.int8):
        return t.contiguous(), _lib.SPK_BOOL
    if t.dtype == torch.float32:
        return t.contiguous(), _lib.SPK_F32
    return ((t != 0) if for_compaction else (t > 0)).contiguous(), _lib.SPK_BOOL


def _bitpack_kernel(outs=None, axis=0, **kw):
    def kernel(arr):
        t = as_tensor(arr)
        require_cuda(t)
        if t.ndim == 1:
            s, kind = _spk_tensor(t)
            n = s.shape[0]
            packed = torch.empty((n + 31) // 32, dtype=torch.int32, device=s.device)
            _lib.check(_lib.lib().be_b200_bitpack_1d(ptr(s), kind, n, 1, ptr(packed), stream_ptr(s.device)),
                       "bitpack_1d")
            return packed
        if t.ndim != 2:
            raise ValueError("bitpack supports 1-D and 2-D arrays")
        ax = axis % 2
        # pack along axis 1 natively; axis 0 is the same kernel on the transposed matrix
        m = t if ax == 1 else t.t()
        s, kind = _spk_tensor(m)
        n, b = s.shape
        wpr = (b + 31) // 32
        packed = torch.empty((n, wpr), dtype=torch.int32, device=s.device)
        ids = torch.empty(n, dtype=torch.int32, device=s.device)
        na = torch.empty(1, dtype=torch.int32, device=s.device)
        ws = workspace(s.device, n, b)
        _lib.check(_lib.lib().be_b200_compact_2d(ptr(s), kind, n, b, 1, ptr(packed), ptr(ids), ptr(na),
                                                 ptr(ws), ws.numel(), stream_ptr(s.device)), "bitpack_2d")
        return packed if ax == 1 else packed.t()
    return kernel


def _compact_1d_kernel(outs=None, **kw):
    def kernel(spikes):
        s, kind = _spk_tensor(spikes)
        n = s.shape[0]
        ids = torch.empty(n, dtype=torch.int32, device=s.device)
        na = torch.empty(1, dtype=torch.int32, device=s.device)
        ws = workspace(s.device, n, 1)
        _lib.check(_lib.lib().be_b200_compact_1d(ptr(s), kind, n, 1, ptr(ids), ptr(na), None, ptr(ws),
                                                 ws.numel(), stream_ptr(s.device)), "compact_1d")
        return ids, na
    return kernel


def _compact_2d_kernel(outs=None, with_packed=True, **kw):
    def kernel(spikes):
        s, kind = _spk_tensor(spikes)
        n, b = s.shape
        packed = torch.empty((n, (b + 31) // 32), dtype=torch.int32, device=s.device) if with_packed else None
        ids = torch.empty(n, dtype=torch.int32, device=s.device)
        na = torch.empty(1, dtype=torch.int32, device=s.device)
        ws = workspace(s.device, n, b)
        _lib.check(_lib.lib().be_b200_compact_2d(ptr(s), kind, n, b, 1, ptr(packed), ptr(ids), ptr(na),
                                                 ptr(ws), ws.numel(), stream_ptr(s.device)), "compact_2d")
        return (packed, ids, na) if with_packed else (ids, na)
    return kernel


bitpack_p = XLACustomKernel("bitpack")
bitpack_p.def_kernel("b200", "gpu", _bitpack_kernel, asdefault=True)
bitpack_p.def_tags("event", "bitpack")

binary_1d_array_index_p = XLACustomKernel("binary_1d_array_index")
binary_1d_array_index_p.def_kernel("b200", "gpu", _compact_1d_kernel, asdefault=True)
binary_1d_array_index_p.def_tags("event", "compact")

binary_2d_array_index_p = XLACustomKernel("binary_2d_array_index")
binary_2d_array_index_p.def_kernel("b200", "gpu", lambda **kw: _compact_2d_kernel(with_packed=True, **kw), asdefault=True)
binary_2d_array_index_p.def_tags("event", "compact")

binary_2d_compact_only_p = XLACustomKernel("binary_2d_compact_only")
binary_2d_compact_only_p.def_kernel("b200", "gpu", lambda **kw: _compact_2d_kernel(with_packed=False, **kw), asdefault=True)
binary_2d_compact_only_p.def_tags("event", "compact")


def bitpack(arr, axis: int = 0, backend=None):
    """uint32 LSB-first words along `axis` (ref: _event/bitpack_binary.py:32-76).  The words are
    returned as int32 tensors (torch has no uint32 arithmetic); bit patterns are identical."""
    return bitpack_p(arr, axis=axis, backend=backend)


def binary_1d_array_index_p_call(spikes, backend=None):
    """-> (active_ids int32[n] ascending + zero tail, n_active int32[1]); ref: _event/compact.py:406."""
    return binary_1d_array_index_p(spikes, backend=backend)


def binary_2d_array_index_p_call(spikes, backend=None):
    """-> (packed[n, ceil(b/32)], active_ids, n_active); ref: _event/compact.py:589."""
    return binary_2d_array_index_p(spikes, backend=backend)


def binary_2d_compact_only_p_call(spikes, backend=None):
    """-> (active_ids, n_active); ref: _event/compact.py:258."""
    return binary_2d_compact_only_p(spikes, backend=backend)


# ---------------------------------------------------------------------------------------------
# containers
# ---------------------------------------------------------------------------------------------
class EventRepresentation:
    """ref: _event/base.py"""

    def __init__(self, value):
        self._value = value

    @property
    def value(self):
        return self._value

    @property
    def shape(self):
        return tuple(self._value.shape)

    @property
    def ndim(self):
        return self._value.ndim

    @property
    def dtype(self):
        return self._value.dtype

    def __repr__(self):
        return f"{type(self).__name__}(shape={self.shape}, dtype={self.dtype})"


def _is_dense_operand(x):
    return isinstance(x, torch.Tensor)


class BinaryArray(EventRepresentation):
    """Dense bool/float spikes; active = True / > 0 (ref: _event/binary.py:31-321)."""

    def __init__(self, value):
        super().__init__(as_tensor(value))

    @property
    def T(self):
        return BinaryArray(self._value.t())

    def __matmul__(self, oc):
        from ._dense import binary_densemm, binary_densemv
        if _is_dense_operand(oc):
            if self.ndim not in (1, 2):
                raise MathError(f"Matrix multiplication is only supported for 1D and 2D arrays. Got {self.ndim}D array.")
            assert oc.ndim == 2, f"Right operand must be a 2D array in matrix multiplication. Got {oc.ndim}D array."
            assert self.shape[-1] == oc.shape[0], (
                f"Incompatible dimensions for matrix multiplication: {self.shape[-1]} and {oc.shape[0]}.")
            if self.ndim == 1:
                return binary_densemv(oc, self._value, transpose=True)
            return binary_densemm(oc, self._value.t(), transpose=True).t()
        return oc.__rmatmul__(self)

    def __rmatmul__(self, oc):
        from ._dense import binary_densemm, binary_densemv
        if _is_dense_operand(oc):
            if self.ndim not in (1, 2):
                raise MathError(f"Matrix multiplication is only supported for 1D and 2D arrays. Got {self.ndim}D array.")
            assert oc.ndim == 2, f"Left operand must be a 2D array in matrix multiplication. Got {oc.ndim}D array."
            assert oc.shape[-1] == self.shape[0], (
                f"Incompatible dimensions for matrix multiplication: {oc.shape[-1]} and {self.shape[0]}.")
            if self.ndim == 1:
                return binary_densemv(oc, self._value, transpose=False)
            return binary_densemm(oc, self._value, transpose=False)
        return oc.__matmul__(self)


class BitPackedBinary(EventRepresentation):
    """Bit-packed spikes (ref: _event/bitpack_binary.py:78-339).

    `packed` is the kernel-facing encoding: 1-D -> ceil(n/32) words; 2-D [n, batch] -> packed along
    the batch axis, [n, ceil(batch/32)].  `value` is kept only when constructed from a dense array
    with keep_value=True (the reference always keeps it)."""

    def __init__(self, arr=None, *, packed=None, shape=None, keep_value=False):
        if packed is not None:
            assert shape is not None, "BitPackedBinary(packed=...) needs the original shape"
            self._packed = packed
            self._shape = tuple(shape)
            super().__init__(None)
        else:
            t = as_tensor(arr)
            self._shape = tuple(t.shape)
            self._packed = bitpack(t, axis=0 if t.ndim == 1 else 1)
            super().__init__(t if keep_value else None)

    @property
    def packed(self):
        return self._packed

    @property
    def original_shape(self):
        return self._shape

    @property
    def shape(self):
        return self._shape

    @property
    def ndim(self):
        return len(self._shape)

    @property
    def dtype(self):
        return torch.bool

    def to_dense(self):
        p = self._packed.to(torch.int64) & 0xFFFFFFFF
        bits = ((p.unsqueeze(-1) >> torch.arange(32, device=p.device)) & 1).to(torch.bool)
        if self.ndim == 1:
            return bits.reshape(-1)[: self._shape[0]]
        return bits.reshape(self._shape[0], -1)[:, : self._shape[1]]

    def __matmul__(self, oc):
        if _is_dense_operand(oc):
            return BinaryArray(self.to_dense()) @ oc
        return oc.__rmatmul__(self)

    def __rmatmul__(self, oc):
        if _is_dense_operand(oc):
            return oc @ BinaryArray(self.to_dense())
        return oc.__matmul__(self)


class CompactBinary(EventRepresentation):
    """packed words + ascending active ids + count (ref: _event/compact_binary.py:53-430)."""

    def __init__(self, packed, ids_and_count, shape, value=None):
        self._packed = packed
        self._ids_and_count = ids_and_count   # int32[n + 1]: ids then n_active
        self._shape = tuple(shape)
        super().__init__(value)

    @classmethod
    def from_array(cls, x, bit_width: int = 32, keep_value: bool = False):
        if bit_width != 32:
            raise ValueError(f"Only bit_width=32 is supported, got {bit_width}.")
        t = as_tensor(x)
        require_cuda(t)
        if t.ndim not in (1, 2):
            raise ValueError(f"CompactBinary supports 1D and 2D arrays, got {t.ndim}D.")
        s, kind = _spk_tensor(t)
        n = s.shape[0]
        b = s.shape[1] if s.ndim == 2 else 1
        buf = torch.zeros(n + 1, dtype=torch.int32, device=s.device)
        ws = workspace(s.device, n, b)
        L = _lib.lib()
        if s.ndim == 1:
            packed = torch.empty((n + 31) // 32, dtype=torch.int32, device=s.device)
            _lib.check(L.be_b200_compact_1d(ptr(s), kind, n, 1, ptr(buf), ctypes.c_void_p(buf.data_ptr() + 4 * n),
                                            ptr(packed), ptr(ws), ws.numel(), stream_ptr(s.device)), "compact_1d")
        else:
            packed = torch.empty((n, (b + 31) // 32), dtype=torch.int32, device=s.device)
            _lib.check(L.be_b200_compact_2d(ptr(s), kind, n, b, 1, ptr(packed), ptr(buf),
                                            ctypes.c_void_p(buf.data_ptr() + 4 * n), ptr(ws), ws.numel(),
                                            stream_ptr(s.device)), "compact_2d")
        return cls(packed, buf, t.shape, t if keep_value else None)

    @property
    def packed(self):
        return self._packed

    @property
    def active_ids(self):
        return self._ids_and_count[:-1]

    @property
    def n_active(self):
        return self._ids_and_count[-1:]

    @property
    def n_orig(self):
        return self._shape[0]

    @property
    def batch_size(self):
        return self._shape[1] if len(self._shape) == 2 else None

    @property
    def bit_width(self):
        return 32

    @property
    def shape(self):
        return self._shape

    @property
    def ndim(self):
        return len(self._shape)

    @property
    def dtype(self):
        return torch.bool

    def to_dense(self):
        return BitPackedBinary(packed=self._packed, shape=self._shape).to_dense()

    def __matmul__(self, oc):
        if _is_dense_operand(oc):
            return BinaryArray(self.to_dense()) @ oc
        return oc.__rmatmul__(self)

    def __rmatmul__(self, oc):
        if _is_dense_operand(oc):
            return oc @ BinaryArray(self.to_dense())
        return oc.__matmul__(self)
