"""Plumbing between torch CUDA tensors and the C ABI: dtype codes, stream handles, workspaces,
event-argument normalisation.  torch is used for device memory and streams only."""
import ctypes

import torch

from . import _lib

_WT = {torch.float32: _lib.F32, torch.float64: _lib.F64, torch.float16: _lib.F16,
       torch.bfloat16: _lib.BF16}


def wt_code(dtype) -> int:
    try:
        return _WT[dtype]
    except KeyError:
        raise TypeError(f"weights must be float32/float64/float16/bfloat16, got {dtype}") from None


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    """Handle of the current stream of `device`; also tells the library binding which device the call that
    takes this handle must run on (see _lib.note_device)."""
    if device is not None:
        idx = torch.device(device).index
        if idx is not None:
            _lib.note_device(idx)
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.KernelExecutionError(
                "brainevent_b200 operators run on CUDA tensors only (no CPU fallback); "
                f"got a tensor on {t.device}")


def as_tensor(x, device=None, dtype=None):
    if isinstance(x, torch.Tensor):
        t = x
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        if device is not None and t.device != torch.device(device):
            t = t.to(device)
        return t
    return torch.as_tensor(x, dtype=dtype, device=device)


# -------------------------------------------------------------------------------------------
# workspaces: one zero-initialised, self-resetting buffer per (device, stream)
# -------------------------------------------------------------------------------------------
_WS = {}


def workspace(device, n_src: int, n_batch: int = 1):
    """The carve-up of a workspace depends on (n_src, n_batch) and its header / tile-status words
    must be zero on first use, so buffers are cached per layout and never shared across layouts."""
    dev = torch.device(device)
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream, int(n_src), int(n_batch))
    buf = _WS.get(key)
    if buf is None:
        need = int(_lib.lib().be_b200_workspace_bytes(int(n_src), int(n_batch)))
        buf = torch.zeros(need, dtype=torch.uint8, device=dev)
        _WS[key] = buf
    return buf


def csc_workspace(device, m, n_post, nse):
    need = int(_lib.lib().be_b200_csc_workspace_bytes(int(m), int(n_post), int(nse)))
    return torch.empty(need, dtype=torch.uint8, device=device)


# -------------------------------------------------------------------------------------------
# events
# -------------------------------------------------------------------------------------------
class EventArgs:
    """(spikes pointer, aux pointer, kind) + the tensors that must outlive the launch."""
    __slots__ = ("spikes", "aux", "kind", "keep", "n", "batch", "device")


def event_args(v, n_expected=None):
    """Normalise BinaryArray / BitPackedBinary / CompactBinary / raw tensors."""
    from ._event import BinaryArray, BitPackedBinary, CompactBinary
    ea = EventArgs()
    if isinstance(v, CompactBinary):
        ea.spikes, ea.aux, ea.kind = v.packed, v._ids_and_count, _lib.SPK_COMPACT
        ea.n, ea.batch = v.shape[0], (v.shape[1] if len(v.shape) == 2 else None)
        ea.keep = (v.packed, v._ids_and_count)
    elif isinstance(v, BitPackedBinary):
        ea.spikes, ea.aux, ea.kind = v.packed, None, _lib.SPK_PACKED
        ea.n, ea.batch = v.shape[0], (v.shape[1] if len(v.shape) == 2 else None)
        ea.keep = (v.packed,)
    else:
        if isinstance(v, BinaryArray):
            v = v.value
        t = as_tensor(v)
        require_cuda(t)
        if t.dtype in (torch.bool, torch.uint8, torch.int8):
            kind = _lib.SPK_BOOL
        elif t.dtype == torch.float32:
            kind = _lib.SPK_F32
        elif t.dtype.is_floating_point:
            t, kind = (t > 0), _lib.SPK_BOOL     # other float widths: apply the `> 0` predicate first
        else:
            t, kind = (t != 0), _lib.SPK_BOOL    # integer spikes are cast to bool (ref: _dense/binary.py)
        t = t.contiguous()
        ea.spikes, ea.aux, ea.kind = t, None, kind
        ea.n, ea.batch = t.shape[0], (t.shape[1] if t.ndim == 2 else None)
        ea.keep = (t,)
    ea.device = ea.spikes.device
    if n_expected is not None and ea.n != n_expected:
        raise AssertionError(f"event length {ea.n} does not match the matrix dimension {n_expected}")
    return ea
