"""Dense event operators (ref: brainevent/_dense/binary.py: binary_densemv :79, binary_densemm :487)."""
import torch

from . import _lib
from ._op import XLACustomKernel, maybe_decimal, split_mantissa_unit
from ._runtime import as_tensor, event_args, ptr, require_cuda, stream_ptr, workspace, wt_code


def _dense_kernel(outs=None, transpose=False, matmat=False, **kw):
    def kernel(W, ev):
        rows, cols = W.shape
        dev = W.device
        L = _lib.lib()
        if matmat:
            n = ev.batch
            m = cols if transpose else rows
            out = torch.empty((m, n), dtype=W.dtype, device=dev)
            ws = workspace(dev, ev.n, n)
            rc = L.be_b200_binary_densemm(ptr(W), wt_code(W.dtype), rows, cols, ptr(ev.spikes), ptr(ev.aux),
                                          ev.kind, n, int(transpose), ptr(out), ptr(ws), ws.numel(),
                                          stream_ptr(dev))
            _lib.check(rc, "binary_densemm")
        else:
            out = torch.empty(cols if transpose else rows, dtype=W.dtype, device=dev)
            ws = workspace(dev, ev.n, 1)
            rc = L.be_b200_binary_densemv(ptr(W), wt_code(W.dtype), rows, cols, ptr(ev.spikes), ptr(ev.aux),
                                          ev.kind, int(transpose), ptr(out), ptr(ws), ws.numel(),
                                          stream_ptr(dev))
            _lib.check(rc, "binary_densemv")
        return [out]
    return kernel


binary_densemv_p = XLACustomKernel("binary_densemv")
binary_densemv_p.def_kernel("b200", "gpu", lambda **kw: _dense_kernel(matmat=False, **kw), asdefault=True)
binary_densemv_p.def_tags("dense", "binary")
binary_densemm_p = XLACustomKernel("binary_densemm")
binary_densemm_p.def_kernel("b200", "gpu", lambda **kw: _dense_kernel(matmat=True, **kw), asdefault=True)
binary_densemm_p.def_tags("dense", "binary")


def _dense_p_call(prim, weights, spikes, transpose, backend, matmat):
    W = as_tensor(weights)
    require_cuda(W)
    assert W.ndim == 2, f"weights must be 2-D, got {W.ndim}-D"
    assert W.dtype.is_floating_point, "Weights must be a floating-point type."
    ev = event_args(spikes)
    assert (ev.batch is not None) == matmat
    k = W.shape[0] if transpose else W.shape[1]
    assert ev.n == k, f"Incompatible dimensions: weights {tuple(W.shape)} (transpose={transpose}) vs {ev.n} events"
    return prim(W.contiguous(), ev, outs=None, transpose=bool(transpose), backend=backend)


def binary_densemv_p_call(weights, spikes, *, transpose, backend=None):
    return _dense_p_call(binary_densemv_p, weights, spikes, transpose, backend, False)


def binary_densemm_p_call(weights, spikes, *, transpose, backend=None):
    return _dense_p_call(binary_densemm_p, weights, spikes, transpose, backend, True)


def binary_densemv(weights, spikes, *, transpose, backend=None):
    """weights[m,k] @ spikes[k] (transpose=False) or spikes[k] @ weights[k,n] (transpose=True)."""
    weights, wunit = split_mantissa_unit(weights)
    spikes, sunit = split_mantissa_unit(spikes)
    from . import _autograd
    if _autograd.needs_grad(weights, spikes):
        r = _autograd.densemv(weights, spikes, transpose, backend)
    else:
        r = binary_densemv_p_call(weights, spikes, transpose=transpose, backend=backend)[0]
    return maybe_decimal(r, wunit, sunit)


def binary_densemm(weights, spikes, *, transpose, backend=None):
    """weights[k,m].T @ spikes[k,n] (transpose=True) or weights[m,k] @ spikes[k,n] -> [m, n]."""
    weights, wunit = split_mantissa_unit(weights)
    spikes, sunit = split_mantissa_unit(spikes)
    r = binary_densemm_p_call(weights, spikes, transpose=transpose, backend=backend)[0]
    return maybe_decimal(r, wunit, sunit)
