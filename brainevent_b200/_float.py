"""Float (non-event) products through the same connectivity formats -- the callers either side of
the event path.

Mirrors ref: _csr/float.py (csrmv :49, csrmm :559, *_p_call :470, :1010), _fcn/float.py (fcnmv :33,
fcnmm :136), _jit_scalar/float.py (jitsmv :159, jitsmm :1024), _jit_normal/float.py (jitnmv, jitnmm),
_jit_uniform/float.py (jitumv, jitumm).  Same names, argument meaning and assertion behaviour; the
operand holds VALUES (cast to the weight dtype), not events.
"""
import torch

from . import _lib
from ._csr import _prep_structure
from ._jitc import (_initialize_conn_length, _initialize_seed, _param)
from ._op import XLACustomKernel, maybe_decimal, split_mantissa_unit
from ._runtime import as_tensor, ptr, require_cuda, stream_ptr, wt_code


def _operand(x, dtype, device):
    x = as_tensor(x, device=device)
    require_cuda(x)
    if x.dtype == torch.bool or not x.dtype.is_floating_point or x.dtype != dtype:
        x = x.to(dtype)
    return x.contiguous()


# ---------------------------------------------------------------------------------------------
# CSR
# ---------------------------------------------------------------------------------------------
def _csr_kernel(outs=None, shape=None, transpose=False, matmat=False, **kw):
    m, k = shape

    def kernel(data, indices, indptr, x):
        n_out = k if transpose else m
        dev = data.device
        homo = int(data.numel() == 1)
        L = _lib.lib()
        if matmat:
            n = x.shape[1]
            out = torch.empty((n_out, n), dtype=data.dtype, device=dev)
            rc = L.be_b200_csrmm(ptr(data), wt_code(data.dtype), homo, ptr(indices), ptr(indptr),
                                 int(indptr.dtype == torch.int64), ptr(x), m, k, indices.numel(), n,
                                 int(transpose), ptr(out), stream_ptr(dev))
            _lib.check(rc, "csrmm")
        else:
            out = torch.empty(n_out, dtype=data.dtype, device=dev)
            rc = L.be_b200_csrmv(ptr(data), wt_code(data.dtype), homo, ptr(indices), ptr(indptr),
                                 int(indptr.dtype == torch.int64), ptr(x), m, k, indices.numel(),
                                 int(transpose), ptr(out), stream_ptr(dev))
            _lib.check(rc, "csrmv")
        return [out]
    return kernel


csrmv_p = XLACustomKernel("csrmv")
csrmv_p.def_kernel("b200", "gpu", lambda **kw: _csr_kernel(matmat=False, **kw), asdefault=True)
csrmv_p.def_tags("csr", "float")
csrmm_p = XLACustomKernel("csrmm")
csrmm_p.def_kernel("b200", "gpu", lambda **kw: _csr_kernel(matmat=True, **kw), asdefault=True)
csrmm_p.def_tags("csr", "float")


def _csr_p_call(prim, data, indices, indptr, x, shape, transpose, backend, matmat):
    data, indices, indptr, _ = _prep_structure(data, indices, indptr)
    assert len(shape) == 2, "shape must be (rows, cols)"
    m, k = int(shape[0]), int(shape[1])
    assert indptr.ndim == 1 and indptr.shape[0] == m + 1, (
        f"indptr must have shape ({m + 1},), got {tuple(indptr.shape)}")
    assert indices.ndim == 1
    assert data.shape == (1,) or data.shape == indices.shape, (
        f"data must have shape (1,) or {tuple(indices.shape)}, got {tuple(data.shape)}")
    x = _operand(x, data.dtype, data.device)
    assert x.ndim == (2 if matmat else 1), f"operand must be {'2' if matmat else '1'}-D"
    n_src = m if transpose else k
    assert x.shape[0] == n_src, (f"shape mismatch: operand has {x.shape[0]} rows but the "
                                 f"{'transposed ' if transpose else ''}matrix expects {n_src}")
    return prim(data, indices, indptr, x, outs=None, shape=(m, k), transpose=bool(transpose), backend=backend)


def csrmv_p_call(data, indices, indptr, v, *, shape, transpose=False, backend=None):
    return _csr_p_call(csrmv_p, data, indices, indptr, v, shape, transpose, backend, False)


def csrmm_p_call(data, indices, indptr, B, *, shape, transpose=False, backend=None):
    return _csr_p_call(csrmm_p, data, indices, indptr, B, shape, transpose, backend, True)


def csrmv(data, indices, indptr, v, *, shape, transpose=False, backend=None):
    """y = A @ v (or A^T @ v) for a CSR matrix and a dense vector (ref: _csr/float.py:49)."""
    data, unitd = split_mantissa_unit(data)
    v, unitv = split_mantissa_unit(v)
    return maybe_decimal(csrmv_p_call(data, indices, indptr, v, shape=shape, transpose=transpose, backend=backend)[0],
                         unitd, unitv)


def csrmm(data, indices, indptr, B, *, shape, transpose=False, backend=None):
    """Y = A @ B (or A^T @ B), B [k or m, n] dense (ref: _csr/float.py:559)."""
    data, unitd = split_mantissa_unit(data)
    B, unitb = split_mantissa_unit(B)
    return maybe_decimal(csrmm_p_call(data, indices, indptr, B, shape=shape, transpose=transpose, backend=backend)[0],
                         unitd, unitb)


# ---------------------------------------------------------------------------------------------
# fixed number of connections (ELL)
# ---------------------------------------------------------------------------------------------
def _fcn_kernel(outs=None, shape=None, transpose=False, matmat=False, **kw):
    def kernel(weights, indices, x):
        n_rows, n_conn = indices.shape
        n_other = shape[1]
        n_out = n_other if transpose else n_rows
        dev = weights.device
        homo = int(weights.numel() == 1)
        L = _lib.lib()
        if matmat:
            n = x.shape[1]
            out = torch.empty((n_out, n), dtype=weights.dtype, device=dev)
            rc = L.be_b200_fcnmm(ptr(weights), wt_code(weights.dtype), homo, ptr(indices), n_rows, n_conn, ptr(x),
                                 n_other, n, int(transpose), ptr(out), stream_ptr(dev))
            _lib.check(rc, "fcnmm")
        else:
            out = torch.empty(n_out, dtype=weights.dtype, device=dev)
            rc = L.be_b200_fcnmv(ptr(weights), wt_code(weights.dtype), homo, ptr(indices), n_rows, n_conn, ptr(x),
                                 n_other, int(transpose), ptr(out), stream_ptr(dev))
            _lib.check(rc, "fcnmv")
        return [out]
    return kernel


fcnmv_p = XLACustomKernel("fcnmv")
fcnmv_p.def_kernel("b200", "gpu", lambda **kw: _fcn_kernel(matmat=False, **kw), asdefault=True)
fcnmv_p.def_tags("fcn", "float")
fcnmm_p = XLACustomKernel("fcnmm")
fcnmm_p.def_kernel("b200", "gpu", lambda **kw: _fcn_kernel(matmat=True, **kw), asdefault=True)
fcnmm_p.def_tags("fcn", "float")


def _fcn_p_call(prim, weights, indices, x, shape, transpose, backend, matmat):
    weights = as_tensor(weights)
    if weights.ndim == 0:
        weights = weights.reshape(1)
    indices = as_tensor(indices)
    require_cuda(weights, indices)
    assert weights.dtype.is_floating_point, "Weights must be a floating-point type."
    assert indices.dtype == torch.int32 and indices.ndim == 2, "indices must be int32 [n_rows, n_conn]"
    assert weights.numel() == 1 or tuple(weights.shape) == tuple(indices.shape), (
        f"weights must have shape (1,) or {tuple(indices.shape)}, got {tuple(weights.shape)}")
    n_pre, n_post = int(shape[0]), int(shape[1])
    assert indices.shape[0] == n_pre, f"Pre size mismatch, got {indices.shape[0]} != {n_pre}"
    x = _operand(x, weights.dtype, weights.device)
    assert x.ndim == (2 if matmat else 1), f"operand must be {'2' if matmat else '1'}-D"
    n_src = n_pre if transpose else n_post
    assert x.shape[0] == n_src, f"shape mismatch: operand has {x.shape[0]} rows, expected {n_src}"
    return prim(weights.contiguous(), indices.contiguous(), x, outs=None, shape=(n_pre, n_post),
                transpose=bool(transpose), backend=backend)


def fcnmv_p_call(weights, indices, vector, *, shape, transpose=False, backend=None):
    return _fcn_p_call(fcnmv_p, weights, indices, vector, shape, transpose, backend, False)


def fcnmm_p_call(weights, indices, matrix, *, shape, transpose=False, backend=None):
    return _fcn_p_call(fcnmm_p, weights, indices, matrix, shape, transpose, backend, True)


def fcnmv(weights, indices, vector, *, shape, transpose=False, backend=None):
    """transpose=False: y[i] = sum_k w[i,k] v[idx[i,k]];  transpose=True: y[idx[i,k]] += w[i,k] v[i]
    (ref: _fcn/float.py:33)."""
    weights, uw = split_mantissa_unit(weights)
    vector, uv = split_mantissa_unit(vector)
    return maybe_decimal(fcnmv_p_call(weights, indices, vector, shape=shape, transpose=transpose, backend=backend)[0], uv, uw)


def fcnmm(weights, indices, matrix, *, shape, transpose=False, backend=None):
    """ref: _fcn/float.py:136."""
    weights, uw = split_mantissa_unit(weights)
    matrix, um = split_mantissa_unit(matrix)
    return maybe_decimal(fcnmm_p_call(weights, indices, matrix, shape=shape, transpose=transpose, backend=backend)[0], um, uw)


# ---------------------------------------------------------------------------------------------
# just-in-time connectivity
# ---------------------------------------------------------------------------------------------
_KIND = {"scalar": _lib.JIT_SCALAR, "normal": _lib.JIT_NORMAL, "uniform": _lib.JIT_UNIFORM}


def _jit_kernel(outs=None, shape=None, transpose=False, corder=True, kind="scalar", matmat=False, **kw):
    from ._jitc import _normalize_chunk_size

    def kernel(p0, p1, clen, seed, x):
        out_len = shape[1] if transpose else shape[0]
        vec_len = shape[0] if transpose else shape[1]
        chunk = _normalize_chunk_size(shape[1])
        dev = x.device
        L = _lib.lib()
        if matmat:
            n = x.shape[1]
            out = torch.empty((out_len, n), dtype=p0.dtype, device=dev)
            rc = L.be_b200_jitmm(_KIND[kind], ptr(p0), ptr(p1), wt_code(p0.dtype), ptr(clen), ptr(seed), ptr(x),
                                 vec_len, out_len, n, chunk, int(corder), ptr(out), stream_ptr(dev))
        else:
            out = torch.empty(out_len, dtype=p0.dtype, device=dev)
            rc = L.be_b200_jitmv(_KIND[kind], ptr(p0), ptr(p1), wt_code(p0.dtype), ptr(clen), ptr(seed), ptr(x),
                                 vec_len, out_len, chunk, int(corder), ptr(out), stream_ptr(dev))
        _lib.check(rc, "jitmm" if matmat else "jitmv")
        return [out]
    return kernel


def _make_prim(name, kind, matmat):
    p = XLACustomKernel(name)
    p.def_kernel("b200", "gpu", lambda **kw: _jit_kernel(kind=kind, matmat=matmat, **kw), asdefault=True)
    p.def_tags("jit_" + kind, "float")
    return p


jitsmv_p = _make_prim("jitsmv", "scalar", False)
jitsmm_p = _make_prim("jitsmm", "scalar", True)
jitnmv_p = _make_prim("jitnmv", "normal", False)
jitnmm_p = _make_prim("jitnmm", "normal", True)
jitumv_p = _make_prim("jitumv", "uniform", False)
jitumm_p = _make_prim("jitumm", "uniform", True)


def _jit_p_call(prim, params, clen, x, seed, shape, transpose, corder, backend, matmat):
    x = as_tensor(x)
    require_cuda(x)
    dev = x.device
    assert len(shape) == 2, "The matrix shape should be a tuple of two integers."
    shape = (int(shape[0]), int(shape[1]))
    p0 = _param(params[0], dev, None)
    p1 = _param(params[1], dev, p0.dtype) if len(params) > 1 else None
    clen = as_tensor(clen, device=dev, dtype=torch.int32).reshape(-1)
    seed = as_tensor(seed, device=dev, dtype=torch.int32).reshape(-1)
    assert clen.shape == (1,), f"The clen shape should be (1,), but got {tuple(clen.shape)}."
    assert seed.shape == (1,), f"The seed shape should be (1,), but got {tuple(seed.shape)}."
    x = _operand(x, p0.dtype, dev)
    assert x.ndim == (2 if matmat else 1), "operand rank does not match the operator"
    if transpose:
        assert shape[0] == x.shape[0], f"The matrix shape and vector length do not match. {x.shape[0]} @ {shape}"
    else:
        assert shape[1] == x.shape[0], f"The matrix shape and vector length do not match. {shape} @ {x.shape[0]}"
    return prim(p0, p1, clen, seed, x, outs=None, shape=shape, transpose=bool(transpose), corder=bool(corder),
                backend=backend)


def _public(prim, n_params, matmat, doc):
    def op(*args, shape, transpose=False, corder=True, backend=None):
        params, (prob, x, seed) = args[:n_params], args[n_params:]
        vals, units = [], []
        for p in params:
            v, un = split_mantissa_unit(p)
            vals.append(v)
            units.append(un)
        x, unx = split_mantissa_unit(x)
        dev = as_tensor(x).device
        res = _jit_p_call(prim, vals, _initialize_conn_length(prob, dev), x, _initialize_seed(seed, dev), shape,
                          transpose, corder, backend, matmat)[0]
        return maybe_decimal(res, units[0], unx)
    op.__doc__ = doc
    return op


jitsmv = _public(jitsmv_p, 1, False, "jitsmv(weight, prob, vector, seed, *, shape, transpose, corder); ref: _jit_scalar/float.py:159")
jitsmm = _public(jitsmm_p, 1, True, "jitsmm(weight, prob, B, seed, *, shape, transpose, corder); ref: _jit_scalar/float.py:1024")
jitnmv = _public(jitnmv_p, 2, False, "jitnmv(w_loc, w_scale, prob, vector, seed, ...); ref: _jit_normal/float.py")
jitnmm = _public(jitnmm_p, 2, True, "jitnmm(w_loc, w_scale, prob, B, seed, ...); ref: _jit_normal/float.py")
jitumv = _public(jitumv_p, 2, False, "jitumv(w_low, w_high, prob, vector, seed, ...); ref: _jit_uniform/float.py")
jitumm = _public(jitumm_p, 2, True, "jitumm(w_low, w_high, prob, B, seed, ...); ref: _jit_uniform/float.py")

JIT_FLOAT_PRIMS = {"scalar": (jitsmv_p, jitsmm_p), "normal": (jitnmv_p, jitnmm_p), "uniform": (jitumv_p, jitumm_p)}
