"""Just-in-time connectivity event operators and matrix classes.

Mirrors ref: brainevent/_jit_scalar/binary.py (:44, :171), _jit_normal/binary.py (:44, :177),
_jit_uniform/binary.py (:44, :177) and the R/C classes in _jit_*/main.py.  The connectivity is
regenerated in-kernel from (seed, row, chunk, lane); `corder` picks gather (True) vs scatter
(False) exactly as in the reference (_jit_scalar/binary.py:438-439), `transpose` only fixes the
output shape, and chunk_size = ceil(shape[1]/4) of the *passed* shape (_misc.py:74-122).
"""
import math

import numpy as np
import torch

from . import _lib
from ._csr import _EVENT_TYPES
from ._event import BinaryArray
from ._op import XLACustomKernel, maybe_decimal, split_mantissa_unit
from ._runtime import EventArgs, as_tensor, event_args, ptr, stream_ptr, workspace, wt_code

MV_STRIDE, MM_STRIDE = 32, 4   # ref: _misc.py:37-38
_KIND = {"scalar": _lib.JIT_SCALAR, "normal": _lib.JIT_NORMAL, "uniform": _lib.JIT_UNIFORM}


def _normalize_chunk_size(n_cols, chunk_size=None, target_chunks=4):
    """ref: _misc.py:74-122"""
    if chunk_size is None:
        target_chunks = int(target_chunks)
        if target_chunks <= 0:
            raise ValueError("target_chunks must be positive")
        chunk_size = max(1, (int(n_cols) + target_chunks - 1) // target_chunks)
    chunk_size = int(chunk_size)
    if chunk_size <= 0:
        raise ValueError("chunk_size must be positive")
    return chunk_size


def _initialize_seed(seed=None, device="cuda"):
    """int32[1] on the device (ref: _data.py:1181-1210)."""
    if seed is None:
        seed = np.random.randint(0, int(1e8))
    if isinstance(seed, torch.Tensor):
        return seed.reshape(-1)[:1].to(device=device, dtype=torch.int32)
    return torch.tensor([int(np.asarray(seed).reshape(-1)[0])], dtype=torch.int64, device=device).to(torch.int32)


def _initialize_conn_length(conn_prob, device="cuda"):
    """clen = ceil(2 / prob) as int32[1] (ref: _data.py:1212-1245)."""
    clen = int(np.asarray(np.ceil(2.0 / float(conn_prob)), dtype=np.int32))
    return torch.tensor([clen], dtype=torch.int32, device=device)


def _jit_kernel(outs=None, shape=None, transpose=False, corder=True, kind="scalar", matmat=False,
                row_window=None, col_window=None, **kw):
    def kernel(p0, p1, clen, seed, ev):
        out_len = shape[1] if transpose else shape[0]
        vec_len = shape[0] if transpose else shape[1]
        chunk = _normalize_chunk_size(shape[1])
        dev = p0.device
        L = _lib.lib()
        if matmat:
            n = ev.batch
            out = torch.empty((out_len, n), dtype=p0.dtype, device=dev)
            ws = workspace(dev, ev.n, n)
            rc = L.be_b200_binary_jitmm(_KIND[kind], ptr(p0), ptr(p1), wt_code(p0.dtype), ptr(clen), ptr(seed),
                                        ptr(ev.spikes), ptr(ev.aux), ev.kind, vec_len, out_len, n, chunk,
                                        int(corder), ptr(out), ptr(ws), ws.numel(), stream_ptr(dev))
            _lib.check(rc, "binary_jitmm")
        else:
            r0, rn = row_window if row_window is not None else (0, -1)
            c0, c1 = col_window if col_window is not None else (0, -1)
            n_local = (rn if (corder and rn >= 0) else out_len) if corder else ((c1 - c0) if c1 >= 0 else out_len)
            out = torch.empty(n_local, dtype=p0.dtype, device=dev)
            ws = workspace(dev, ev.n, 1)
            rc = L.be_b200_binary_jitmv(_KIND[kind], ptr(p0), ptr(p1), wt_code(p0.dtype), ptr(clen), ptr(seed),
                                        ptr(ev.spikes), ptr(ev.aux), ev.kind, vec_len, out_len, chunk,
                                        int(corder), r0, rn, c0, c1, ptr(out), ptr(ws), ws.numel(),
                                        stream_ptr(dev))
            _lib.check(rc, "binary_jitmv")
        return [out]
    return kernel


def _make_prim(name, kind, matmat):
    p = XLACustomKernel(name)
    p.def_kernel("b200", "gpu", lambda **kw: _jit_kernel(kind=kind, matmat=matmat, **kw), asdefault=True)
    p.def_tags("jit_" + kind, "binary")
    return p


binary_jitsmv_p = _make_prim("binary_jitsmv", "scalar", False)
binary_jitsmm_p = _make_prim("binary_jitsmm", "scalar", True)
binary_jitnmv_p = _make_prim("binary_jitnmv", "normal", False)
binary_jitnmm_p = _make_prim("binary_jitnmm", "normal", True)
binary_jitumv_p = _make_prim("binary_jitumv", "uniform", False)
binary_jitumm_p = _make_prim("binary_jitumm", "uniform", True)


def _param(x, device, dtype):
    t = as_tensor(x, device=device)
    if not t.dtype.is_floating_point:
        t = t.to(dtype or torch.float32)
    elif dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.reshape(-1)[:1].contiguous()


def _jit_p_call(prim, params, clen, x, seed, shape, transpose, corder, backend, matmat, **extra):
    ev = x if isinstance(x, EventArgs) else event_args(x)
    dev = ev.device
    assert len(shape) == 2, "The matrix shape should be a tuple of two integers."
    shape = (int(shape[0]), int(shape[1]))
    p0 = _param(params[0], dev, None)
    p1 = _param(params[1], dev, p0.dtype) if len(params) > 1 else None
    clen = as_tensor(clen, device=dev, dtype=torch.int32).reshape(-1)
    seed = as_tensor(seed, device=dev, dtype=torch.int32).reshape(-1)
    assert clen.shape == (1,), f"The clen shape should be (1,), but got {tuple(clen.shape)}."
    assert seed.shape == (1,), f"The seed shape should be (1,), but got {tuple(seed.shape)}."
    assert (ev.batch is not None) == matmat, "event rank does not match the operator"
    if transpose:
        assert shape[0] == ev.n, f"The matrix shape and vector length do not match. {ev.n} @ {shape}"
    else:
        assert shape[1] == ev.n, f"The matrix shape and vector length do not match. {shape} @ {ev.n}"
    return prim(p0, p1, clen, seed, ev, outs=None, shape=shape, transpose=bool(transpose),
                corder=bool(corder), backend=backend, **extra)


def _public(prim, n_params, matmat, doc):
    def op(*args, shape, transpose=False, corder=True, backend=None, **extra):
        params, (prob, x, seed) = args[:n_params], args[n_params:]
        units = []
        vals = []
        for p in params:
            v, un = split_mantissa_unit(p)
            vals.append(v)
            units.append(un)
        x, unx = split_mantissa_unit(x)
        dev = event_args(x).device
        seed = _initialize_seed(seed, dev)
        clen = _initialize_conn_length(prob, dev)
        call = lambda ps: _jit_p_call(prim, ps, clen, x, seed, shape, transpose, corder, backend, matmat, **extra)[0]
        if torch.is_grad_enabled() and any(isinstance(v, torch.Tensor) and v.requires_grad for v in vals):
            from . import _autograd           # the operator is linear in its weight parameters (ref: the primitives' jvp rules)
            res = _autograd.linear_in_params(call, [as_tensor(v, device=dev) for v in vals])
        else:
            res = call(vals)
        return maybe_decimal(res, units[0], unx)
    op.__doc__ = doc
    return op


binary_jitsmv = _public(binary_jitsmv_p, 1, False, "binary_jitsmv(weight, prob, vector, seed, *, shape, transpose, corder); ref: _jit_scalar/binary.py:44")
binary_jitsmm = _public(binary_jitsmm_p, 1, True, "binary_jitsmm(weight, prob, B, seed, *, shape, transpose, corder); ref: _jit_scalar/binary.py:171")
binary_jitnmv = _public(binary_jitnmv_p, 2, False, "binary_jitnmv(w_loc, w_scale, prob, vector, seed, ...); ref: _jit_normal/binary.py:44")
binary_jitnmm = _public(binary_jitnmm_p, 2, True, "binary_jitnmm(w_loc, w_scale, prob, B, seed, ...); ref: _jit_normal/binary.py:177")
binary_jitumv = _public(binary_jitumv_p, 2, False, "binary_jitumv(w_low, w_high, prob, vector, seed, ...); ref: _jit_uniform/binary.py:44")
binary_jitumm = _public(binary_jitumm_p, 2, True, "binary_jitumm(w_low, w_high, prob, B, seed, ...); ref: _jit_uniform/binary.py:177")


def jitc_row_connections(prob, seed, rows, k, chunk_size, stride=MV_STRIDE, device="cuda", cap=None):
    """Connectivity dump for parity: (counts int32[n_rows], cols int32[n_rows, cap]) of generated
    rows `rows` = (begin, n) in walk order.  ref: _jit_scalar/csr.py (count/fill)."""
    begin, n_rows = rows
    clen = _initialize_conn_length(prob, device)
    seed_t = _initialize_seed(seed, device)
    if cap is None:
        cap = int(2.5 * k * float(prob)) + 64
    counts = torch.zeros(n_rows, dtype=torch.int32, device=device)
    cols = torch.full((n_rows, cap), -1, dtype=torch.int32, device=device)
    _lib.check(_lib.lib().be_b200_jitc_row_connections(ptr(clen), ptr(seed_t), begin, n_rows, k, chunk_size,
                                                       stride, ptr(counts), ptr(cols), cap,
                                                       stream_ptr(torch.device(device))), "jitc_row_connections")
    return counts, cols


# ---------------------------------------------------------------------------------------------
# JITC -> CSR materialisation (ref: _jit_scalar/csr.py:488 jits_to_csr, _jit_normal/csr.py:465,
# _jit_uniform/csr.py)
# ---------------------------------------------------------------------------------------------
MM_STRIDE = 4


def _normalize_matrix_mode(matrix_mode):
    if matrix_mode not in ("mv", "mm"):
        raise ValueError(f"matrix_mode must be 'mv' or 'mm', got {matrix_mode!r}")
    return matrix_mode


def _jit_to_csr(kind, params, prob, seed, shape, corder, matrix_mode, device):
    from ._csr import CSR
    n_rows, n_cols = int(shape[0]), int(shape[1])
    stride = MV_STRIDE if _normalize_matrix_mode(matrix_mode) == "mv" else MM_STRIDE
    units = [split_mantissa_unit(p)[1] for p in params]
    p0 = _param(split_mantissa_unit(params[0])[0], device, None)
    p1 = _param(split_mantissa_unit(params[1])[0], device, p0.dtype) if len(params) > 1 else None
    dev = p0.device
    if n_rows == 0 or n_cols == 0 or float(prob) == 0.0:
        return CSR((torch.zeros(0, dtype=p0.dtype, device=dev), torch.zeros(0, dtype=torch.int32, device=dev),
                    torch.zeros(n_rows + 1, dtype=torch.int32, device=dev)), shape=(n_rows, n_cols))
    clen, seed_t = _initialize_conn_length(prob, dev), _initialize_seed(seed, dev)
    chunk = _normalize_chunk_size(n_cols)
    L = _lib.lib()
    ws = torch.empty(L.be_b200_jitc_csr_workspace_bytes(n_rows, 0), dtype=torch.uint8, device=dev)
    indptr = torch.empty(n_rows + 1, dtype=torch.int64, device=dev)
    _lib.check(L.be_b200_jitc_csr_count(ptr(clen), ptr(seed_t), n_rows, n_cols, chunk, stride, int(bool(corder)),
                                        ptr(indptr), ptr(ws), ws.numel(), stream_ptr(dev)), "jitc_csr_count")
    nnz = int(indptr[-1].item())      # the one host read-back, as in the reference (csr.py:530)
    indices = torch.empty(nnz, dtype=torch.int32, device=dev)
    data = torch.empty(nnz, dtype=p0.dtype, device=dev)
    if nnz:
        ws = torch.empty(L.be_b200_jitc_csr_workspace_bytes(n_rows, nnz), dtype=torch.uint8, device=dev)
        _lib.check(L.be_b200_jitc_csr_fill(_KIND[kind], ptr(p0), ptr(p1), wt_code(p0.dtype), ptr(clen), ptr(seed_t),
                                           n_rows, n_cols, chunk, stride, int(bool(corder)), ptr(indptr), nnz,
                                           ptr(indices), ptr(data), ptr(ws), ws.numel(), stream_ptr(dev)),
                   "jitc_csr_fill")
    if nnz < 2 ** 31:
        indptr = indptr.to(torch.int32)
    return CSR((maybe_decimal(data, units[0]), indices, indptr), shape=(n_rows, n_cols))


def jits_to_csr(weight, prob, seed, *, shape, corder, matrix_mode, backend=None, device="cuda"):
    """CSR of the matrix jitsmv ('mv') / jitsmm ('mm') applies with transpose=False
    (ref: _jit_scalar/csr.py:488)."""
    return _jit_to_csr("scalar", (weight,), prob, seed, shape, corder, matrix_mode, device)


def jitn_to_csr(w_loc, w_scale, prob, seed, *, shape, corder, matrix_mode, backend=None, device="cuda"):
    """ref: _jit_normal/csr.py:465."""
    return _jit_to_csr("normal", (w_loc, w_scale), prob, seed, shape, corder, matrix_mode, device)


def jitu_to_csr(w_low, w_high, prob, seed, *, shape, corder, matrix_mode, backend=None, device="cuda"):
    """ref: _jit_uniform/csr.py."""
    return _jit_to_csr("uniform", (w_low, w_high), prob, seed, shape, corder, matrix_mode, device)


class _ModeView:
    """`mat.mv` / `mat.mm`: the mv (stride 32) and mm (stride 4) kernels draw different matrices,
    so materialisation names the one it wants (ref: _jit_scalar/main.py:40-86)."""
    __slots__ = ("_mat", "_mode")

    def __init__(self, mat, mode):
        self._mat, self._mode = mat, mode

    def _generated(self, device):
        m = self._mat
        gen_shape, gen_transpose = m._materialize_params()
        # the products of a column-oriented class run the generator with transpose=True: its rows
        # are then the generated matrix's COLUMNS for corder=True (and its rows for corder=False),
        # i.e. the materialisation with the opposite corder
        corder = (not m.corder) if gen_transpose else m.corder
        g = _jit_to_csr(type(m)._family, m.params, m.prob, m.seed, gen_shape, corder, self._mode, device)
        return g, gen_transpose

    def tocsr(self, device="cuda"):
        from ._csr import CSR, csr_to_csc_index
        g, gen_transpose = self._generated(device)
        if not gen_transpose:
            return g
        # column-oriented classes generate the transpose: CSR(M) = CSC mirror of the generated CSR
        indptr, indices, perm = csr_to_csc_index(g.indptr, g.indices, shape=g.shape)
        return CSR((g.data[perm.long()], indices, indptr), shape=self._mat.shape)

    def tocsc(self, device="cuda"):
        from ._csr import CSC, csr_to_csc_index
        g, gen_transpose = self._generated(device)
        if gen_transpose:
            return g.T          # the generated CSR compresses the columns of the logical matrix
        indptr, indices, perm = csr_to_csc_index(g.indptr, g.indices, shape=g.shape)
        return CSC((g.data[perm.long()], indices, indptr), shape=g.shape)

    def todense(self, device="cuda"):
        g, gen_transpose = self._generated(device)
        return g.todense().t() if gen_transpose else g.todense()

    def __repr__(self):
        return f"{type(self._mat).__name__}.{self._mode}"


# ---------------------------------------------------------------------------------------------
# matrix classes
# ---------------------------------------------------------------------------------------------
class JITCMatrix:
    """Base of the just-in-time connectivity matrices (ref: _data.py JITCMatrix)."""
    _prims = None

    def __init__(self, params, prob, seed, *, shape, corder=False, backend=None):
        self.params = tuple(params)
        self.prob = float(prob)
        self.seed = seed
        self.shape = (int(shape[0]), int(shape[1]))
        self.corder = bool(corder)
        self.backend = backend

        self._dev = {}

    @property
    def mv(self):
        return _ModeView(self, "mv")

    @property
    def mm(self):
        return _ModeView(self, "mm")

    def tocsr(self):
        raise NotImplementedError("tocsr() is ambiguous for JITC matrices: the mv and mm kernels draw different "
                                  "matrices. Use mat.mv.tocsr() or mat.mm.tocsr().")

    todense = tocsr

    def _device_params(self, device):
        """weight parameters, clen and seed as one-element device tensors, created once per device
        (the reference keeps them as jax arrays); keeps the per-call path free of host->device
        copies so it can be captured in a CUDA graph."""
        key = str(device)
        if key not in self._dev:
            vals = [_param(split_mantissa_unit(p)[0], device, None) for p in self.params]
            vals = [v.to(vals[0].dtype) for v in vals]
            self._dev[key] = (vals, _initialize_conn_length(self.prob, device), _initialize_seed(self.seed, device))
        return self._dev[key]

    def _call(self, other, shape, transpose, corder):
        if other.ndim not in (1, 2):
            raise NotImplementedError(f"matmul with object of shape {other.shape}")
        return other.ndim == 2

    def _run(self, x, shape, transpose, corder, matmat):
        if not isinstance(x, _EVENT_TYPES):
            return self._run_float(x, shape, transpose, corder, matmat)
        prim = type(self)._prims[1 if matmat else 0]
        ev = event_args(x)
        vals, clen, seed = self._device_params(ev.device)
        res = _jit_p_call(prim, vals, clen, ev, seed, shape, transpose, corder, self.backend, matmat)[0]
        return maybe_decimal(res, split_mantissa_unit(self.params[0])[1])


    def _run_float(self, x, shape, transpose, corder, matmat):
        """plain (non-event) operand: jit?mv / jit?mm (ref: _jit_scalar/main.py:941-1066)."""
        from . import _float
        prim = _float.JIT_FLOAT_PRIMS[type(self)._family][1 if matmat else 0]
        x = as_tensor(x)
        vals, clen, seed = self._device_params(x.device)
        res = _float._jit_p_call(prim, vals, clen, x, seed, shape, transpose, corder, self.backend, matmat)[0]
        return maybe_decimal(res, split_mantissa_unit(self.params[0])[1])


class _JITC_R(JITCMatrix):
    """Row-oriented: the generated matrix has the logical shape (ref: _jit_scalar/main.py:885-1066)."""

    def _materialize_params(self):
        return self.shape, False

    def __matmul__(self, other):
        return self._run(other, self.shape, False, self.corder, self._call(other, None, None, None))

    def __rmatmul__(self, other):
        if self._call(other, None, None, None):
            if not isinstance(other, _EVENT_TYPES):
                return self._run(as_tensor(other).t(), self.shape, True, not self.corder, True).t()
            if not isinstance(other, BinaryArray):
                raise NotImplementedError("2-D packed events must be packed along the batch axis")
            return self._run(BinaryArray(other.value.t()), self.shape, True, not self.corder, True).t()
        return self._run(other, self.shape, True, not self.corder, False)


class _JITC_C(JITCMatrix):
    """Column-oriented: the generator is invoked with shape[::-1] (ref: _jit_scalar/main.py:1392-1593),
    so R and C of the same logical shape draw different matrices."""

    def _materialize_params(self):
        return self.shape[::-1], True

    def __matmul__(self, other):
        return self._run(other, self.shape[::-1], True, self.corder, self._call(other, None, None, None))

    def __rmatmul__(self, other):
        if self._call(other, None, None, None):
            if not isinstance(other, _EVENT_TYPES):
                return self._run(as_tensor(other).t(), self.shape[::-1], False, not self.corder, True).t()
            if not isinstance(other, BinaryArray):
                raise NotImplementedError("2-D packed events must be packed along the batch axis")
            return self._run(BinaryArray(other.value.t()), self.shape[::-1], False, not self.corder, True).t()
        return self._run(other, self.shape[::-1], False, not self.corder, False)


def _cls(name, base, prims, n_params, doc):
    def __init__(self, data, *, shape, corder=False, backend=None):
        *params, prob, seed = data
        assert len(params) == n_params, f"{name} expects ({', '.join(['w'] * n_params)}, prob, seed)"
        base.__init__(self, params, prob, seed, shape=shape, corder=corder, backend=backend)
    family = {"S": "scalar", "N": "normal", "U": "uniform"}[name[4]]
    return type(name, (base,), {"__init__": __init__, "_prims": prims, "_family": family, "__doc__": doc})


JITCScalarR = _cls("JITCScalarR", _JITC_R, (binary_jitsmv_p, binary_jitsmm_p), 1, "JITCScalarR((weight, prob, seed), shape=...)")
JITCScalarC = _cls("JITCScalarC", _JITC_C, (binary_jitsmv_p, binary_jitsmm_p), 1, "JITCScalarC((weight, prob, seed), shape=...)")
JITCNormalR = _cls("JITCNormalR", _JITC_R, (binary_jitnmv_p, binary_jitnmm_p), 2, "JITCNormalR((loc, scale, prob, seed), shape=...)")
JITCNormalC = _cls("JITCNormalC", _JITC_C, (binary_jitnmv_p, binary_jitnmm_p), 2, "JITCNormalC((loc, scale, prob, seed), shape=...)")
JITCUniformR = _cls("JITCUniformR", _JITC_R, (binary_jitumv_p, binary_jitumm_p), 2, "JITCUniformR((low, high, prob, seed), shape=...)")
JITCUniformC = _cls("JITCUniformC", _JITC_C, (binary_jitumv_p, binary_jitumm_p), 2, "JITCUniformC((low, high, prob, seed), shape=...)")
