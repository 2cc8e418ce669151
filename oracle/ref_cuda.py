"""ctypes loader for the reference's OWN CUDA kernels built by `oracle/ref_cuda/Makefile` into
`oracle/_ref/libref_<op>.so` — TEST / BENCH INFRASTRUCTURE ONLY.

Nothing under `brainevent_b200/` imports this module.  It exists so that (1) the GPU parity tests
have a second oracle that is the reference itself (modulo atomic ordering), and (2)
`tests/bench_vs_reference_cuda.py` can time "the reference's kernels recompiled for sm_100a" beside
the new ones on the same B200, same inputs.  Each function mirrors the Python side of the
reference's kernel generator for that op (cited per function): same operands, same attributes.
"""
import ctypes
import math
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "_ref")

# BE::DType codes, /root/reference/brainevent/include/brainevent/tensor.h:42-59
_DT = {torch.float16: 0, torch.float32: 1, torch.float64: 2, torch.bfloat16: 3, torch.int8: 4,
       torch.int16: 5, torch.int32: 6, torch.int64: 7, torch.uint8: 8, torch.bool: 12}
_U32 = 10


class _RefTensor(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("shape", ctypes.c_int64 * 4), ("ndim", ctypes.c_int32),
                ("dtype", ctypes.c_int32)]


_libs = {}


def available(name: str) -> bool:
    return os.path.exists(os.path.join(_DIR, f"libref_{name}.so"))


def _lib(name):
    if name not in _libs:
        path = os.path.join(_DIR, f"libref_{name}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle/ref_cuda` where /root/reference exists")
        _libs[name] = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
    return _libs[name]


def _desc(t, dtype_code=None):
    assert t.is_cuda and t.is_contiguous() and t.dim() <= 4
    d = _RefTensor()
    d.data = t.data_ptr()
    for i, s in enumerate(t.shape):
        d.shape[i] = s
    d.ndim = t.dim()
    d.dtype = _DT[t.dtype] if dtype_code is None else dtype_code
    return d


def _call(lib, fn, tensors, attrs=()):
    f = getattr(_lib(lib), "ref_" + fn)
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
    arr = (_RefTensor * len(tensors))(*tensors)
    at = (ctypes.c_int64 * max(1, len(attrs)))(*attrs)
    rc = f(ctypes.cast(arr, ctypes.c_void_p), ctypes.cast(at, ctypes.c_void_p),
           torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise RuntimeError(f"reference CUDA entry {fn} failed")


def _spk8(v):
    """bool spikes travel as one byte each (XLA PRED / int8), `_fcn/binary.py:256-288`."""
    return v.view(torch.uint8) if v.dtype == torch.bool else v


# ---------------------------------------------------------------- FCN scatter (`_fcn/binary.py:256-288`)
def binary_fcnmv_scatter(weights, indices, spikes, n_post, out=None):
    out = torch.empty(n_post, dtype=torch.float32, device=indices.device) if out is None else out
    kind = "homo" if weights.numel() == 1 else "hetero"
    _call("fcnmv", f"binary_fcnmv_scatter_{kind}_bool_f32",
          [_desc(weights), _desc(indices), _desc(_spk8(spikes)), _desc(out)])
    return out


# ---------------------------------------------------------------- CSR NT gather (`_csr/binary.py:622-655`)
def binary_csrmv_nt(weights, indices, indptr, vector, n_rows, out=None):
    out = torch.empty(n_rows, dtype=torch.float32, device=indices.device) if out is None else out
    kind = "homo" if weights.numel() == 1 else "hetero"
    _call("csrmv", f"binary_csrmv_nt_auto_{kind}_f32_bool",
          [_desc(weights), _desc(indices), _desc(indptr), _desc(_spk8(vector)), _desc(out)])
    return out


# ------------------------------------- CSR T scatter, the GPU default (`_csr/binary.py:564-621`, hybrid)
def hybrid_task_capacity(indptr, tpr_threshold=128, task_nnz=4096):
    """`_csr/hybrid_config.py:298-323`: sum over rows longer than the threshold of ceil(nnz/task_nnz)."""
    nnz = (indptr[1:] - indptr[:-1]).to(torch.int64)
    big = nnz[nnz > tpr_threshold]
    return int(((big + task_nnz - 1) // task_nnz).sum().item())


class HybridWorkspace:
    """`_make_binary_csrmv_workspace`, `_csr/binary.py:104-116`."""

    def __init__(self, indptr):
        self.capacity = hybrid_task_capacity(indptr)
        self.task_begin = torch.empty(self.capacity, dtype=indptr.dtype, device=indptr.device)
        self.task_end = torch.empty(self.capacity, dtype=indptr.dtype, device=indptr.device)
        self.status = torch.empty(2, dtype=torch.int32, device=indptr.device)


def binary_csrmv_t_hybrid(weights, indices, indptr, vector, n_post, ws, out=None):
    out = torch.empty(n_post, dtype=torch.float32, device=indices.device) if out is None else out
    kind = "homo" if weights.numel() == 1 else "hetero"
    _call("csrmv_hybrid", f"binary_csrmv_wat_hybrid_{kind}_f32_bool",
          [_desc(weights), _desc(indices), _desc(indptr), _desc(_spk8(vector), _DT[torch.bool]),
           _desc(ws.task_begin), _desc(ws.task_end), _desc(ws.status), _desc(out)], [ws.capacity])
    return out


def binary_csrmm_t_hybrid(weights, indices, indptr, matrix, n_post, ws, out=None):
    """matrix `[n_pre, n_batch]`; physical output `[n_batch, n_post]` (`_csr/binary.py:1263-1286`)."""
    nb = matrix.shape[1]
    out = torch.empty(nb, n_post, dtype=torch.float32, device=indices.device) if out is None else out
    kind = "homo" if weights.numel() == 1 else "hetero"
    _call("csrmm_hybrid", f"binary_csrmm_sraw_hybrid_{kind}_f32_bool",
          [_desc(weights), _desc(indices), _desc(indptr), _desc(_spk8(matrix), _DT[torch.bool]),
           _desc(ws.task_begin), _desc(ws.task_end), _desc(ws.status), _desc(out)], [ws.capacity])
    return out


# ---------------------------------------------------------------- JITC event matvec
_scalars = {}


def _i32(value, dev):
    """int32[1] device scalar, cached so repeated calls (and CUDA-graph capture) issue no host copy."""
    key = (int(value), str(dev))
    if key not in _scalars:
        _scalars[key] = torch.tensor([int(value)], dtype=torch.int32, device=dev)
    return _scalars[key]


def _jit_common(lib, params, prob, vector, seed, shape, transpose, corder, out):
    """`_jit_scalar/binary.py:424-460` (same in `_jit_normal`, `_jit_uniform`): pack_bool, then
    `notrans` (corder) or `trans` with attrs vector_size = k, chunk_size = ceil(shape[1]/4)."""
    dev = vector.device
    k = vector.shape[0]
    n_out = shape[1] if transpose else shape[0]
    out = torch.empty(n_out, dtype=torch.float32, device=dev) if out is None else out
    clen = _i32(math.ceil(2.0 / prob), dev)                                          # `_data.py:1212-1245`
    seed_t = _i32(seed, dev)
    packed = torch.empty((k + 31) // 32, dtype=torch.int32, device=dev)
    spk = _spk8(vector)
    _call(lib, "pack_bool", [_desc(spk, _DT[torch.int8]), _desc(packed, _U32)])
    chunk = max(1, -(-int(shape[1]) // 4))                                           # `_misc.py:74-122`
    _call(lib, "notrans_f32" if corder else "trans_f32",
          [_desc(p) for p in params] + [_desc(clen), _desc(seed_t), _desc(packed, _U32), _desc(out)], [k, chunk])
    return out


def binary_jitsmv(weight, prob, vector, seed, *, shape, transpose, corder, out=None):
    return _jit_common("jitsmv", [weight], prob, vector, seed, shape, transpose, corder, out)


def binary_jitnmv(w_loc, w_scale, prob, vector, seed, *, shape, transpose, corder, out=None):
    return _jit_common("jitnmv", [w_loc, w_scale], prob, vector, seed, shape, transpose, corder, out)


def binary_jitumv(w_low, w_high, prob, vector, seed, *, shape, transpose, corder, out=None):
    return _jit_common("jitumv", [w_low, w_high], prob, vector, seed, shape, transpose, corder, out)


# ---------------------------------------------------------------- dense event matmul (`_dense/binary.py:933-990`)
def binary_densemm_transpose(weights, spikes, out=None):
    """weights `[k, m]`, spikes `[k, n_batch]` -> physical `[n_batch, m]`."""
    k, m = weights.shape
    out = torch.empty(spikes.shape[1], m, dtype=torch.float32, device=weights.device) if out is None else out
    _call("densemm", "binary_densemm_transpose_f32_bool", [_desc(weights), _desc(_spk8(spikes)), _desc(out)])
    return out


# ---------------------------------------------------------------- compaction (`_event/compact.py:406`)
def compact_1d(spikes):
    n = spikes.numel()
    ids = torch.zeros(n, dtype=torch.int32, device=spikes.device)
    cnt = torch.zeros(1, dtype=torch.int32, device=spikes.device)
    _call("compact", "compact_1d_bool", [_desc(_spk8(spikes)), _desc(ids), _desc(cnt)])
    return ids, cnt
