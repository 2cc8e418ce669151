"""Binding of libbrainevent_b200.so into chaobrain/brainevent (the reference), for a machine that has jax.

What a maintainer adds to the reference: this ONE module next to the package, and a call to `install()` before the
first operator runs.  Nothing else changes: every hot-path primitive keeps its own `cuda_raw` kernel generator
(`_binary_csrmv_cuda_kernel`, `_ell_binary_matvec_cuda_kernel`, `_binary_jitsmv_cuda_kernel`, ... -- the full table is
PRIMITIVES below), with its own argument preparation, `jax.ffi.ffi_call` target names, `input_output_aliases`
and scalar attributes.  Those generators reach native code through exactly one function,

    brainevent._op.load_cuda_file(path_to_cu, name=<module>, ...)          ref: _op/kernix_pipeline.py:448-473

which nvcc-compiles the `.cu` file and registers one FFI target "<module>.<fn>" per `// @BE` function
(ref: _op/kernix_runtime.py:161-189, 308-314).  `install()` replaces that function with `load_cuda_file` below, which
compiles nothing: it registers the same target names against the symbols `be_<module>__<fn>` of
libbrainevent_b200.so (include/brainevent_b200_ref_aliases.h; hand-written sm_100a kernels behind them).  The
hybrid CSR modules carry the launch-config suffix of the reference (`csr_binary_csrmv_hybrid_b256_s2048_t128_n4096`,
ref: _csr/binary.py:583-589); the suffix is accepted and ignored -- this library has no launch configuration.

`register_b200_backend()` additionally registers the same generators under the backend name 'b200' as the GPU default,
so `brainevent.config.set_backend('gpu', 'b200')` and per-call `backend='b200'` work.  Only the hot path is covered
(SURVEY 8a); modules this library does not serve (`float_*.cu`, plasticity, cuBLAS, ...) fall through to the reference's
own loader.

This file imports jax lazily: with no jax in the process (this repository's test image) the table and the symbol
resolution are still testable, and tests/test_ffi_replay_gpu.py drives every family through a hand-built call frame.
"""
import ctypes
import json
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("BRAINEVENT_B200_LIB", os.path.join(_ROOT, "brainevent_b200", "libbrainevent_b200.so"))

with open(os.path.join(_HERE, "ffi_targets.json")) as _f:
    TARGETS = json.load(_f)["targets"]          # "<module>.<fn>" -> symbol
MODULES = sorted({t.split(".", 1)[0] for t in TARGETS})

# (reference module, primitive, its cuda_raw kernel generator, reference file:line of the registration)
PRIMITIVES = [
    ("brainevent._csr.binary", "binary_csrmv_p", "_binary_csrmv_cuda_kernel", "_csr/binary.py:1016"),
    ("brainevent._csr.binary", "binary_csrmm_p", "_binary_csrmm_cuda_kernel", "_csr/binary.py:1636"),
    ("brainevent._csr.binary_indexed", "binary_csrmv_indexed_p", "_binary_csrmv_indexed_cuda_kernel", "_csr/binary_indexed.py:601"),
    ("brainevent._csr.binary_indexed", "binary_csrmm_indexed_p", "_binary_csrmm_indexed_cuda_kernel", "_csr/binary_indexed.py:1153"),
    ("brainevent._fcn.binary", "binary_fcnmv_p", "_ell_binary_matvec_cuda_kernel", "_fcn/binary.py:546"),
    ("brainevent._fcn.binary", "binary_fcnmm_p", "_binary_fcnmm_cuda_kernel", "_fcn/binary.py:1167"),
    ("brainevent._jit_scalar.binary", "binary_jitsmv_p", "_binary_jitsmv_cuda_kernel", "_jit_scalar/binary.py:824"),
    ("brainevent._jit_scalar.binary", "binary_jitsmm_p", "_binary_jitsmm_cuda_kernel", "_jit_scalar/binary.py:1372"),
    ("brainevent._jit_normal.binary", "binary_jitnmv_p", "_binary_jitnmv_cuda_kernel", "_jit_normal/binary.py:694"),
    ("brainevent._jit_normal.binary", "binary_jitnmm_p", "_binary_jitnmm_cuda_kernel", "_jit_normal/binary.py:1116"),
    ("brainevent._jit_uniform.binary", "binary_jitumv_p", "_binary_jitumv_cuda_kernel", "_jit_uniform/binary.py:875"),
    ("brainevent._jit_uniform.binary", "binary_jitumm_p", "_binary_jitumm_cuda_kernel", "_jit_uniform/binary.py:1504"),
    ("brainevent._dense.binary", "binary_densemv_p", "_binary_densemv_cuda_kernel", "_dense/binary.py:463"),
    ("brainevent._dense.binary", "binary_densemm_p", "_binary_densemm_cuda_kernel", "_dense/binary.py:1044"),
    ("brainevent._event.compact", "binary_1d_array_index_p", "_binary_1d_array_index_cuda_kernel", "_event/compact.py:415"),
    ("brainevent._event.compact", "binary_2d_array_index_p", "_binary_2d_array_index_cuda_kernel", "_event/compact.py:598"),
    ("brainevent._event.compact", "binary_2d_compact_only_p", "_binary_2d_compact_only_cuda_kernel", "_event/compact.py:268"),
]

_lib = None
_registered = set()
_reference_loader = None


def lib():
    """The shared library (loaded once).  There is no fallback: a missing library is an error."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError(f"{LIB_PATH} not found: build it with `make -C brainevent_b200/csrc` (sm_100a, nvcc 12.9)")
        _lib = ctypes.CDLL(LIB_PATH)
    return _lib


def module_stem(module: str):
    """'csr_binary_csrmv_hybrid_b256_s2048_t128_n4096' -> 'csr_binary_csrmv_hybrid'; None if this library does not
    serve the module.  Longest served stem wins ('csr_binary_csrmv_hybrid' before 'csr_binary_csrmv')."""
    best = None
    for stem in MODULES:
        if module == stem or (module.startswith(stem + "_") and re.fullmatch(r"(_[a-z]\d+)+", module[len(stem):])):
            if best is None or len(stem) > len(best):
                best = stem
    return best


def symbol_for_target(target: str):
    """Symbol of libbrainevent_b200.so behind the reference FFI target '<module>.<fn>' (None if not served)."""
    module, _, fn = target.partition(".")
    stem = module_stem(module)
    return TARGETS.get(f"{stem}.{fn}") if stem else None


def handler_address(target: str) -> int:
    sym = symbol_for_target(target)
    if sym is None:
        raise KeyError(f"no B200 kernel behind FFI target {target!r}")
    fn = getattr(lib(), sym)
    fn.restype = ctypes.c_void_p
    fn.argtypes = [ctypes.c_void_p]
    return ctypes.cast(fn, ctypes.c_void_p).value


def register_module(module: str) -> int:
    """Register every FFI target of one reference module name (as the reference spells it, config suffix included)
    with jax: the twin of CompiledModule registration, ref: _op/kernix_runtime.py:308-314."""
    import jax
    stem = module_stem(module)
    if stem is None:
        raise KeyError(module)
    n = 0
    for target, sym in TARGETS.items():
        mod, _, fn = target.partition(".")
        if mod != stem:
            continue
        name = f"{module}.{fn}"
        if name in _registered:
            continue
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(lib(), sym)), platform="CUDA")
        _registered.add(name)
        n += 1
    return n


class _Module:
    """What load_cuda_file returns (the reference returns a CompiledModule; its callers ignore the value)."""

    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return f"<brainevent_b200 module {self.name}>"


def load_cuda_file(filepath, functions=None, *, name=None, **kwargs):
    """Drop-in for brainevent._op.load_cuda_file (ref: _op/kernix_pipeline.py:448-473).  Same signature; `filepath`,
    `functions`, `extra_cuda_cflags`, `allow_cuda_graph` are accepted and unused -- nothing is compiled."""
    module = name if name is not None else os.path.splitext(os.path.basename(str(filepath)))[0]
    if module_stem(module) is None:
        if _reference_loader is None:
            raise KeyError(f"module {module!r} is not served by libbrainevent_b200.so and no reference loader is installed")
        return _reference_loader(filepath, functions, name=name, **kwargs)
    register_module(module)
    return _Module(module)


def install():
    """Point the reference's hot-path modules at this library.  Idempotent."""
    global _reference_loader
    import importlib
    import brainevent._op as op
    if getattr(op.load_cuda_file, "__module__", "") != __name__:
        _reference_loader = op.load_cuda_file
        op.load_cuda_file = load_cuda_file
    for modname in sorted({m for m, _, _, _ in PRIMITIVES}):
        m = importlib.import_module(modname)
        if hasattr(m, "load_cuda_file"):
            m.load_cuda_file = load_cuda_file          # `from brainevent._op import load_cuda_file` bound the old one


def register_b200_backend(asdefault: bool = True):
    """install() + every hot-path primitive gets backend 'b200' on 'gpu': the reference's own generator, now bound to
    this library (ref: XLACustomKernel.def_kernel, _op/main.py:343-416)."""
    import importlib
    install()
    done = []
    for modname, prim, gen, _ in PRIMITIVES:
        m = importlib.import_module(modname)
        getattr(m, prim).def_kernel("b200", "gpu", getattr(m, gen), asdefault=asdefault)
        done.append(prim)
    return done
