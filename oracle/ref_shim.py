"""Tier-1 shim: import the UNMODIFIED reference (/root/reference) without jax/brainunit.

TEST INFRASTRUCTURE ONLY. Works only in the build container (where /root/reference exists);
it is used by tests/golden/make_golden.py to generate golden fixtures and by
tests/test_oracle_vs_reference.py (skipped when the reference is absent) to pin the C/numpy
oracle in oracle/ against the reference's own numba kernel bodies.

Recipe: jax / brainunit are replaced by sys.modules stubs whose
`jax.numpy` falls through to real numpy, then `numba_kernel` in each reference module is
monkey-patched to a direct call `fn(*ins, *zeros(outs))` (the semantics of
brainevent/_op/numba_ffi.py:1041-1044: inputs first, outputs after, written in place).
Nothing from the reference is copied; it is imported from where it lies.
"""
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types
from unittest import mock

import numpy as np

REFERENCE_ROOT = os.environ.get("BRAINEVENT_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "brainevent"))


class SDS:
    """Stand-in for jax.ShapeDtypeStruct (shape/dtype/size only)."""

    def __init__(self, shape, dtype, **kw):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.size = int(np.prod(self.shape)) if len(self.shape) else 1
        self.ndim = len(self.shape)


class _MockLoader(importlib.abc.Loader):
    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__name__ = spec.name
        m.__path__ = []
        m.__spec__ = spec
        m.__loader__ = self
        return m

    def exec_module(self, module):
        pass


class _MockFinder(importlib.abc.MetaPathFinder):
    PREFIXES = ("jax", "jaxlib", "brainunit", "brainstate", "braintools")

    def find_spec(self, fullname, path, target=None):
        root = fullname.split(".")[0]
        if root in self.PREFIXES and fullname not in sys.modules:
            return importlib.machinery.ModuleSpec(fullname, _MockLoader(), is_package=True)
        return None


_installed = False


def _install_stubs():
    global _installed
    if _installed:
        return
    try:
        import ml_dtypes  # noqa: F401  (registers bfloat16 with numpy)
    except Exception:
        pass

    jax = types.ModuleType("jax")
    jax.__path__ = []
    jax.__version__ = "0.8.0"
    jax.__version_info__ = (0, 8, 0)
    jax.ShapeDtypeStruct = SDS

    class _Array:  # isinstance targets
        pass

    jax.Array = _Array
    jax.typing = types.ModuleType("jax.typing")
    jax.typing.ArrayLike = object
    jax.typing.DTypeLike = object

    tree_util = types.ModuleType("jax.tree_util")
    tree_util.register_pytree_node_class = lambda c: c
    tree_util.register_pytree_node = lambda *a, **k: None
    jax.tree_util = tree_util

    class _JNP(types.ModuleType):
        def __getattr__(self, name):
            return getattr(np, name)

    jnp = _JNP("jax.numpy")
    jnp.__path__ = []
    jax.numpy = jnp

    class Primitive:
        def __init__(self, name):
            self.name = name
            self.multiple_results = False

        def def_impl(self, *a, **k):
            return None

        def def_abstract_eval(self, *a, **k):
            return None

        def bind(self, *a, **k):
            raise RuntimeError("stub primitive cannot bind")

    class Tracer:
        pass

    core = types.ModuleType("jax.core")
    core.Primitive = Primitive
    core.Tracer = Tracer
    core.ShapedArray = SDS
    jax.core = core
    extend = types.ModuleType("jax.extend")
    extend.__path__ = []
    extend_core = types.ModuleType("jax.extend.core")
    extend_core.Primitive = Primitive
    extend.core = extend_core
    jax.extend = extend

    sys.modules["jax"] = jax
    sys.modules["jax.numpy"] = jnp
    sys.modules["jax.tree_util"] = tree_util
    sys.modules["jax.core"] = core
    sys.modules["jax.extend"] = extend
    sys.modules["jax.extend.core"] = extend_core
    sys.modules["jax.typing"] = jax.typing

    bu = types.ModuleType("brainunit")
    bu.__path__ = []

    class Quantity:
        pass

    bu.Quantity = Quantity
    bu_sparse = types.ModuleType("brainunit.sparse")

    class SparseMatrix:
        pass

    bu_sparse.SparseMatrix = SparseMatrix
    bu_sparse.__getattr__ = lambda name: mock.MagicMock(name="brainunit.sparse." + name)
    bu.sparse = bu_sparse
    _mm = mock.MagicMock(name="brainunit.math")
    bu.math = _mm
    bu.__getattr__ = lambda name: mock.MagicMock(name="brainunit." + name)
    sys.modules["brainunit"] = bu
    sys.modules["brainunit.sparse"] = bu_sparse

    sys.meta_path.insert(0, _MockFinder())
    # let attribute access on the stub jax module fall back to mocks
    jax.__getattr__ = lambda name: importlib.import_module("jax." + name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def _shim_numba_kernel(fn, outs, **kw):
    single = not isinstance(outs, (list, tuple))
    outs_l = [outs] if single else list(outs)

    def call(*ins):
        o = [np.zeros(s.shape, s.dtype) for s in outs_l]
        fn(*[np.ascontiguousarray(i) if isinstance(i, np.ndarray) else i for i in ins], *o)
        return o

    return call


_mod_cache = {}


def ref_module(name: str):
    """Import `brainevent.<name>` from the reference with numba_kernel patched to a direct call."""
    if not available():
        raise RuntimeError("reference not present at " + REFERENCE_ROOT)
    _install_stubs()
    if name in _mod_cache:
        return _mod_cache[name]
    m = importlib.import_module("brainevent." + name)
    if hasattr(m, "numba_kernel"):
        m.numba_kernel = _shim_numba_kernel
    _mod_cache[name] = m
    return m
