"""Primitive + backend dispatch, mirroring ref: brainevent/_op/main.py (XLACustomKernel) and
brainevent/_registry.py -- without a tracing compiler underneath.

A primitive owns named kernel generators per (platform, backend).  A kernel generator is
`kg(**static_kwargs) -> kernel(*arrays) -> outputs` exactly as in the reference
(_op/main.py:343-416); here the only registered backend is 'b200', whose kernels call the C ABI
of libbrainevent_b200.so on torch CUDA tensors.  Resolution order is the reference's:
per-call `backend=` > config.get_backend(platform) > primitive default > first registered
(_op/main.py:504-548).
"""
from typing import Callable, Dict, List, Optional

from . import config
from ._error import KernelNotAvailableError, KernelRegistrationError

_REGISTRY: Dict[str, "XLACustomKernel"] = {}


class KernelEntry:
    def __init__(self, backend: str, platform: str, kernel_generator: Callable):
        self.backend, self.platform, self.kernel_generator = backend, platform, kernel_generator


class XLACustomKernel:
    def __init__(self, name: str, doc: Optional[str] = None):
        self.name = name
        self.__doc__ = doc
        self._kernels: Dict[str, Dict[str, KernelEntry]] = {}
        self._defaults: Dict[str, str] = {}
        self._tags = set()
        self._call = None
        if name in _REGISTRY:
            raise KernelRegistrationError(f"primitive {name!r} is already registered")
        _REGISTRY[name] = self

    # -- registration (ref: _op/main.py:343-416) ------------------------------------------
    def def_kernel(self, backend: str, platform: str, kg: Callable, asdefault: bool = False):
        self._kernels.setdefault(platform, {})[backend] = KernelEntry(backend, platform, kg)
        if asdefault or platform not in self._defaults:
            self._defaults[platform] = backend

    def def_call(self, fn: Callable):
        self._call = fn

    def def_tags(self, *tags):
        self._tags.update(tags)

    @property
    def tags(self):
        return set(self._tags)

    def available_backends(self, platform: str = "gpu") -> List[str]:
        return list(self._kernels.get(platform, {}))

    def set_default(self, platform: str, backend: str):
        if backend not in self._kernels.get(platform, {}):
            raise KernelNotAvailableError(f"{self.name}: no backend {backend!r} on {platform!r}")
        self._defaults[platform] = backend

    def get_default(self, platform: str = "gpu"):
        return self._defaults.get(platform)

    # -- dispatch (ref: _op/main.py:469-588) -----------------------------------------------
    def _resolve(self, platform: str, backend: Optional[str]) -> KernelEntry:
        table = self._kernels.get(platform, {})
        if not table:
            raise KernelNotAvailableError(f"{self.name}: no kernels registered for {platform!r}")
        for cand in (backend, config.get_backend(platform), self._defaults.get(platform)):
            if cand is None:
                continue
            if cand in table:
                return table[cand]
            if cand is backend:
                raise KernelNotAvailableError(
                    f"{self.name}: backend {backend!r} is not available on {platform!r}; "
                    f"available: {list(table)}")
        return next(iter(table.values()))

    def __call__(self, *ins, outs=None, backend: Optional[str] = None, platform: str = "gpu", **static):
        entry = self._resolve(platform, backend)
        kernel = entry.kernel_generator(outs=outs, **static)
        return kernel(*ins)

    def call(self, *args, **kwargs):
        if self._call is None:
            raise KernelRegistrationError(f"{self.name}: def_call was never set")
        return self._call(*args, **kwargs)


def get_registry() -> Dict[str, XLACustomKernel]:
    """ref: _registry.py:36"""
    return dict(_REGISTRY)


def get_all_primitive_names() -> List[str]:
    return sorted(_REGISTRY)


def get_primitives_by_tags(*tags) -> Dict[str, XLACustomKernel]:
    want = set(tags)
    return {n: p for n, p in _REGISTRY.items() if want <= p.tags}


# -- brainunit-aware helpers (duck-typed: brainunit is not a dependency) ----------------------
def split_mantissa_unit(x):
    """`u.split_mantissa_unit`: (mantissa, unit) for Quantity-like objects, (x, None) otherwise."""
    if hasattr(x, "mantissa") and hasattr(x, "unit"):
        return x.mantissa, x.unit
    return x, None


def maybe_decimal(value, *units):
    """`u.maybe_decimal(res * unit...)`: re-attach units when any operand carried one."""
    out = value
    for un in units:
        if un is not None:
            out = out * un
    return out
