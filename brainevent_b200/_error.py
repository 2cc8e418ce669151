"""Exception taxonomy (names follow ref: brainevent/_error.py:43-405)."""
from ._lib import KernelExecutionError, KernelLoadError  # noqa: F401


class BrainEventError(Exception):
    pass


class MathError(BrainEventError):
    pass


class UnsupportedOperationError(BrainEventError, NotImplementedError):
    pass


class KernelError(BrainEventError):
    pass


class KernelNotAvailableError(KernelError):
    pass


class KernelRegistrationError(KernelError):
    pass
