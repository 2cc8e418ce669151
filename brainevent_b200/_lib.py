"""ctypes binding of libbrainevent_b200.so (the C ABI in include/brainevent_b200.h).

The library is the product path: if it is missing or fails to load, importing any operator
raises -- there is NO CPU or PyTorch fallback (north star: "no CPU fallback").
"""
import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BE_B200_LIBPATH", os.path.join(_HERE, "libbrainevent_b200.so"))   # the env override serves A/B builds only
CSRC = os.path.join(_HERE, "csrc")

# dtype / event-kind codes (include/brainevent_b200.h)
F32, F64, F16, BF16 = 0, 1, 2, 3
SPK_BOOL, SPK_F32, SPK_PACKED, SPK_COMPACT = 0, 1, 2, 3
SPK_COMPACT_CLEARED = 4      # COMPACT whose producer already cleared the consumer's output (dist.P2PSpikeExchange)
JIT_SCALAR, JIT_NORMAL, JIT_UNIFORM = 0, 1, 2

_lib = None
_tls = threading.local()


def note_device(index):
    """_runtime.stream_ptr() records the device whose stream the next C call is enqueued on; the guarded
    library below switches the current device for that one call if it differs (kernels, memsets and
    function attributes apply to the CURRENT device, not to the stream's)."""
    _tls.device = index


class _GuardedLib:
    """ctypes library whose calls run with the device recorded by note_device() current."""

    def __init__(self, L):
        self._L = L

    def __getattr__(self, name):
        fn = getattr(self._L, name)

        def call(*args):
            dev = _tls.__dict__.pop("device", None)
            if dev is not None:
                import torch
                if dev != torch.cuda.current_device():
                    with torch.cuda.device(dev):
                        return fn(*args)
            return fn(*args)

        call.__name__ = name
        setattr(self, name, call)
        return call


class KernelLoadError(RuntimeError):
    """The native library could not be loaded (ref: brainevent/_error.py KernelLoadError)."""


class KernelExecutionError(RuntimeError):
    """A native entry point returned a non-zero status (ref: _error.py KernelExecutionError)."""


def build(verbose: bool = False, jobs: int = 8) -> str:
    """Compile every CUDA source for sm_100a into libbrainevent_b200.so (nvcc cross-compiles
    without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, f"-j{jobs}"], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
    if r.returncode != 0:
        raise KernelLoadError("building libbrainevent_b200.so failed:\n" + r.stdout[-4000:])
    return LIB_PATH


_P, _I, _L, _Z = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_size_t
_D = ctypes.c_double

_SIGNATURES = {
    "be_b200_last_error": (ctypes.c_char_p, []),
    "be_b200_abi_version": (_I, []),
    "be_b200_check_device": (_I, []),
    "be_b200_workspace_bytes": (_Z, [_L, _L]),
    "be_b200_profile_begin": (_I, [_I]),
    "be_b200_profile_end": (_I, [_P, _I, _P]),
    "be_b200_bitpack_1d": (_I, [_P, _I, _L, _I, _P, _P]),
    "be_b200_compact_1d": (_I, [_P, _I, _L, _I, _P, _P, _P, _P, _Z, _P]),
    "be_b200_compact_2d": (_I, [_P, _I, _L, _L, _I, _P, _P, _P, _P, _Z, _P]),
    "be_b200_binary_csrmv": (_I, [_P, _I, _I, _P, _P, _I, _P, _I, _P, _P, _I, _L, _L, _L, _I, _P, _P, _Z, _P]),
    "be_b200_binary_csrmv_auto": (_I, [_P, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P, _I, _L, _L, _L, ctypes.c_float, _P, _P, _Z, _P]),
    "be_b200_binary_csrmm": (_I, [_P, _I, _I, _P, _P, _I, _P, _I, _P, _P, _I, _L, _L, _L, _L, _I, _I, _P, _P, _Z, _P]),
    "be_b200_binary_fcnmv": (_I, [_P, _I, _I, _P, _L, _L, _P, _P, _I, _L, _L, _I, _P, _P, _Z, _P]),
    "be_b200_binary_fcnmm": (_I, [_P, _I, _I, _P, _L, _L, _P, _P, _I, _L, _L, _L, _I, _I, _P, _P, _Z, _P]),
    "be_b200_binary_jitmv": (_I, [_I, _P, _P, _I, _P, _P, _P, _P, _I, _L, _L, _L, _I, _L, _L, _L, _L, _P, _P, _Z, _P]),
    "be_b200_binary_jitmm": (_I, [_I, _P, _P, _I, _P, _P, _P, _P, _I, _L, _L, _L, _L, _I, _P, _P, _Z, _P]),
    "be_b200_csrmv": (_I, [_P, _I, _I, _P, _P, _I, _P, _L, _L, _L, _I, _P, _P]),
    "be_b200_csrmm": (_I, [_P, _I, _I, _P, _P, _I, _P, _L, _L, _L, _L, _I, _P, _P]),
    "be_b200_fcnmv": (_I, [_P, _I, _I, _P, _L, _L, _P, _L, _I, _P, _P]),
    "be_b200_fcnmm": (_I, [_P, _I, _I, _P, _L, _L, _P, _L, _L, _I, _P, _P]),
    "be_b200_jitmv": (_I, [_I, _P, _P, _I, _P, _P, _P, _L, _L, _L, _I, _P, _P]),
    "be_b200_jitmm": (_I, [_I, _P, _P, _I, _P, _P, _P, _L, _L, _L, _L, _I, _P, _P]),
    "be_b200_jitc_csr_workspace_bytes": (_Z, [_L, _L]),
    "be_b200_jitc_csr_count": (_I, [_P, _P, _L, _L, _L, _I, _I, _P, _P, _Z, _P]),
    "be_b200_jitc_csr_fill": (_I, [_I, _P, _P, _I, _P, _P, _L, _L, _L, _I, _I, _P, _L, _P, _P, _P, _Z, _P]),
    "be_b200_binary_2d_pair_stream_encode": (_I, [_P, _I, _L, _L, _P, _P, _P, _Z, _P]),
    "be_b200_binary_2d_csr_encode": (_I, [_P, _I, _L, _L, _P, _P, _P, _Z, _P]),
    "be_b200_binary_2d_csc_encode": (_I, [_P, _I, _L, _L, _P, _P, _P, _P, _Z, _P]),
    "be_b200_binary_2d_row_sparse_encode": (_I, [_P, _I, _L, _L, _L, _P, _P, _P]),
    "be_b200_plasticity_rows": (_I, [_P, _I, _P, _P, _I, _L, _P, _I, _P, _P, _I, _P, _L, _L, _L, _D, _D, _P, _Z, _P]),
    "be_b200_plasticity_dense": (_I, [_P, _I, _L, _L, _P, _P, _I, _P, _I, _D, _D, _P, _Z, _P]),
    "be_b200_clip": (_I, [_P, _I, _L, _D, _D, _P]),
    "be_b200_jitc_row_connections": (_I, [_P, _P, _L, _L, _L, _L, _I, _P, _P, _L, _P]),
    "be_b200_binary_densemv": (_I, [_P, _I, _L, _L, _P, _P, _I, _I, _P, _P, _Z, _P]),
    "be_b200_binary_densemm": (_I, [_P, _I, _L, _L, _P, _P, _I, _L, _I, _P, _P, _Z, _P]),
    "be_b200_csc_workspace_bytes": (_Z, [_L, _L, _L]),
    "be_b200_csr_to_csc": (_I, [_P, _I, _P, _L, _L, _L, _P, _P, _P, _P, _Z, _P]),
    "be_b200_exchange_ctl_bytes": (_Z, []),
    "be_b200_pack_publish": (_I, [_P, _I, _L, _P, _P, _I, _I, _L, _P, _P]),
    "be_b200_wait_exchange": (_I, [_P, _P, _I, _L, _P, _P, _P]),
    "be_b200_pack_publish_reset": (_I, [_P, _I, _L, _P, _P, _I, _I, _L, _P, _P, _P]),
    "be_b200_wait_exchange_compact": (_I, [_P, _P, _I, _L, _L, _P, _P, _P, _Z, _P, _P]),
    "be_b200_pack_publish_clear": (_I, [_P, _I, _L, _P, _P, _I, _I, _L, _P, _P, _P, _Z, _P]),
    "be_b200_sharded_symm_bytes": (_Z, [_I, _L]),
    "be_b200_sharded_create": (_P, [_I, _I, _L, _L, _L, _P, _P]),
    "be_b200_sharded_set_fcn": (_I, [_P, _P, _I, _I, _P, _L, _L]),
    "be_b200_sharded_set_csr": (_I, [_P, _P, _I, _I, _P, _P, _I, _L, _L]),
    "be_b200_sharded_set_jitc": (_I, [_P, _I, _P, _P, _I, _P, _P, _L, _L, _I]),
    "be_b200_sharded_set_dense": (_I, [_P, _P, _I]),
    "be_b200_sharded_set_option": (_I, [_P, ctypes.c_char_p, _L]),
    "be_b200_sharded_step": (_I, [_P, _P, _I, _P, _P]),
    "be_b200_sharded_publish": (_I, [_P, _P, _I, _P, _P]),
    "be_b200_sharded_consume": (_I, [_P, _P, _P]),
    "be_b200_sharded_timed_out": (_I, [_P]),
    "be_b200_sharded_host_init": (_I, [_P, _I, _I]),
    "be_b200_sharded_submit_host": (_I, [_P, _P, _P]),
    "be_b200_sharded_wait": (_I, [_P]),
    "be_b200_sharded_run_host": (_I, [_P, _L, _P, _I, _P, _I]),
    "be_b200_sharded_destroy": (None, [_P]),
    "be_b200_projection_create_csr": (_P, [_P, _I, _P, _P, _L, _L]),
    "be_b200_projection_create_fcn": (_P, [_P, _I, _P, _L, _L, _L]),
    "be_b200_projection_adopt_fcn": (_P, [_P, _I, _P, _L, _L, _L]),
    "be_b200_projection_create_jitc": (_P, [_I, ctypes.c_float, ctypes.c_float, ctypes.c_int32, ctypes.c_int32, _L, _L, _L, _I]),
    "be_b200_projection_apply_host": (_I, [_P, _P, _P]),
    "be_b200_projection_submit_host": (_I, [_P, _P, _P]),
    "be_b200_projection_wait": (_I, [_P]),
    "be_b200_projection_destroy": (None, [_P]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """Load the shared library (once).  Raises KernelLoadError if it is absent: run
    `python -c "import __graft_entry__ as g; g.build()"` or `make -C brainevent_b200/csrc`."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KernelLoadError(
                f"{LIB_PATH} not found: the CUDA extension is the only execution path of "
                "brainevent_b200 (no CPU fallback). Build it with __graft_entry__.build().")
        try:
            L = ctypes.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise KernelLoadError(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = _GuardedLib(L)
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().be_b200_last_error().decode("utf-8", "replace")
        raise KernelExecutionError(f"{what} failed (code {rc}): {msg}")
