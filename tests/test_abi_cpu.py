"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/*.h declares, and the host mirror fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for h in ("brainevent_b200.h", "brainevent_b200_xla.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(be_b200_\w+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    from brainevent_b200 import _lib
    L = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared()
    assert len(declared) >= 35
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    # and the ctypes table binds exactly the plain-C part
    plain = {n for n in declared if not n.startswith("be_b200_xla_")}
    assert plain == set(_lib.EXPORTED_SYMBOLS), plain ^ set(_lib.EXPORTED_SYMBOLS)


def test_abi_version_and_workspace_query():
    from brainevent_b200 import _lib
    L = _lib.lib()
    assert L.be_b200_abi_version() == 1
    a = L.be_b200_workspace_bytes(1_000_000, 1)
    b = L.be_b200_workspace_bytes(1_000_000, 64)
    assert a >= 4 * 1_000_000 and b >= a and a % 256 == 0


def test_no_cpu_fallback():
    import brainevent_b200 as be
    w = torch.ones(3)
    idx = torch.tensor([0, 2, 1], dtype=torch.int32)
    ptr = torch.tensor([0, 2, 3], dtype=torch.int32)
    with pytest.raises(be.KernelExecutionError):
        be.binary_csrmv(w, idx, ptr, torch.tensor([True, False, True]), shape=(2, 3))
    with pytest.raises(be.KernelExecutionError):
        be.BinaryArray(torch.tensor([True, False])) @ be.CSR((w, idx, ptr), shape=(2, 3))


def test_primitive_registry_mirrors_reference_names():
    import brainevent_b200 as be
    names = set(be.get_all_primitive_names())
    for n in ("binary_csrmv", "binary_csrmm", "binary_csrmv_indexed", "binary_csrmm_indexed", "binary_fcnmv",
              "binary_fcnmm", "binary_jitsmv", "binary_jitsmm", "binary_jitnmv", "binary_jitnmm", "binary_jitumv",
              "binary_jitumm", "binary_densemv", "binary_densemm", "binary_1d_array_index", "binary_2d_array_index",
              "binary_2d_compact_only"):
        assert n in names
        p = be.get_registry()[n]
        assert p.available_backends("gpu") == ["b200"] and p.get_default("gpu") == "b200"
    assert set(be.get_primitives_by_tags("csr", "binary")) >= {"binary_csrmv", "binary_csrmm"}
    with pytest.raises(be.KernelNotAvailableError):
        be.binary_csrmv_p._resolve("gpu", "numba")
    with pytest.raises(ValueError):
        be.config.set_backend("gpu", "pallas")


def test_shape_assertions_match_reference_behaviour():
    import brainevent_b200 as be
    from brainevent_b200 import _csr
    w = torch.ones(3)
    idx = torch.tensor([0, 2, 1], dtype=torch.int32)
    ptr = torch.tensor([0, 2, 3], dtype=torch.int32)
    # wrong vector length -> assertion before any launch (ref: _csr/binary.py:827-987 asserts)
    with pytest.raises((AssertionError, be.KernelExecutionError)):
        be.binary_csrmv(w, idx, ptr, torch.tensor([True, False]), shape=(2, 3))
    assert _csr._prep_structure  # host logic is importable without CUDA


def test_reference_cuda_shim_exports():
    """oracle/_ref/libref_*.so (the reference's own kernels behind a plain C ABI, built in the
    container that has /root/reference) load and export the entry points oracle/ref_cuda.py binds."""
    import ctypes
    from oracle import ref_cuda as R
    want = {"fcnmv": ["binary_fcnmv_scatter_homo_bool_f32", "binary_fcnmv_scatter_hetero_bool_f32"],
            "csrmv": ["binary_csrmv_nt_auto_homo_f32_bool", "binary_csrmv_nt_auto_hetero_f32_bool"],
            "csrmv_hybrid": ["binary_csrmv_wat_hybrid_homo_f32_bool", "binary_csrmv_wat_hybrid_hetero_f32_bool"],
            "csrmm_hybrid": ["binary_csrmm_sraw_hybrid_homo_f32_bool", "binary_csrmm_sraw_hybrid_hetero_f32_bool"],
            "jitsmv": ["pack_bool", "notrans_f32", "trans_f32"], "jitnmv": ["pack_bool", "notrans_f32", "trans_f32"],
            "jitumv": ["pack_bool", "notrans_f32", "trans_f32"], "densemm": ["binary_densemm_transpose_f32_bool"],
            "compact": ["compact_1d_bool"]}
    built = [n for n in want if R.available(n)]
    if not built:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    for name in built:
        lib = ctypes.CDLL(os.path.join(os.path.dirname(R.__file__), "_ref", f"libref_{name}.so"))
        for fn in want[name]:
            assert hasattr(lib, "ref_" + fn), (name, fn)


def test_exchange_entry_points_validate_before_launching():
    """The peer-memory exchange entries reject bad arguments with BE_B200_EINVAL and a message before
    they touch the device (safe to call without a GPU), and the ctypes table matches the header."""
    from brainevent_b200 import _lib
    L = _lib.lib()
    assert L.be_b200_exchange_ctl_bytes() >= 16
    null = ctypes.c_void_p(0)
    one = ctypes.c_void_p(16)            # never dereferenced: validation fails first
    rc = L.be_b200_wait_exchange_compact(null, null, 2, 4, 128, null, null, null, 0, null, null)
    assert rc != 0 and b"wait_exchange_compact" in L.be_b200_last_error()
    rc = L.be_b200_wait_exchange_compact(one, one, 2, 4, 4 * 2 * 32 + 1, one, one, null, 0, one, null)
    assert rc != 0 and b"n_bits" in L.be_b200_last_error()            # ids would overrun the gathered vector
    rc = L.be_b200_pack_publish_reset(one, _lib.SPK_BOOL, 129, one, one, 2, 0, 4, one, null, null)
    assert rc != 0 and b"n_local" in L.be_b200_last_error()           # 129 spikes do not fit 4 words
    rc = L.be_b200_pack_publish_reset(one, _lib.SPK_BOOL, 64, one, one, 2, 5, 4, one, null, null)
    assert rc != 0 and b"rank" in L.be_b200_last_error()
    assert _lib.SPK_COMPACT_CLEARED == 4
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "brainevent_b200.h")).read()
    assert "#define BE_B200_SPK_COMPACT_CLEARED 4" in hdr
