"""CPU side of SURVEY 8(b) last row: every reference `// @BE` name of the hot path has a symbol in
libbrainevent_b200.so, every FFI target the reference's own kernel generators were recorded calling resolves to one,
and the symbols answer the XLA metadata handshake (no CUDA call on that path)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _backend():
    from integration import brainevent_b200_backend as B
    return B


def test_alias_header_symbols_are_exported():
    B = _backend()
    src = open(os.path.join(ROOT, "include", "brainevent_b200_ref_aliases.h")).read()
    names = re.findall(r"XLA_FFI_Error\* (be_\w+__\w+)\(", src)
    assert len(names) >= 260 and len(set(names)) == len(names)
    L = C.CDLL(B.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing[:5]
    assert sorted(B.TARGETS.values()) == sorted(names)


def test_generated_files_are_current():
    """tools/gen_ffi_aliases.py is the single source of the alias list"""
    import subprocess, sys
    before = {p: open(os.path.join(ROOT, p)).read() for p in ("include/brainevent_b200_ref_aliases.h",
                                                              "brainevent_b200/csrc/xla_ref_aliases.inc",
                                                              "integration/ffi_targets.json")}
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_ffi_aliases.py")], check=True, capture_output=True)
    for p, txt in before.items():
        assert open(os.path.join(ROOT, p)).read() == txt, p + " is stale: run tools/gen_ffi_aliases.py"


def test_recorded_reference_targets_resolve():
    B = _backend()
    z = np.load(os.path.join(ROOT, "tests", "golden", "ffi_calls.npz"))
    meta = json.loads(bytes(z["__meta__"]).decode())
    targets = sorted({c["target"] for c in meta["calls"]})
    assert len(targets) >= 60
    assert [t for t in targets if B.symbol_for_target(t) is None] == []
    # the hybrid modules arrive with the reference's launch-config suffix
    assert any("_hybrid_b" in t for t in targets)
    assert B.module_stem("csr_binary_csrmv_hybrid_b256_s2048_t128_n4096") == "csr_binary_csrmv_hybrid"
    assert B.module_stem("csr_binary_csrmv") == "csr_binary_csrmv"
    assert B.module_stem("csr_float_csrmv") is None and B.symbol_for_target("dense_binary_cublas.binary_densemv_cublas_t_f32_bool") is None
    # every primitive of the path has its reference generator in the table (ref: def_cuda_raw_kernel sites)
    assert len(B.PRIMITIVES) == 17


def test_metadata_handshake_on_alias_symbols():
    B = _backend()
    from tests.test_xla_ffi_gpu import Api, ApiVersion, CallFrame, ExtBase, Metadata, MetadataExt
    L = B.lib()
    api = Api()
    api.struct_size = C.sizeof(Api)
    api.api_version = ApiVersion(C.sizeof(ApiVersion), None, 0, 1)
    for target in ("csr_binary_csrmv_hybrid_b256_s2048_t128_n4096.binary_csrmv_wat_hybrid_hetero_f32_bool",
                   "fcn_binary_mv.binary_fcnmv_scatter_hetero_bool_f32", "jit_normal_binary_jitnmv.notrans_f32",
                   "jit_scalar_binary_jitsmm.pack", "dense_binary_mm.binary_densemm_transpose_f32_bool", "compact.compact_1d_bool"):
        md = Metadata()
        md.struct_size = C.sizeof(Metadata)
        ext = MetadataExt(ExtBase(C.sizeof(MetadataExt), 1, None), C.pointer(md))
        f = CallFrame()
        f.struct_size = C.sizeof(CallFrame)
        f.api = C.pointer(api)
        f.stage = 3
        f.extension_start = C.cast(C.pointer(ext), C.POINTER(ExtBase))
        fn = getattr(L, B.symbol_for_target(target))
        fn.restype = C.c_void_p
        fn.argtypes = [C.POINTER(CallFrame)]
        assert fn(C.byref(f)) is None
        assert (md.api_version.major, md.api_version.minor, md.traits) == (0, 1, 1), target
