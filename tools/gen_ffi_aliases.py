#!/usr/bin/env python
"""Generates the reference-name-compatible FFI surface (SURVEY 8(b) last row, Appendix D):

  brainevent_b200/csrc/xla_ref_aliases.inc   one `extern "C" XLA_FFI_Error* be_<module>__<fn>(frame)` per
                                             reference `// @BE` name on the hot path, forwarding to the generic
                                             decoder of its family in xla_ffi.cu
  include/brainevent_b200_ref_aliases.h      their declarations
  integration/ffi_targets.json               target name "<module>.<fn>" -> symbol, for the binding module

The reference composes a target as f"{module}.{fn}" and binds the symbol "be_" + fn of the module's own
shared library (ref: _op/kernix_runtime.py:161-189,308-314).  All modules share ONE library here, and
different reference modules export the same function names (`pack_bool`, `notrans_f32`, ...), so the symbol
carries the module: be_<module>__<fn>.  The FFI target names -- what jax.ffi.ffi_call is given -- are the
reference's own, unchanged.  Names only are taken from the reference; run:  python tools/gen_ffi_aliases.py
"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DT = ["_f32", "_f64", "_f16", "_bf16"]
SPK = ["_bool", "_float"]
HOMO = ["_homo", "_hetero"]

# (module, function-name list, handler call, reference file:line composing the target)
FAMILIES = []


def fam(module, names, handler, cite):
    FAMILIES.append((module, names, handler, cite))


# ---- CSR (ref: _csr/binary.py:564-621, 1235-1297; _csr/binary_indexed.py:290-392, 845-956) ----------------
fam("csr_binary_csrmv", [f"binary_csrmv_nt_auto{h}{d}{s}" for h in HOMO for d in DT for s in SPK],
    "ref_csr(f, /*mm=*/false, /*hybrid=*/false, /*perm=*/false)", "_csr/binary.py:610-615")
fam("csr_binary_csrmv_hybrid", [f"binary_csrmv_wat_hybrid{h}{d}{s}" for h in HOMO for d in DT for s in SPK],
    "ref_csr(f, false, true, false)", "_csr/binary.py:583-607")
fam("csr_binary_csrmm", [f"binary_csrmm_nt_auto{h}{d}{s}" for h in HOMO for d in DT for s in SPK],
    "ref_csr(f, true, false, false)", "_csr/binary.py:1288-1294")
fam("csr_binary_csrmm_hybrid", [f"binary_csrmm_sraw_hybrid{h}{d}{s}" for h in HOMO for d in DT for s in SPK],
    "ref_csr(f, true, true, false)", "_csr/binary.py:1254-1286")
fam("csr_binary_indexed_csrmv", [f"binary_csrmv_nt_auto_perm_hetero{d}{s}" for d in DT for s in SPK],
    "ref_csr(f, false, false, true)", "_csr/binary_indexed.py:381-389")
fam("csr_binary_indexed_csrmv_hybrid", [f"binary_indexed_csrmv_wat_hybrid_hetero{d}{s}" for d in DT for s in SPK],
    "ref_csr(f, false, true, true)", "_csr/binary_indexed.py:340-366")
fam("csr_binary_indexed_csrmm", [f"binary_csrmm_nt_auto_perm_hetero{d}{s}" for d in DT for s in SPK],
    "ref_csr(f, true, false, true)", "_csr/binary_indexed.py:845-956")
fam("csr_binary_indexed_csrmm_hybrid", [f"binary_indexed_csrmm_sraw_hybrid_hetero{d}{s}" for d in DT for s in SPK],
    "ref_csr(f, true, true, true)", "_csr/binary_indexed.py:845-956")
# ---- FCN (ref: _fcn/binary.py:256-288, 770-864) -------------------------------------------------------------
fam("fcn_binary_mv", [f"binary_fcnmv_scatter{h}_bool{d}" for h in HOMO for d in DT],
    "ref_fcnmv_scatter(f)", "_fcn/binary.py:281-282")
fam("fcn_binary_mm", [f"binary_fcnmm_sraw{h}{s}{d}" for h in HOMO for s in SPK for d in DT],
    "ref_fcnmm(f, /*scatter=*/true, /*packed=*/false)", "_fcn/binary.py:849-852")
fam("fcn_binary_mm", [f"binary_fcnmm_gather{h}{z}{s}{d}" for h in HOMO for z in ("_warp", "_basic") for s in SPK for d in DT],
    "ref_fcnmm(f, false, false)", "_fcn/binary.py:821-825")
fam("fcn_binary_mm", [f"binary_fcnmm_gather_packed{h}{z}{d}" for h in HOMO for z in ("_warp", "_basic") for d in DT],
    "ref_fcnmm(f, false, true)", "_fcn/binary.py:814")
fam("fcn_binary_mm", ["binary_fcnmm_pack_bool"] + [f"binary_fcnmm_pack{d}" for d in DT],
    "ref_fcnmm_pack(f)", "_fcn/binary.py:808-810")
# ---- JITC (ref: _jit_scalar/binary.py:424-460, 960-1001 and the normal / uniform twins) ---------------------
for kind, tag, letter in (("BE_B200_JIT_SCALAR", "scalar", "s"), ("BE_B200_JIT_NORMAL", "normal", "n"),
                          ("BE_B200_JIT_UNIFORM", "uniform", "u")):
    mv, mm = f"jit_{tag}_binary_jit{letter}mv", f"jit_{tag}_binary_jit{letter}mm"
    cite = f"_jit_{tag}/binary.py"
    fam(mv, ["pack_bool"], "ref_jit_pack_bool(f)", cite)
    fam(mv, [f"notrans{d}" for d in DT], f"ref_jitmv(f, {kind}, /*corder=*/1)", cite)
    fam(mv, [f"trans{d}" for d in DT], f"ref_jitmv(f, {kind}, 0)", cite)
    fam(mm, ["pack"], "ref_jit_pack_mm(f)", cite)
    fam(mm, [f"notrans{d}" for d in DT], f"ref_jitmm(f, {kind}, 1)", cite)
    fam(mm, [f"trans{d}" for d in DT], f"ref_jitmm(f, {kind}, 0)", cite)
# ---- dense (ref: _dense/binary.py:228-250, 933-990) ---------------------------------------------------------
fam("dense_binary_mv", [f"binary_densemv_{t}{d}{s}" for t in ("transpose", "no_transpose") for d in DT for s in SPK],
    None, "_dense/binary.py:244-247")
fam("dense_binary_mm", [f"binary_densemm_{t}{d}{s}" for t in ("transpose", "no_transpose") for d in DT for s in SPK],
    None, "_dense/binary.py:975-988")
# ---- compaction (ref: _event/compact.py:135-156, 282-303, 428-449) ------------------------------------------
fam("compact", [f"compact_1d{s}" for s in SPK], "be_b200_xla_compact_1d(f)", "_event/compact.py:428-449")
fam("compact", [f"compact_2d_only{s}" for s in SPK], "be_b200_xla_compact_2d_only(f)", "_event/compact.py:282-303")
fam("compact", [f"fused_bitpack_compact_2d{s}" for s in SPK], "be_b200_xla_compact_2d(f)", "_event/compact.py:135-156")


def handler_for(module, name, handler):
    if handler is not None:
        return handler
    t = 0 if "no_transpose" in name else 1
    return f"ref_dense(f, /*mm=*/{'true' if module.endswith('_mm') else 'false'}, /*transpose=*/{t})"


def main():
    inc, hdr, table = [], [], {}
    inc.append("// GENERATED by tools/gen_ffi_aliases.py -- do not edit.  One symbol per reference `// @BE` name on the hot path.\n")
    hdr.append("/* GENERATED by tools/gen_ffi_aliases.py -- do not edit.\n"
               " * Reference-name-compatible XLA-FFI symbols of libbrainevent_b200.so: be_<module>__<fn> serves the reference's\n"
               " * FFI target \"<module>.<fn>\" (hybrid modules: any config suffix after the module stem) with the reference's own\n"
               " * argument order, result order, aliasing and attributes.  Bind with integration/brainevent_b200_backend.py. */\n"
               "#ifndef BRAINEVENT_B200_REF_ALIASES_H_\n#define BRAINEVENT_B200_REF_ALIASES_H_\n"
               "#include \"brainevent_b200_xla.h\"\n#ifdef __cplusplus\nextern \"C\" {\n#endif\n")
    for module, names, handler, cite in FAMILIES:
        hdr.append(f"/* ref: {cite} -- module {module} */\n")
        for n in names:
            sym = f"be_{module}__{n}"
            assert sym not in table.values(), sym
            table[f"{module}.{n}"] = sym
            inc.append(f"XLA_FFI_Error* {sym}(XLA_FFI_CallFrame* f) {{ return {handler_for(module, n, handler)}; }}\n")
            hdr.append(f"XLA_FFI_Error* {sym}(XLA_FFI_CallFrame* frame);\n")
    hdr.append("#ifdef __cplusplus\n}\n#endif\n#endif\n")
    open(os.path.join(ROOT, "brainevent_b200", "csrc", "xla_ref_aliases.inc"), "w").write("".join(inc))
    open(os.path.join(ROOT, "include", "brainevent_b200_ref_aliases.h"), "w").write("".join(hdr))
    os.makedirs(os.path.join(ROOT, "integration"), exist_ok=True)
    with open(os.path.join(ROOT, "integration", "ffi_targets.json"), "w") as f:
        json.dump({"_comment": "reference FFI target -> symbol in libbrainevent_b200.so (generated by tools/gen_ffi_aliases.py); "
                               "hybrid modules match by stem: the reference appends its launch-config suffix to the module name",
                   "targets": table}, f, indent=0, sort_keys=True)
    print(f"{len(table)} alias symbols")


if __name__ == "__main__":
    main()
