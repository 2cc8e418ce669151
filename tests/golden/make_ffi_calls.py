#!/usr/bin/env python
"""Records what the REFERENCE's own cuda_raw kernel generators hand to `jax.ffi.ffi_call`, per hot-path
primitive, into tests/golden/ffi_calls.npz -- run here (build container: /root/reference present, no jax, no GPU):

    python tests/golden/make_ffi_calls.py

The unmodified generators (`_binary_csrmv_cuda_kernel`, `_ell_binary_matvec_cuda_kernel`,
`_binary_jitsmv_cuda_kernel`, ... ref: integration/brainevent_b200_backend.py PRIMITIVES) are imported through the
jax/brainunit stub shim (oracle/ref_shim.py) with two substitutions: `load_cuda_file` records the module name instead
of compiling, and `jax.ffi.ffi_call` records (target, operand arrays, result specs, input_output_aliases, scalar
attributes) and hands back zero arrays of the result specs so the generator's Python continues (pack -> kernel chains
are kept as data-flow links).  Each case also stores the expected result of the primitive from the oracle (the pinned
CPU restatement of the reference's numba kernels).  tests/test_ffi_replay_gpu.py replays every recorded call on the
B200 through libbrainevent_b200.so's reference-named symbols with a hand-built XLA_FFI_CallFrame and compares.
"""
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O                      # noqa: E402
from oracle import ref_shim as RS       # noqa: E402
from tests import cases                 # noqa: E402

SDS = RS.SDS
RS._install_stubs()
import jax                              # noqa: E402  (the stub)

CALLS, MODULES = [], []
_live = {}            # id(array) -> (call index, result index); arrays are kept alive in CALLS


def _fake_ffi_call(name, out_specs, input_output_aliases=None, **kw):
    single = not isinstance(out_specs, (list, tuple))
    specs = [out_specs] if single else list(out_specs)

    def call(*arrays, **attrs):
        ins = []
        for a in arrays:
            if id(a) in _live:
                ins.append(("link",) + _live[id(a)])
            else:
                ins.append(("array", np.ascontiguousarray(a)))
        outs = [np.zeros(s.shape, s.dtype) for s in specs]
        idx = len(CALLS)
        CALLS.append(dict(target=name, ins=ins, outs=[(tuple(s.shape), np.dtype(s.dtype).name) for s in specs],
                          aliases={int(k): int(v) for k, v in (input_output_aliases or {}).items()},
                          attrs={k: (int(v), np.asarray(v).dtype.name) for k, v in attrs.items()}, keep=outs))
        for j, o in enumerate(outs):
            _live[id(o)] = (idx, j)
        return outs[0] if single else outs
    return call


def _fake_load(path, functions=None, *, name=None, **kw):
    MODULES.append((name, os.path.basename(str(path))))


jax.ffi = types.SimpleNamespace(ffi_call=_fake_ffi_call)
sys.modules["brainunit"].math.asarray = np.asarray


def mod(name):
    m = RS.ref_module(name)
    m.load_cuda_file = _fake_load
    return m


CASES = []


def record(name, fn, expected, out_index=0, transpose_result=False, rtol=1e-5, atol=1e-5, exact=False, sort_ids=False):
    """fn() runs one reference kernel closure; the calls it makes since the last record() belong to this case."""
    first = len(CALLS)
    result = fn()
    res = result[out_index] if isinstance(result, (list, tuple)) else result
    # which recorded output carries the primitive's result (possibly through the generator's jnp.transpose)
    src = None
    base = res.base if (isinstance(res, np.ndarray) and res.base is not None) else res
    for cand in (res, base):
        if id(cand) in _live:
            src = _live[id(cand)]
    assert src is not None, name
    transposed = bool(src is not None and tuple(res.shape) != tuple(CALLS[src[0]]["keep"][src[1]].shape))
    CASES.append(dict(name=name, first=first, last=len(CALLS), result=src, transposed=transposed,
                      expected=np.asarray(expected), rtol=rtol, atol=atol, exact=exact, sort_ids=sort_ids))


def main():
    rng = np.random.default_rng(2024)
    f32 = np.float32
    # ---------------------------------------------------------------- CSR mv / mm (+ indexed)
    csr = mod("_csr.binary")
    csri = mod("_csr.binary_indexed")
    m_, k_ = 700, 500
    for homo in (False, True):
        c = cases.csr_case(1, m_, k_, 12, 0.0, homo=homo, ragged=True)
        w, idx, ptr = np.atleast_1d(c["data"]).astype(f32), c["indices"], c["indptr"]
        cap = 3
        tb, te, st = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(2, np.int32)
        for spk_float in (False, True):
            for transpose in (True, False):
                v = cases.spikes(3, m_ if transpose else k_, 0.2, float_spikes=spk_float)
                n_out = k_ if transpose else m_
                kern = csr._binary_csrmv_cuda_kernel(
                    weight_info=SDS(w.shape, w.dtype), vector_info=SDS(v.shape, v.dtype), transpose=transpose,
                    indices_info=SDS(idx.shape, idx.dtype), indptr_info=SDS(ptr.shape, ptr.dtype),
                    outs=[SDS((n_out,), f32), SDS(tb.shape, tb.dtype), SDS(te.shape, te.dtype), SDS(st.shape, st.dtype)],
                    task_capacity=cap)
                record(f"binary_csrmv homo={homo} float={spk_float} T={transpose}",
                       lambda: kern(w, idx, ptr, v, tb, te, st),
                       O.binary_csrmv(w, idx, ptr, v, shape=(m_, k_), transpose=transpose))
                B = cases.spikes(4, m_ if transpose else k_, 0.15, float_spikes=spk_float, batch=40)
                kern = csr._binary_csrmm_cuda_kernel(
                    weight_info=SDS(w.shape, w.dtype), vector_info=SDS(B.shape, B.dtype), transpose=transpose,
                    indices_info=SDS(idx.shape, idx.dtype), indptr_info=SDS(ptr.shape, ptr.dtype),
                    outs=[SDS((n_out, 40), f32), SDS(tb.shape, tb.dtype), SDS(te.shape, te.dtype), SDS(st.shape, st.dtype)],
                    task_capacity=cap)
                record(f"binary_csrmm homo={homo} float={spk_float} T={transpose}",
                       lambda: kern(w, idx, ptr, B, tb, te, st),
                       O.binary_csrmm(w, idx, ptr, B, shape=(m_, k_), transpose=transpose), atol=1e-4)
        # indexed (CSC-mirror traversal reading weights[perm[j]])
        perm = rng.permutation(idx.size).astype(np.int32)
        for transpose in (True, False):
            v = cases.spikes(5, m_ if transpose else k_, 0.2)
            n_out = k_ if transpose else m_
            kern = csri._binary_csrmv_indexed_cuda_kernel(
                weight_info=SDS(w.shape, w.dtype), vector_info=SDS(v.shape, v.dtype), transpose=transpose,
                indices_info=SDS(idx.shape, idx.dtype), indptr_info=SDS(ptr.shape, ptr.dtype),
                outs=[SDS((n_out,), f32), SDS(tb.shape, tb.dtype), SDS(te.shape, te.dtype), SDS(st.shape, st.dtype)],
                task_capacity=cap)
            record(f"binary_csrmv_indexed homo={homo} T={transpose}",
                   lambda: kern(w, idx, ptr, perm, v, tb, te, st),
                   O.binary_csrmv_indexed(w, idx, ptr, perm, v, shape=(m_, k_), transpose=transpose))
            B = cases.spikes(6, m_ if transpose else k_, 0.15, batch=33)
            kern = csri._binary_csrmm_indexed_cuda_kernel(
                weight_info=SDS(w.shape, w.dtype), vector_info=SDS(B.shape, B.dtype), transpose=transpose,
                indices_info=SDS(idx.shape, idx.dtype), indptr_info=SDS(ptr.shape, ptr.dtype),
                outs=[SDS((n_out, 33), f32), SDS(tb.shape, tb.dtype), SDS(te.shape, te.dtype), SDS(st.shape, st.dtype)],
                task_capacity=cap)
            record(f"binary_csrmm_indexed homo={homo} T={transpose}",
                   lambda: kern(w, idx, ptr, perm, B, tb, te, st),
                   O.binary_csrmm_indexed(w, idx, ptr, perm, B, shape=(m_, k_), transpose=transpose), atol=1e-4)
    # ---------------------------------------------------------------- FCN mv / mm
    fcn = mod("_fcn.binary")
    n_pre, n_post, n_conn = 600, 450, 20
    for homo in (False, True):
        f = cases.fcn_case(7, n_pre, n_post, n_conn, homo=homo)
        w, idx = np.atleast_1d(f["weights"]).astype(f32), f["indices"]
        v = cases.spikes(8, n_pre, 0.2)
        kern = fcn._ell_binary_matvec_cuda_kernel(transpose=True, spike_info=SDS(v.shape, v.dtype), indices_info=SDS(idx.shape, idx.dtype),
                                                  weight_info=SDS(w.shape, w.dtype), outs=[SDS((n_post,), f32)])
        record(f"binary_fcnmv scatter homo={homo}", lambda: kern(w, idx, v),
               O.binary_fcnmv(w, idx, v, shape=(n_pre, n_post), transpose=True))
        for spk_float in (False, True):
            B = cases.spikes(9, n_pre, 0.15, float_spikes=spk_float, batch=24)
            kern = fcn._binary_fcnmm_cuda_kernel(transpose=True, weight_info=SDS(w.shape, w.dtype), matrix_info=SDS(B.shape, B.dtype),
                                                 indices_info=SDS(idx.shape, idx.dtype), outs=[SDS((24, n_post), f32)])
            record(f"binary_fcnmm sraw homo={homo} float={spk_float}", lambda: kern(w, idx, B),
                   O.binary_fcnmm(w, idx, B, shape=(n_pre, n_post), transpose=True).T, atol=1e-4)
            # gather, small matrix: unpacked kernels (n_conn <= 32 -> _warp)
            Bg = cases.spikes(10, n_post, 0.15, float_spikes=spk_float, batch=24)
            kern = fcn._binary_fcnmm_cuda_kernel(transpose=False, weight_info=SDS(w.shape, w.dtype), matrix_info=SDS(Bg.shape, Bg.dtype),
                                                 indices_info=SDS(idx.shape, idx.dtype), outs=[SDS((n_pre, 24), f32)])
            record(f"binary_fcnmm gather homo={homo} float={spk_float}", lambda: kern(w, idx, Bg),
                   O.binary_fcnmm(w, idx, Bg, shape=(n_pre, n_post), transpose=False), atol=1e-4)
    # gather, matrix above 1 MB: pack + packed gather chain (n_conn > 32 -> _basic)
    n_pre2, n_conn2, nb = 300, 40, 64
    for spk_float in (False, True):
        n_post2 = 4_200 if spk_float else 16_500        # just above the reference's 1 MB switch to the packed path
        f = cases.fcn_case(11, n_pre2, n_post2, n_conn2)
        w, idx = f["weights"].astype(f32), f["indices"]
        Bg = cases.spikes(12, n_post2, 0.1, float_spikes=spk_float, batch=nb)
        kern = fcn._binary_fcnmm_cuda_kernel(transpose=False, weight_info=SDS(w.shape, w.dtype), matrix_info=SDS(Bg.shape, Bg.dtype),
                                             indices_info=SDS(idx.shape, idx.dtype), outs=[SDS((n_pre2, nb), f32)])
        record(f"binary_fcnmm gather packed float={spk_float}", lambda: kern(w, idx, Bg),
               O.binary_fcnmm(w, idx, Bg, shape=(n_pre2, n_post2), transpose=False), atol=1e-4)
    # ---------------------------------------------------------------- JITC mv / mm, three families
    fam = {"s": ("_jit_scalar.binary", "_binary_jitsmv_cuda_kernel", "_binary_jitsmm_cuda_kernel", (1.5,), O.binary_jitsmv, O.binary_jitsmm),
           "n": ("_jit_normal.binary", "_binary_jitnmv_cuda_kernel", "_binary_jitnmm_cuda_kernel", (1.0, 0.1), O.binary_jitnmv, O.binary_jitnmm),
           "u": ("_jit_uniform.binary", "_binary_jitumv_cuda_kernel", "_binary_jitumm_cuda_kernel", (-0.5, 1.5), O.binary_jitumv, O.binary_jitumm)}
    shape, prob, seed = (900, 1300), 0.02, 77
    clen = np.asarray([O.conn_length(prob)], np.int32)
    seed_a = np.asarray([seed], np.int32)
    for letter, (mname, gmv, gmm, params, omv, omm) in fam.items():
        jm = mod(mname)
        ps = [np.asarray([p], f32) for p in params]
        winfo = {"s": dict(weight_info=SDS((1,), f32)),
                 "n": dict(w_loc_info=SDS((1,), f32), w_scale_info=SDS((1,), f32)),
                 "u": dict(w_low_info=SDS((1,), f32), w_high_info=SDS((1,), f32))}[letter]
        for transpose in (False, True):
            for corder in (True, False):
                vlen = shape[0] if transpose else shape[1]
                olen = shape[1] if transpose else shape[0]
                v = cases.spikes(13, vlen, 0.1)
                kern = getattr(jm, gmv)(corder=corder, vector_info=SDS(v.shape, v.dtype), shape=shape, transpose=transpose,
                                        outs=[SDS((olen,), f32)], **winfo)
                args = (*ps, clen, v, seed_a) + ((None,) if letter == "s" else ())
                # the Normal family's oracle matches the strict normal01 of the reference; its CUDA / numba builds
                # re-associate: tolerance as in tests/test_gpu_parity.py
                record(f"binary_jit{letter}mv T={transpose} corder={corder}", lambda: kern(*args),
                       omv(*params, prob, v, seed, shape=shape, transpose=transpose, corder=corder), atol=2e-4 if letter == "n" else 2e-5)
                B = cases.spikes(14, vlen, 0.1, batch=20)
                kern = getattr(jm, gmm)(corder=corder, B_info=SDS(B.shape, B.dtype), shape=shape, transpose=transpose,
                                        outs=[SDS((olen, 20), f32)], out_info=SDS((olen, 20), f32), **winfo)
                args = (*ps, clen, B, seed_a) + ((None,) if letter == "s" else ())
                record(f"binary_jit{letter}mm T={transpose} corder={corder}", lambda: kern(*args),
                       omm(*params, prob, B, seed, shape=shape, transpose=transpose, corder=corder), atol=3e-4 if letter == "n" else 1e-4)
    # ---------------------------------------------------------------- dense mv / mm
    dn = mod("_dense.binary")
    W = rng.standard_normal((300, 260)).astype(f32)
    for transpose in (True, False):
        for spk_float in (False, True):
            n_src = W.shape[0] if transpose else W.shape[1]
            n_o = W.shape[1] if transpose else W.shape[0]
            v = cases.spikes(15, n_src, 0.2, float_spikes=spk_float)
            kern = dn._binary_densemv_cuda_kernel(spk_info=SDS(v.shape, v.dtype), transpose=transpose, weight_info=SDS(W.shape, W.dtype),
                                                  outs=[SDS((n_o,), f32)])
            record(f"binary_densemv T={transpose} float={spk_float}", lambda: kern(W, v), O.binary_densemv(W, v, transpose=transpose), atol=1e-4)
            B = cases.spikes(16, n_src, 0.2, float_spikes=spk_float, batch=70)
            kern = dn._binary_densemm_cuda_kernel(spk_info=SDS(B.shape, B.dtype), weight_info=SDS(W.shape, W.dtype), transpose=transpose,
                                                  outs=[SDS((n_o, 70), f32)])
            record(f"binary_densemm T={transpose} float={spk_float}", lambda: kern(W, B), O.binary_densemm(W, B, transpose=transpose), atol=1e-4)
    # ---------------------------------------------------------------- compaction
    cp = mod("_event.compact")
    for spk_float in (False, True):
        v = cases.spikes(17, 5000, 0.07, float_spikes=spk_float)
        if spk_float:
            v = np.abs(v)          # compaction uses `!= 0`; the operators `> 0` (kept consistent for the shared handler)
        kern = cp._binary_1d_array_index_cuda_kernel(spikes_info=SDS(v.shape, v.dtype), outs=[SDS((5000,), np.int32), SDS((1,), np.int32)])
        ids, na = O.binary_1d_array_index(v)
        record(f"binary_1d_array_index float={spk_float} ids", lambda: kern(v), ids, out_index=0, exact=True, sort_ids=True)
        record(f"binary_1d_array_index float={spk_float} count", lambda: kern(v), na, out_index=1, exact=True)
        M = cases.spikes(18, 900, 0.05, float_spikes=spk_float, batch=40)
        if spk_float:
            M = np.abs(M)
        outs2 = [SDS((900, 2), np.uint32), SDS((900,), np.int32), SDS((1,), np.int32)]
        kern = cp._binary_2d_array_index_cuda_kernel(spikes_info=SDS(M.shape, M.dtype), outs=outs2)
        exp = O.binary_2d_array_index(M)
        for j, nm in enumerate(("packed", "ids", "count")):
            record(f"binary_2d_array_index float={spk_float} {nm}", lambda: kern(M), exp[j], out_index=j, exact=True, sort_ids=(j == 1))
        kern = cp._binary_2d_compact_only_cuda_kernel(spikes_info=SDS(M.shape, M.dtype), outs=outs2[1:])
        exp = O.binary_2d_compact_only(M)
        for j, nm in enumerate(("ids", "count")):
            record(f"binary_2d_compact_only float={spk_float} {nm}", lambda: kern(M), exp[j], out_index=j, exact=True, sort_ids=(j == 0))

    # ---------------------------------------------------------------- save
    arrays, meta_calls = {}, []
    for i, c in enumerate(CALLS):
        ins = []
        for j, item in enumerate(c["ins"]):
            if item[0] == "array":
                key = f"c{i}_in{j}"
                arrays[key] = item[1]
                ins.append({"array": key})
            else:
                ins.append({"link": [item[1], item[2]]})
        meta_calls.append(dict(target=c["target"], ins=ins, outs=c["outs"], aliases=c["aliases"], attrs=c["attrs"]))
    meta_cases = []
    for i, c in enumerate(CASES):
        arrays[f"case{i}_expected"] = c["expected"]
        meta_cases.append({k: c[k] for k in ("name", "first", "last", "result", "transposed", "rtol", "atol", "exact", "sort_ids")})
    meta = dict(calls=meta_calls, cases=meta_cases, modules=sorted(set(MODULES)),
                generated_by="tests/golden/make_ffi_calls.py from the unmodified reference kernel generators")
    # identical operand arrays are stored once
    seen, alias = {}, {}
    for k in list(arrays):
        a = arrays[k]
        h = (a.dtype.str, a.shape, a.tobytes())
        if h in seen:
            alias[k] = seen[h]
            del arrays[k]
        else:
            seen[h] = k
    meta["array_alias"] = alias
    out = os.path.join(ROOT, "tests", "golden", "ffi_calls.npz")
    np.savez_compressed(out, __meta__=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)
    targets = sorted({c["target"] for c in CALLS})
    print(f"{len(CASES)} cases, {len(CALLS)} ffi calls, {len(targets)} distinct targets, {len(set(MODULES))} modules -> {out} "
          f"({os.path.getsize(out) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
