"""SURVEY 8(b) last row: the reference-named FFI symbols, driven by the REFERENCE's own argument preparation.

tests/golden/ffi_calls.npz holds what the unmodified reference cuda_raw kernel generators passed to
`jax.ffi.ffi_call` for every hot-path primitive (recorded in the build container by tests/golden/make_ffi_calls.py:
target names, operand arrays in the reference's order, result specs, input_output_aliases, scalar attributes,
pack -> kernel data flow) plus the oracle's result for each primitive.  Here every recorded call is replayed on the
B200: the target is resolved exactly as integration/brainevent_b200_backend.py would register it, the operands
become device buffers of a hand-built XLA_FFI_CallFrame (results pre-filled with garbage, aliased results sharing
their operand's buffer like XLA does), and the primitive's result is compared with the oracle."""
import ctypes as C
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from tests.test_xla_ffi_gpu import CallFrame, FakeXla  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_TORCH = {"float32": torch.float32, "float64": torch.float64, "int32": torch.int32, "int64": torch.int64,
          "int8": torch.int8, "uint8": torch.uint8, "bool": torch.bool, "uint32": torch.int32}


def _load():
    z = np.load(os.path.join(ROOT, "tests", "golden", "ffi_calls.npz"))
    meta = json.loads(bytes(z["__meta__"]).decode())
    return z, meta


Z, META = _load()


def _array(key):
    return Z[META["array_alias"].get(key, key)]


@pytest.fixture(scope="module")
def backend():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from integration import brainevent_b200_backend as B
    return B


@pytest.mark.parametrize("ci", range(len(META["cases"])), ids=[c["name"].replace(" ", "_") for c in META["cases"]])
def test_replay_reference_ffi_calls(backend, ci):
    case = META["cases"][ci]
    fx = FakeXla()
    produced = {}
    for i in range(case["first"], case["last"]):
        call = META["calls"][i]
        sym = backend.symbol_for_target(call["target"])
        assert sym is not None, call["target"]
        ins, u32_in = [], []
        for j, item in enumerate(call["ins"]):
            if "array" in item:
                a = _array(item["array"])
                if a.dtype == np.uint32:
                    a, _ = a.view(np.int32), u32_in.append(j)
                ins.append(torch.as_tensor(np.ascontiguousarray(a), device="cuda"))
            else:
                src = produced[tuple(item["link"])]
                if src[1]:
                    u32_in.append(j)
                ins.append(src[0])
        outs, u32_out = [], []
        alias_of = {int(v): int(k) for k, v in call["aliases"].items()}          # result index -> operand index
        for j, (shape, dt) in enumerate(call["outs"]):
            if j in alias_of:
                outs.append(ins[alias_of[j]])                           # XLA hands the operand's buffer back
            else:
                o = torch.empty(tuple(shape), dtype=_TORCH[dt], device="cuda")
                if o.dtype.is_floating_point:
                    o.fill_(float("nan"))                               # results are uninitialised on entry
                elif o.numel():
                    o.fill_(-7 if o.dtype != torch.bool else 1)
                outs.append(o)
            if dt == "uint32":
                u32_out.append(j)
        attrs = {k: v[0] for k, v in call["attrs"].items()}
        frame = fx.frame(ins, outs, attrs, packed_u32=tuple(u32_in), packed_u32_out=tuple(u32_out))
        fn = getattr(backend.lib(), sym)
        fn.restype = C.c_void_p
        fn.argtypes = [C.POINTER(CallFrame)]
        rc = fn(C.byref(frame))
        torch.cuda.synchronize()
        assert rc is None, (call["target"], fx.errors)
        for j, o in enumerate(outs):
            produced[(i, j)] = (o, call["outs"][j][1] == "uint32")
    assert not fx.live, "scratch not returned to the allocator"
    got = produced[tuple(case["result"])][0].cpu().numpy()
    if case["transposed"]:
        got = got.T
    want = _array(f"case{ci}_expected")
    if want.dtype == np.uint32:
        got = got.view(np.uint32)
    if case["exact"]:
        if case["sort_ids"]:
            # the reference's CUDA compaction order is unspecified; ours is ascending = the oracle's
            n = int((want != 0).sum()) + (1 if want.size and want[0] == 0 and (want != 0).any() else 0)
            np.testing.assert_array_equal(np.sort(got[:n]), np.sort(want[:n]))
        else:
            np.testing.assert_array_equal(got.reshape(want.shape), want)
    else:
        np.testing.assert_allclose(got, want.reshape(got.shape), rtol=case["rtol"], atol=case["atol"])
