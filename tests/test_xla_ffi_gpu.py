"""Exercise the header-less XLA-FFI handlers with a hand-built XLA_FFI_CallFrame (jax is not
installable in this image): a fake XLA_FFI_Api table supplies the stream, scratch memory and error
objects, exactly the four services the handlers use."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import oracle as O  # noqa: E402
from tests import cases  # noqa: E402


class ExtBase(C.Structure):
    pass


ExtBase._fields_ = [("struct_size", C.c_size_t), ("type", C.c_int), ("next", C.POINTER(ExtBase))]


class ApiVersion(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("major", C.c_int), ("minor", C.c_int)]


class Metadata(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("api_version", ApiVersion), ("traits", C.c_uint32), ("state_type_id", C.c_int64)]


class MetadataExt(C.Structure):
    _fields_ = [("base", ExtBase), ("metadata", C.POINTER(Metadata))]


class Buffer(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("dtype", C.c_int), ("data", C.c_void_p),
                ("rank", C.c_int64), ("dims", C.POINTER(C.c_int64))]


class ByteSpan(C.Structure):
    _fields_ = [("ptr", C.c_char_p), ("len", C.c_size_t)]


class Scalar(C.Structure):
    _fields_ = [("dtype", C.c_int), ("value", C.c_void_p)]


class Args(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("size", C.c_int64),
                ("types", C.POINTER(C.c_int)), ("args", C.POINTER(C.c_void_p))]


class Attrs(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("size", C.c_int64),
                ("types", C.POINTER(C.c_int)), ("names", C.POINTER(C.POINTER(ByteSpan))), ("attrs", C.POINTER(C.c_void_p))]


class ErrArgs(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("message", C.c_char_p), ("errc", C.c_int)]


class StreamArgs(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("ctx", C.c_void_p), ("stream", C.c_void_p)]


class AllocArgs(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("ctx", C.c_void_p), ("size", C.c_size_t),
                ("alignment", C.c_size_t), ("data", C.c_void_p)]


class FreeArgs(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("ctx", C.c_void_p), ("size", C.c_size_t),
                ("data", C.c_void_p)]


ERR_F = C.CFUNCTYPE(C.c_void_p, C.POINTER(ErrArgs))
STREAM_F = C.CFUNCTYPE(C.c_void_p, C.POINTER(StreamArgs))
ALLOC_F = C.CFUNCTYPE(C.c_void_p, C.POINTER(AllocArgs))
FREE_F = C.CFUNCTYPE(C.c_void_p, C.POINTER(FreeArgs))


class Api(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.c_void_p), ("api_version", ApiVersion),
                ("internal_api", C.c_void_p), ("Error_Create", ERR_F), ("Error_GetMessage", C.c_void_p),
                ("Error_Destroy", C.c_void_p), ("Handler_Register", C.c_void_p), ("Stream_Get", STREAM_F),
                ("Type_Register", C.c_void_p), ("ExecutionContext_Get", C.c_void_p), ("State_Set", C.c_void_p),
                ("State_Get", C.c_void_p), ("DeviceMemory_Allocate", ALLOC_F), ("DeviceMemory_Free", FREE_F),
                ("ThreadPool_Schedule", C.c_void_p), ("ThreadPool_NumThreads", C.c_void_p), ("Future_Create", C.c_void_p),
                ("Future_SetAvailable", C.c_void_p), ("Future_SetError", C.c_void_p), ("RunId_Get", C.c_void_p),
                ("DeviceOrdinal_Get", C.c_void_p)]


class CallFrame(C.Structure):
    _fields_ = [("struct_size", C.c_size_t), ("extension_start", C.POINTER(ExtBase)), ("api", C.POINTER(Api)),
                ("ctx", C.c_void_p), ("stage", C.c_int), ("args", Args), ("rets", Args), ("attrs", Attrs),
                ("future", C.c_void_p)]


XLA = {torch.bool: 1, torch.int8: 2, torch.int32: 4, torch.int64: 5, torch.uint8: 6, torch.float16: 10,
       torch.float32: 11, torch.float64: 12, torch.bfloat16: 16}


class FakeXla:
    def __init__(self):
        self.errors, self.live, self.keep = [], {}, []

        def err(a):
            self.errors.append((a.contents.errc, a.contents.message.decode()))
            return 0xdead

        def stream(a):
            a.contents.stream = torch.cuda.current_stream().cuda_stream
            return None

        def alloc(a):
            t = torch.empty(a.contents.size + 256, dtype=torch.uint8, device="cuda")
            p = (t.data_ptr() + 255) // 256 * 256
            self.live[p] = t
            a.contents.data = p
            return None

        def free(a):
            self.live.pop(a.contents.data, None)
            return None

        self.cb = (ERR_F(err), STREAM_F(stream), ALLOC_F(alloc), FREE_F(free))
        self.api = Api()
        self.api.struct_size = C.sizeof(Api)
        self.api.api_version = ApiVersion(C.sizeof(ApiVersion), None, 0, 1)
        self.api.Error_Create, self.api.Stream_Get, self.api.DeviceMemory_Allocate, self.api.DeviceMemory_Free = self.cb

    def _buffers(self, tensors, packed_u32=()):
        n = len(tensors)
        types = (C.c_int * n)(*([1] * n))
        ptrs = (C.c_void_p * n)()
        self.keep += list(tensors)   # the frame only stores raw pointers: keep the tensors alive
        for i, t in enumerate(tensors):
            dims = (C.c_int64 * max(t.ndim, 1))(*t.shape)
            dt = 8 if i in packed_u32 else XLA[t.dtype]
            b = Buffer(C.sizeof(Buffer), None, dt, t.data_ptr(), t.ndim, dims)
            self.keep += [dims, b]
            ptrs[i] = C.addressof(b)
        self.keep += [types, ptrs]
        return Args(C.sizeof(Args), None, n, types, ptrs)

    def frame(self, ins, outs, attrs, packed_u32=(), packed_u32_out=()):
        f = CallFrame()
        f.struct_size = C.sizeof(CallFrame)
        f.api = C.pointer(self.api)
        f.stage = 3
        f.args = self._buffers(ins, packed_u32)
        f.rets = self._buffers(outs, packed_u32_out)
        names = sorted(attrs)
        n = len(names)
        types = (C.c_int * max(n, 1))(*([3] * n))
        nptr = (C.POINTER(ByteSpan) * max(n, 1))()
        aptr = (C.c_void_p * max(n, 1))()
        for i, k in enumerate(names):
            bs = ByteSpan(k.encode(), len(k))
            if isinstance(attrs[k], float):
                val = C.c_float(attrs[k])
                sc = Scalar(11, C.addressof(val))       # F32 scalar attribute
            else:
                val = C.c_int32(int(attrs[k]))
                sc = Scalar(4, C.addressof(val))
            self.keep += [bs, val, sc]
            nptr[i] = C.pointer(bs)
            aptr[i] = C.addressof(sc)
        self.keep += [types, nptr, aptr]
        f.attrs = Attrs(C.sizeof(Attrs), None, n, types, nptr, aptr)
        return f


@pytest.fixture(scope="module")
def lib():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from brainevent_b200 import _lib
    L = C.CDLL(_lib.LIB_PATH)
    return L


def call(L, name, fx, frame):
    fn = getattr(L, name)
    fn.restype = C.c_void_p
    fn.argtypes = [C.POINTER(CallFrame)]
    r = fn(C.byref(frame))
    torch.cuda.synchronize()
    return r


def t(x):
    return torch.as_tensor(np.ascontiguousarray(x), device="cuda")


def test_metadata_handshake(lib):
    fx = FakeXla()
    md = Metadata()
    md.struct_size = C.sizeof(Metadata)
    ext = MetadataExt(ExtBase(C.sizeof(MetadataExt), 1, None), C.pointer(md))
    f = fx.frame([], [], {})
    f.extension_start = C.cast(C.pointer(ext), C.POINTER(ExtBase))
    assert call(lib, "be_b200_xla_binary_fcnmv", fx, f) is None
    assert (md.api_version.major, md.api_version.minor, md.traits) == (0, 1, 1)


def test_fcnmv_and_csrmv_frames(lib):
    fx = FakeXla()
    f = cases.fcn_case(1, 2000, 1500, 40)
    s = cases.spikes(2, 2000, 0.1)
    out = torch.empty(1500, device="cuda")
    fr = fx.frame([t(f["weights"]), t(f["indices"]), t(s)], [out], {"transpose": 1})
    assert call(lib, "be_b200_xla_binary_fcnmv", fx, fr) is None, fx.errors
    np.testing.assert_allclose(out.cpu().numpy(), O.binary_fcnmv(f["weights"], f["indices"], s, shape=(2000, 1500), transpose=True),
                               rtol=1e-5, atol=1e-5)
    assert not fx.live           # scratch returned to the allocator
    c = cases.csr_case(3, 1200, 900, 20, 0.1, ragged=True)
    for tr in (0, 1):
        v = cases.spikes(4, 1200 if tr else 900, 0.1, float_spikes=bool(tr))
        out = torch.empty(900 if tr else 1200, device="cuda")
        fr = fx.frame([t(c["data"]), t(c["indices"]), t(c["indptr"]), t(v)], [out], {"transpose": tr})
        assert call(lib, "be_b200_xla_binary_csrmv", fx, fr) is None, fx.errors
        np.testing.assert_allclose(out.cpu().numpy(), O.binary_csrmv(c["data"], c["indices"], c["indptr"], v, shape=(1200, 900),
                                                                     transpose=bool(tr)), rtol=1e-5, atol=1e-5)
    B = cases.spikes(5, 1200, 0.05, batch=64)
    out = torch.empty((900, 64), device="cuda")
    fr = fx.frame([t(c["data"]), t(c["indices"]), t(c["indptr"]), t(B)], [out], {"transpose": 1})
    assert call(lib, "be_b200_xla_binary_csrmm", fx, fr) is None, fx.errors
    np.testing.assert_allclose(out.cpu().numpy(), O.binary_csrmm(c["data"], c["indices"], c["indptr"], B, shape=(1200, 900),
                                                                 transpose=True), rtol=1e-5, atol=1e-5)


def test_jit_dense_compact_frames(lib):
    fx = FakeXla()
    shape, prob, seed = (700, 900), 0.02, 11
    v = cases.spikes(6, 700, 0.2)
    for kind, params, ofn in ((0, [1.5], O.binary_jitsmv), (1, [1.0, 0.1], O.binary_jitnmv), (2, [-0.5, 1.0], O.binary_jitumv)):
        for corder in (0, 1):
            out = torch.empty(900, device="cuda")
            ins = [torch.tensor([p], device="cuda") for p in params] + [
                torch.tensor([O.conn_length(prob)], dtype=torch.int32, device="cuda"),
                torch.tensor([seed], dtype=torch.int32, device="cuda"), t(v)]
            fr = fx.frame(ins, [out], {"kind": kind, "corder": corder, "chunk_size": O.normalize_chunk_size(900)})
            assert call(lib, "be_b200_xla_binary_jitmv", fx, fr) is None, fx.errors
            np.testing.assert_allclose(out.cpu().numpy(), ofn(*params, prob, v, seed, shape=shape, transpose=True, corder=bool(corder)),
                                       rtol=1e-5, atol=2e-5)
    W = np.random.default_rng(0).standard_normal((300, 400)).astype(np.float32)
    sv = cases.spikes(7, 300, 0.1)
    out = torch.empty(400, device="cuda")
    assert call(lib, "be_b200_xla_binary_densemv", fx, fx.frame([t(W), t(sv)], [out], {"transpose": 1})) is None, fx.errors
    np.testing.assert_allclose(out.cpu().numpy(), O.binary_densemv(W, sv, transpose=True), rtol=1e-4, atol=1e-4)
    s1 = cases.spikes(8, 5000, 0.03)
    ids, na = torch.empty(5000, dtype=torch.int32, device="cuda"), torch.empty(1, dtype=torch.int32, device="cuda")
    assert call(lib, "be_b200_xla_compact_1d", fx, fx.frame([t(s1)], [ids, na], {})) is None, fx.errors
    oi, on = O.binary_1d_array_index(s1)
    assert int(na.item()) == int(on[0])
    np.testing.assert_array_equal(ids.cpu().numpy(), oi)
    s2 = cases.spikes(9, 700, 0.02, batch=70)
    pk = torch.empty((700, 3), dtype=torch.int32, device="cuda")
    ids, na = torch.empty(700, dtype=torch.int32, device="cuda"), torch.empty(1, dtype=torch.int32, device="cuda")
    assert call(lib, "be_b200_xla_compact_2d", fx, fx.frame([t(s2)], [pk, ids, na], {})) is None, fx.errors
    op, oi, on = O.binary_2d_array_index(s2)
    np.testing.assert_array_equal(pk.cpu().numpy().view(np.uint32), op)
    np.testing.assert_array_equal(ids.cpu().numpy(), oi)
    # packed (u32) events straight into the scatter: BitPackedBinary through the FFI
    f = cases.fcn_case(10, 2048, 1500, 16)
    s = cases.spikes(11, 2048, 0.1)
    words = t(O.bitpack(s, 0).view(np.int32))
    out = torch.empty(1500, device="cuda")
    fr = fx.frame([t(f["weights"]), t(f["indices"]), words], [out], {"transpose": 1}, packed_u32=(2,))
    assert call(lib, "be_b200_xla_binary_fcnmv", fx, fr) is None, fx.errors
    np.testing.assert_allclose(out.cpu().numpy(), O.binary_fcnmv(f["weights"], f["indices"], s, shape=(2048, 1500), transpose=True),
                               rtol=1e-5, atol=1e-5)


def test_errors_become_ffi_errors(lib):
    fx = FakeXla()
    out = torch.empty(10, device="cuda")
    fr = fx.frame([torch.ones(3, device="cuda")], [out], {})
    assert call(lib, "be_b200_xla_binary_fcnmv", fx, fr) == 0xdead
    assert fx.errors and fx.errors[-1][0] == 3


def test_next_row_frames(lib):
    """float ops, plasticity and the 2-D encoders through hand-built call frames."""
    fx = FakeXla()
    rng = np.random.default_rng(20)
    c = cases.csr_case(21, 1200, 900, 20, 0.1, ragged=True)
    for tr in (0, 1):
        x = rng.standard_normal(1200 if tr else 900).astype(np.float32)
        out = torch.empty(900 if tr else 1200, device="cuda")
        fr = fx.frame([t(c["data"]), t(c["indices"]), t(c["indptr"]), t(x)], [out], {"transpose": tr})
        assert call(lib, "be_b200_xla_csrmv", fx, fr) is None, fx.errors
        np.testing.assert_allclose(out.cpu().numpy(), O.csrmv(c["data"], c["indices"], c["indptr"], x, shape=(1200, 900), transpose=bool(tr)),
                                   rtol=1e-4, atol=1e-4)
    X = rng.standard_normal((900, 12)).astype(np.float32)
    out = torch.empty((1200, 12), device="cuda")
    assert call(lib, "be_b200_xla_csrmm", fx, fx.frame([t(c["data"]), t(c["indices"]), t(c["indptr"]), t(X)], [out], {"transpose": 0})) is None, fx.errors
    np.testing.assert_allclose(out.cpu().numpy(), O.csrmm(c["data"], c["indices"], c["indptr"], X, shape=(1200, 900)), rtol=1e-4, atol=1e-4)
    f = cases.fcn_case(22, 2000, 1500, 40)
    x = rng.standard_normal(2000).astype(np.float32)
    out = torch.empty(1500, device="cuda")
    assert call(lib, "be_b200_xla_fcnmv", fx, fx.frame([t(f["weights"]), t(f["indices"]), t(x)], [out], {"transpose": 1})) is None, fx.errors
    np.testing.assert_allclose(out.cpu().numpy(), O.fcnmv(f["weights"], f["indices"], x, shape=(2000, 1500), transpose=True), rtol=1e-4, atol=1e-4)
    xv = rng.standard_normal(700).astype(np.float32)
    out = torch.empty(900, device="cuda")
    ins = [torch.tensor([1.0], device="cuda"), torch.tensor([0.1], device="cuda"),
           torch.tensor([O.conn_length(0.02)], dtype=torch.int32, device="cuda"), torch.tensor([11], dtype=torch.int32, device="cuda"), t(xv)]
    fr = fx.frame(ins, [out], {"kind": 1, "corder": 0, "chunk_size": O.normalize_chunk_size(900)})
    assert call(lib, "be_b200_xla_jitmv", fx, fr) is None, fx.errors
    np.testing.assert_allclose(out.cpu().numpy(), O.jitnmv(1.0, 0.1, 0.02, xv, 11, shape=(700, 900), transpose=True, corder=False), rtol=1e-4, atol=1e-4)
    # plasticity: aliased (in place) and non-aliased output, bounds as float attributes
    pre = cases.spikes(23, 1200, 0.2)
    tr_post = rng.standard_normal(900).astype(np.float32)
    want = O.update_csr_on_binary_pre(c["data"], c["indices"], c["indptr"], pre, tr_post, -0.5, 0.75, shape=(1200, 900))
    w_in, w_out = t(c["data"]), torch.empty(c["data"].shape[0], device="cuda")
    fr = fx.frame([w_in, t(c["indices"]), t(c["indptr"]), t(pre), t(tr_post)], [w_out], {"w_min": -0.5, "w_max": 0.75})
    assert call(lib, "be_b200_xla_plasticity_rows", fx, fr) is None, fx.errors
    np.testing.assert_array_equal(w_out.cpu().numpy(), want)
    np.testing.assert_array_equal(w_in.cpu().numpy(), c["data"])
    fr = fx.frame([w_in, t(c["indices"]), t(c["indptr"]), t(pre), t(tr_post)], [w_in], {"w_min": -0.5, "w_max": 0.75})
    assert call(lib, "be_b200_xla_plasticity_rows", fx, fr) is None, fx.errors
    np.testing.assert_array_equal(w_in.cpu().numpy(), want)
    fw = t(f["weights"])
    spk = cases.spikes(24, 2000, 0.1)
    trf = rng.standard_normal(1500).astype(np.float32)
    fr = fx.frame([fw, t(f["indices"]), t(spk), t(trf)], [fw], {"has_indptr": 0})
    assert call(lib, "be_b200_xla_plasticity_rows", fx, fr) is None, fx.errors
    np.testing.assert_array_equal(fw.cpu().numpy(), O.fcn_plasticity_row(f["weights"], f["indices"], spk, trf))
    # encoders
    S = cases.spikes(25, 300, 0.05, batch=70)
    idx, ptr_ = torch.empty(300 * 70, dtype=torch.int32, device="cuda"), torch.empty(301, dtype=torch.int32, device="cuda")
    assert call(lib, "be_b200_xla_binary_2d_csr_encode", fx, fx.frame([t(S)], [idx, ptr_], {})) is None, fx.errors
    oi, op = O.binary_2d_csr_encode(S)
    np.testing.assert_array_equal(idx.cpu().numpy(), oi); np.testing.assert_array_equal(ptr_.cpu().numpy(), op)
    idx, ptr_ = torch.empty(300 * 70, dtype=torch.int32, device="cuda"), torch.empty(71, dtype=torch.int32, device="cuda")
    assert call(lib, "be_b200_xla_binary_2d_csc_encode", fx, fx.frame([t(S)], [idx, ptr_], {})) is None, fx.errors
    oi, op = O.binary_2d_csc_encode(S)
    np.testing.assert_array_equal(idx.cpu().numpy(), oi); np.testing.assert_array_equal(ptr_.cpu().numpy(), op)
    pairs, npairs = torch.empty((300 * 70, 2), dtype=torch.int32, device="cuda"), torch.empty(1, dtype=torch.int32, device="cuda")
    assert call(lib, "be_b200_xla_binary_2d_pair_stream_encode", fx, fx.frame([t(S)], [pairs, npairs], {})) is None, fx.errors
    op_, on = O.binary_2d_pair_stream_encode(S)
    np.testing.assert_array_equal(pairs.cpu().numpy(), op_); assert int(npairs.item()) == int(on[0])
    rs, mx = torch.empty((300, 40), dtype=torch.int32, device="cuda"), torch.empty(1, dtype=torch.int32, device="cuda")
    assert call(lib, "be_b200_xla_binary_2d_row_sparse_encode", fx, fx.frame([t(S)], [rs, mx], {})) is None, fx.errors
    np.testing.assert_array_equal(rs.cpu().numpy(), O.binary_2d_row_sparse_encode(S, 40))
    assert int(mx.item()) == int((S != 0).sum(axis=1).max())
    assert not fx.live
