// xla_ffi.cu -- header-less XLA-FFI handlers over the plain C ABI (see brainevent_b200_xla.h).
#include "common.cuh"
#include "../../include/brainevent_b200_xla.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>

namespace {

struct Frame {
    XLA_FFI_CallFrame* f;
    std::string err;
    void* stream = nullptr;
    void* scratch = nullptr;
    size_t scratch_bytes = 0;

    explicit Frame(XLA_FFI_CallFrame* frame) : f(frame) {}

    XLA_FFI_Error* fail(int code, const std::string& msg) {
        XLA_FFI_Error_Create_Args a;
        a.struct_size = sizeof(a); a.extension_start = nullptr; a.message = msg.c_str(); a.errc = code;
        return f->api->XLA_FFI_Error_Create(&a);
    }
    // metadata handshake: returns true when the frame was a metadata query (handled)
    bool metadata() {
        for (XLA_FFI_Extension_Base* e = f->extension_start; e; e = e->next) {
            if (e->type == 1) {
                XLA_FFI_Metadata* m = reinterpret_cast<XLA_FFI_Metadata_Extension*>(e)->metadata;
                m->api_version.major_version = f->api ? f->api->api_version.major_version : 0;
                m->api_version.minor_version = f->api ? f->api->api_version.minor_version : 1;
                m->traits = 1u;  // COMMAND_BUFFER_COMPATIBLE: memsets + kernel launches only
                return true;
            }
        }
        return false;
    }
    XLA_FFI_Buffer* arg(int i) const { return i < f->args.size ? (XLA_FFI_Buffer*)f->args.args[i] : nullptr; }
    XLA_FFI_Buffer* ret(int i) const { return i < f->rets.size ? (XLA_FFI_Buffer*)f->rets.rets[i] : nullptr; }
    bool attr_i64(const char* name, int64_t* out) const {
        for (int64_t i = 0; i < f->attrs.size; ++i) {
            XLA_FFI_ByteSpan* n = f->attrs.names[i];
            if (n->len == strlen(name) && memcmp(n->ptr, name, n->len) == 0 && f->attrs.types[i] == 3) {
                XLA_FFI_Scalar* s = (XLA_FFI_Scalar*)f->attrs.attrs[i];
                switch (s->dtype) {
                    case BE_XLA_PRED: case BE_XLA_S8: case BE_XLA_U8: *out = *(int8_t*)s->value; return true;
                    case BE_XLA_S32: case BE_XLA_U32: *out = *(int32_t*)s->value; return true;
                    case BE_XLA_S64: *out = *(int64_t*)s->value; return true;
                    default: return false;
                }
            }
        }
        return false;
    }
    bool attr_f64(const char* name, double* out) const {
        for (int64_t i = 0; i < f->attrs.size; ++i) {
            XLA_FFI_ByteSpan* n = f->attrs.names[i];
            if (n->len == strlen(name) && memcmp(n->ptr, name, n->len) == 0 && f->attrs.types[i] == 3) {
                XLA_FFI_Scalar* s = (XLA_FFI_Scalar*)f->attrs.attrs[i];
                if (s->dtype == BE_XLA_F32) { *out = *(float*)s->value; return true; }
                if (s->dtype == BE_XLA_F64) { *out = *(double*)s->value; return true; }
                int64_t v;
                if (attr_i64(name, &v)) { *out = (double)v; return true; }
                return false;
            }
        }
        return false;
    }
    XLA_FFI_Error* alloc_bytes(size_t bytes, void** out) {     // extra device scratch (freed by the caller)
        XLA_FFI_DeviceMemory_Allocate_Args a;
        a.struct_size = sizeof(a); a.extension_start = nullptr; a.ctx = f->ctx;
        a.size = bytes; a.alignment = 256; a.data = nullptr;
        XLA_FFI_Error* e = f->api->XLA_FFI_DeviceMemory_Allocate(&a);
        *out = a.data;
        return e;
    }
    void free_bytes(size_t bytes, void* p) {
        XLA_FFI_DeviceMemory_Free_Args a;
        a.struct_size = sizeof(a); a.extension_start = nullptr; a.ctx = f->ctx; a.size = bytes; a.data = p;
        f->api->XLA_FFI_DeviceMemory_Free(&a);
    }
    XLA_FFI_Error* get_stream() {
        XLA_FFI_Stream_Get_Args a;
        a.struct_size = sizeof(a); a.extension_start = nullptr; a.ctx = f->ctx; a.stream = nullptr;
        XLA_FFI_Error* e = f->api->XLA_FFI_Stream_Get(&a);
        stream = a.stream;
        return e;
    }
    // zero-initialised workspace from XLA's scratch allocator (stream-ordered re-use is XLA's job)
    XLA_FFI_Error* get_scratch(int64_t n_src, int64_t n_batch) {
        scratch_bytes = be_b200_workspace_bytes(n_src, n_batch);
        XLA_FFI_DeviceMemory_Allocate_Args a;
        a.struct_size = sizeof(a); a.extension_start = nullptr; a.ctx = f->ctx;
        a.size = scratch_bytes; a.alignment = 256; a.data = nullptr;
        XLA_FFI_Error* e = f->api->XLA_FFI_DeviceMemory_Allocate(&a);
        if (e) return e;
        scratch = a.data;
        // only the control words need zeroing: the 256-byte header and the tile-status tail
        be::Workspace ws = be::ws_carve(scratch, n_src, n_batch);
        if (cudaMemsetAsync(ws.hdr, 0, sizeof(be::WsHeader), (cudaStream_t)stream) != cudaSuccess ||
            cudaMemsetAsync(ws.status, 0, (size_t)((char*)scratch + scratch_bytes - (char*)ws.status), (cudaStream_t)stream) != cudaSuccess)
            return fail(BE_B200_EINTERNAL, "workspace memset failed");
        return nullptr;
    }
    void release() {
        if (!scratch) return;
        XLA_FFI_DeviceMemory_Free_Args a;
        a.struct_size = sizeof(a); a.extension_start = nullptr; a.ctx = f->ctx; a.size = scratch_bytes; a.data = scratch;
        f->api->XLA_FFI_DeviceMemory_Free(&a);
        scratch = nullptr;
    }
    XLA_FFI_Error* finish(int rc, const char* what) {
        release();
        if (rc == BE_B200_OK) return nullptr;
        return fail(rc, std::string(what) + ": " + be_b200_last_error());
    }
};

int wt_of(int xla_dtype) {
    switch (xla_dtype) {
        case BE_XLA_F32: return BE_B200_F32;
        case BE_XLA_F64: return BE_B200_F64;
        case BE_XLA_F16: return BE_B200_F16;
        case BE_XLA_BF16: return BE_B200_BF16;
        default: return -1;
    }
}
int spk_of(int xla_dtype) {
    switch (xla_dtype) {
        case BE_XLA_PRED: case BE_XLA_S8: case BE_XLA_U8: return BE_B200_SPK_BOOL;
        case BE_XLA_F32: return BE_B200_SPK_F32;
        case BE_XLA_U32: return BE_B200_SPK_PACKED;
        default: return -1;
    }
}
int64_t numel(const XLA_FFI_Buffer* b) {
    int64_t n = 1;
    for (int64_t i = 0; i < b->rank; ++i) n *= b->dims[i];
    return n;
}

#define BE_XLA_PROLOGUE(NARGS, NRETS, NAME)                                                          \
    Frame fr(frame);                                                                                 \
    if (fr.metadata()) return nullptr;                                                               \
    if (frame->stage != 3) return nullptr; /* only EXECUTE does work */                              \
    if (frame->args.size != (NARGS) || frame->rets.size < (NRETS))                                   \
        return fr.fail(BE_B200_EINVAL, NAME ": wrong number of buffers");                            \
    if (XLA_FFI_Error* e__ = fr.get_stream()) return e__;

XLA_FFI_Error* csr_common(XLA_FFI_CallFrame* frame, bool matmat, bool indexed, const char* name) {
    const int na = indexed ? 5 : 4;
    BE_XLA_PROLOGUE(na, 1, "csr");
    XLA_FFI_Buffer *w = fr.arg(0), *idx = fr.arg(1), *ptr = fr.arg(2), *perm = indexed ? fr.arg(3) : nullptr,
                   *ev = fr.arg(na - 1), *out = fr.ret(0);
    int64_t transpose = 0, bm = 0;
    fr.attr_i64("transpose", &transpose);
    fr.attr_i64("out_batch_major", &bm);
    const int wt = wt_of(w->dtype), sk = spk_of(ev->dtype);
    if (wt < 0 || sk < 0) return fr.fail(BE_B200_EUNIMPL, std::string(name) + ": unsupported dtype");
    if (idx->dtype != BE_XLA_S32) return fr.fail(BE_B200_EINVAL, std::string(name) + ": indices must be int32");
    const int64_t m = ptr->dims[0] - 1, nse = numel(idx);
    const int64_t n = matmat ? out->dims[bm ? 0 : 1] : 0;
    const int64_t k = transpose ? out->dims[bm ? 1 : 0] : (sk == BE_B200_SPK_PACKED && !matmat ? -1 : ev->dims[0]);
    if (k < 0) return fr.fail(BE_B200_EINVAL, std::string(name) + ": packed events need transpose=1 here");
    const int64_t n_src = transpose ? m : k;
    if (XLA_FFI_Error* e = fr.get_scratch(n_src, matmat ? n : 1)) return e;
    int rc;
    if (matmat)
        rc = be_b200_binary_csrmm(w->data, wt, numel(w) == 1, (const int32_t*)idx->data, ptr->data, ptr->dtype == BE_XLA_S64,
                                  perm ? perm->data : nullptr, perm && perm->dtype == BE_XLA_S64, ev->data, nullptr, sk, m, k,
                                  nse, n, (int)transpose, (int)bm, out->data, fr.scratch, fr.scratch_bytes, fr.stream);
    else
        rc = be_b200_binary_csrmv(w->data, wt, numel(w) == 1, (const int32_t*)idx->data, ptr->data, ptr->dtype == BE_XLA_S64,
                                  perm ? perm->data : nullptr, perm && perm->dtype == BE_XLA_S64, ev->data, nullptr, sk, m, k,
                                  nse, (int)transpose, out->data, fr.scratch, fr.scratch_bytes, fr.stream);
    return fr.finish(rc, name);
}

XLA_FFI_Error* fcn_common(XLA_FFI_CallFrame* frame, bool matmat, const char* name) {
    BE_XLA_PROLOGUE(3, 1, "fcn");
    XLA_FFI_Buffer *w = fr.arg(0), *idx = fr.arg(1), *ev = fr.arg(2), *out = fr.ret(0);
    int64_t transpose = 0, bm = 0;
    fr.attr_i64("transpose", &transpose);
    fr.attr_i64("out_batch_major", &bm);
    const int wt = wt_of(w->dtype), sk = spk_of(ev->dtype);
    if (wt < 0 || sk < 0) return fr.fail(BE_B200_EUNIMPL, std::string(name) + ": unsupported dtype");
    if (idx->dtype != BE_XLA_S32 || idx->rank != 2) return fr.fail(BE_B200_EINVAL, std::string(name) + ": indices must be int32 [n_rows, n_conn]");
    const int64_t n_rows = idx->dims[0], n_conn = idx->dims[1], n_out = out->dims[(matmat && bm) ? 1 : 0];
    const int64_t n = matmat ? out->dims[bm ? 0 : 1] : 0;
    int64_t n_src = transpose ? n_rows : ev->dims[0];
    if (!transpose && sk == BE_B200_SPK_PACKED && !matmat) {
        if (!fr.attr_i64("n_src", &n_src)) return fr.fail(BE_B200_EINVAL, std::string(name) + ": packed events need attr n_src");
    }
    if (XLA_FFI_Error* e = fr.get_scratch(n_src, matmat ? n : 1)) return e;
    int rc;
    if (matmat)
        rc = be_b200_binary_fcnmm(w->data, wt, numel(w) == 1, (const int32_t*)idx->data, n_rows, n_conn, ev->data, nullptr, sk,
                                  n_src, n_out, n, (int)transpose, (int)bm, out->data, fr.scratch, fr.scratch_bytes, fr.stream);
    else
        rc = be_b200_binary_fcnmv(w->data, wt, numel(w) == 1, (const int32_t*)idx->data, n_rows, n_conn, ev->data, nullptr, sk,
                                  n_src, n_out, (int)transpose, out->data, fr.scratch, fr.scratch_bytes, fr.stream);
    return fr.finish(rc, name);
}

XLA_FFI_Error* jit_common(XLA_FFI_CallFrame* frame, bool matmat, const char* name) {
    Frame fr(frame);
    if (fr.metadata()) return nullptr;
    if (frame->stage != 3) return nullptr;
    int64_t kind = 0, corder = 1, chunk = 0;
    fr.attr_i64("kind", &kind);
    fr.attr_i64("corder", &corder);
    if (!fr.attr_i64("chunk_size", &chunk)) return fr.fail(BE_B200_EINVAL, std::string(name) + ": attr chunk_size is required");
    const int np = kind == BE_B200_JIT_SCALAR ? 1 : 2;
    if (frame->args.size != np + 3 || frame->rets.size < 1) return fr.fail(BE_B200_EINVAL, std::string(name) + ": wrong number of buffers");
    if (XLA_FFI_Error* e = fr.get_stream()) return e;
    XLA_FFI_Buffer *p0 = fr.arg(0), *p1 = np == 2 ? fr.arg(1) : nullptr, *clen = fr.arg(np), *seed = fr.arg(np + 1),
                   *ev = fr.arg(np + 2), *out = fr.ret(0);
    const int wt = wt_of(p0->dtype), sk = spk_of(ev->dtype);
    if (wt < 0 || sk < 0) return fr.fail(BE_B200_EUNIMPL, std::string(name) + ": unsupported dtype");
    const int64_t out_len = out->dims[0];
    int64_t vec_len = ev->dims[0];
    if (sk == BE_B200_SPK_PACKED && !matmat) {
        if (!fr.attr_i64("vector_size", &vec_len)) return fr.fail(BE_B200_EINVAL, std::string(name) + ": packed events need attr vector_size");
    }
    const int64_t n = matmat ? out->dims[1] : 1;
    if (XLA_FFI_Error* e = fr.get_scratch(vec_len, n)) return e;
    int rc;
    if (matmat)
        rc = be_b200_binary_jitmm((int)kind, p0->data, p1 ? p1->data : nullptr, wt, (const int32_t*)clen->data,
                                  (const int32_t*)seed->data, ev->data, nullptr, sk, vec_len, out_len, n, chunk, (int)corder,
                                  out->data, fr.scratch, fr.scratch_bytes, fr.stream);
    else
        rc = be_b200_binary_jitmv((int)kind, p0->data, p1 ? p1->data : nullptr, wt, (const int32_t*)clen->data,
                                  (const int32_t*)seed->data, ev->data, nullptr, sk, vec_len, out_len, chunk, (int)corder,
                                  0, -1, 0, -1, out->data, fr.scratch, fr.scratch_bytes, fr.stream);
    return fr.finish(rc, name);
}

XLA_FFI_Error* dense_common(XLA_FFI_CallFrame* frame, bool matmat, const char* name) {
    BE_XLA_PROLOGUE(2, 1, "dense");
    XLA_FFI_Buffer *w = fr.arg(0), *ev = fr.arg(1), *out = fr.ret(0);
    int64_t transpose = 0;
    fr.attr_i64("transpose", &transpose);
    const int wt = wt_of(w->dtype), sk = spk_of(ev->dtype);
    if (wt < 0 || sk < 0 || w->rank != 2) return fr.fail(BE_B200_EUNIMPL, std::string(name) + ": unsupported dtype/rank");
    const int64_t rows = w->dims[0], cols = w->dims[1], n_src = transpose ? rows : cols;
    const int64_t n = matmat ? out->dims[1] : 1;
    if (XLA_FFI_Error* e = fr.get_scratch(n_src, n)) return e;
    int rc;
    if (matmat)
        rc = be_b200_binary_densemm(w->data, wt, rows, cols, ev->data, nullptr, sk, n, (int)transpose, out->data, fr.scratch,
                                    fr.scratch_bytes, fr.stream);
    else
        rc = be_b200_binary_densemv(w->data, wt, rows, cols, ev->data, nullptr, sk, (int)transpose, out->data, fr.scratch,
                                    fr.scratch_bytes, fr.stream);
    return fr.finish(rc, name);
}

// ---- float (non-event) products: SURVEY 8(f)-1 -------------------------------------------------------
XLA_FFI_Error* csr_val_common(XLA_FFI_CallFrame* frame, bool matmat, const char* name) {
    BE_XLA_PROLOGUE(4, 1, "csr");
    XLA_FFI_Buffer *w = fr.arg(0), *idx = fr.arg(1), *ptr = fr.arg(2), *x = fr.arg(3), *out = fr.ret(0);
    int64_t transpose = 0;
    fr.attr_i64("transpose", &transpose);
    const int wt = wt_of(w->dtype);
    if (wt < 0 || x->dtype != w->dtype) return fr.fail(BE_B200_EUNIMPL, std::string(name) + ": operand and weights must share a float dtype");
    if (idx->dtype != BE_XLA_S32) return fr.fail(BE_B200_EINVAL, std::string(name) + ": indices must be int32");
    const int64_t m = ptr->dims[0] - 1, nse = numel(idx);
    const int64_t k = transpose ? out->dims[0] : x->dims[0];
    int rc;
    if (matmat)
        rc = be_b200_csrmm(w->data, wt, numel(w) == 1, (const int32_t*)idx->data, ptr->data, ptr->dtype == BE_XLA_S64, x->data,
                           m, k, nse, x->dims[1], (int)transpose, out->data, fr.stream);
    else
        rc = be_b200_csrmv(w->data, wt, numel(w) == 1, (const int32_t*)idx->data, ptr->data, ptr->dtype == BE_XLA_S64, x->data,
                           m, k, nse, (int)transpose, out->data, fr.stream);
    return fr.finish(rc, name);
}

XLA_FFI_Error* fcn_val_common(XLA_FFI_CallFrame* frame, bool matmat, const char* name) {
    BE_XLA_PROLOGUE(3, 1, "fcn");
    XLA_FFI_Buffer *w = fr.arg(0), *idx = fr.arg(1), *x = fr.arg(2), *out = fr.ret(0);
    int64_t transpose = 0;
    fr.attr_i64("transpose", &transpose);
    const int wt = wt_of(w->dtype);
    if (wt < 0 || x->dtype != w->dtype) return fr.fail(BE_B200_EUNIMPL, std::string(name) + ": operand and weights must share a float dtype");
    if (idx->dtype != BE_XLA_S32 || idx->rank != 2) return fr.fail(BE_B200_EINVAL, std::string(name) + ": indices must be int32 [n_rows, n_conn]");
    const int64_t n_rows = idx->dims[0], n_conn = idx->dims[1];
    const int64_t n_other = transpose ? out->dims[0] : x->dims[0];
    int rc;
    if (matmat)
        rc = be_b200_fcnmm(w->data, wt, numel(w) == 1, (const int32_t*)idx->data, n_rows, n_conn, x->data, n_other, x->dims[1],
                           (int)transpose, out->data, fr.stream);
    else
        rc = be_b200_fcnmv(w->data, wt, numel(w) == 1, (const int32_t*)idx->data, n_rows, n_conn, x->data, n_other,
                           (int)transpose, out->data, fr.stream);
    return fr.finish(rc, name);
}

XLA_FFI_Error* jit_val_common(XLA_FFI_CallFrame* frame, bool matmat, const char* name) {
    Frame fr(frame);
    if (fr.metadata()) return nullptr;
    if (frame->stage != 3) return nullptr;
    int64_t kind = 0, corder = 1, chunk = 0;
    fr.attr_i64("kind", &kind);
    fr.attr_i64("corder", &corder);
    if (!fr.attr_i64("chunk_size", &chunk)) return fr.fail(BE_B200_EINVAL, std::string(name) + ": attr chunk_size is required");
    const int np = kind == BE_B200_JIT_SCALAR ? 1 : 2;
    if (frame->args.size != np + 3 || frame->rets.size < 1) return fr.fail(BE_B200_EINVAL, std::string(name) + ": wrong number of buffers");
    if (XLA_FFI_Error* e = fr.get_stream()) return e;
    XLA_FFI_Buffer *p0 = fr.arg(0), *p1 = np == 2 ? fr.arg(1) : nullptr, *clen = fr.arg(np), *seed = fr.arg(np + 1),
                   *x = fr.arg(np + 2), *out = fr.ret(0);
    const int wt = wt_of(p0->dtype);
    if (wt < 0 || x->dtype != p0->dtype) return fr.fail(BE_B200_EUNIMPL, std::string(name) + ": operand and weights must share a float dtype");
    int rc;
    if (matmat)
        rc = be_b200_jitmm((int)kind, p0->data, p1 ? p1->data : nullptr, wt, (const int32_t*)clen->data, (const int32_t*)seed->data,
                           x->data, x->dims[0], out->dims[0], x->dims[1], chunk, (int)corder, out->data, fr.stream);
    else
        rc = be_b200_jitmv((int)kind, p0->data, p1 ? p1->data : nullptr, wt, (const int32_t*)clen->data, (const int32_t*)seed->data,
                           x->data, x->dims[0], out->dims[0], chunk, (int)corder, out->data, fr.stream);
    return fr.finish(rc, name);
}


// ==========================================================================================================
// Reference-name-compatible decoders (SURVEY 8(b) last row / Appendix D): same argument order, result order,
// aliasing and attributes as the reference's `// @BE` entry points, one generic decoder per family.  The
// per-name symbols are generated into xla_ref_aliases.inc (tools/gen_ffi_aliases.py).
// ==========================================================================================================
size_t xla_dtype_size(int dt) {
    switch (dt) {
        case BE_XLA_PRED: case BE_XLA_S8: case BE_XLA_U8: return 1;
        case BE_XLA_F16: case BE_XLA_BF16: return 2;
        case BE_XLA_S32: case BE_XLA_U32: case BE_XLA_F32: return 4;
        case BE_XLA_S64: case BE_XLA_F64: return 8;
        default: return 0;
    }
}

// out[c, r] = in[r, c]   (the reference's dense / JITC mm kernels write a physical batch-major buffer)
template <typename T>
__global__ void transpose_2d_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t rows, int64_t cols) {
    __shared__ T tile[32][33];
    const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int64_t r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = in[r * cols + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int64_t c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[c * rows + r] = tile[threadIdx.x][i];
    }
}
int launch_transpose(const void* in, void* out, int64_t rows, int64_t cols, size_t elem, cudaStream_t s) {
    if (rows == 0 || cols == 0) return BE_B200_OK;
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32)), block(32, 8);
    if (elem == 4) transpose_2d_kernel<uint32_t><<<grid, block, 0, s>>>((const uint32_t*)in, (uint32_t*)out, rows, cols);
    else if (elem == 8) transpose_2d_kernel<uint64_t><<<grid, block, 0, s>>>((const uint64_t*)in, (uint64_t*)out, rows, cols);
    else if (elem == 2) transpose_2d_kernel<uint16_t><<<grid, block, 0, s>>>((const uint16_t*)in, (uint16_t*)out, rows, cols);
    else return BE_B200_EUNIMPL;
    return cudaGetLastError() == cudaSuccess ? BE_B200_OK : BE_B200_EINTERNAL;
}

// reference jit*mm `pack`: B int8 [k, n] -> packed[n, n_words], bit b of word w of column c = B[32w + b, c] != 0
// (ref: _jit_scalar/binary_jitsmm.cu:130-154)
__global__ void pack_cols_kernel(const int8_t* __restrict__ B, uint32_t* __restrict__ packed, int64_t k, int64_t n,
                                 int64_t n_words) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * n_words) return;
    const int64_t c = i / n_words, w = i - c * n_words;
    uint32_t bits = 0;
    for (int b = 0; b < 32; ++b) {
        const int64_t r = w * 32 + b;
        if (r < k && B[r * n + c] != 0) bits |= 1u << b;
    }
    packed[i] = bits;
}
// column-major packing -> this library's row-major packing [k, ceil(n/32)] (bit c of word cw of row r)
__global__ void repack_rows_kernel(const uint32_t* __restrict__ cols, uint32_t* __restrict__ rows_out, int64_t k, int64_t n,
                                   int64_t n_words) {
    const int64_t wpr = (n + 31) / 32;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k * wpr) return;
    const int64_t r = i / wpr, cw = i - r * wpr;
    uint32_t bits = 0;
    for (int b = 0; b < 32; ++b) {
        const int64_t c = cw * 32 + b;
        if (c < n && ((cols[c * n_words + (r >> 5)] >> (r & 31)) & 1u)) bits |= 1u << b;
    }
    rows_out[i] = bits;
}

XLA_FFI_Error* ref_csr(XLA_FFI_CallFrame* frame, bool mm, bool hybrid, bool has_perm) {
    const int na = 4 + (has_perm ? 1 : 0) + (hybrid ? 3 : 0);
    BE_XLA_PROLOGUE(na, hybrid ? 4 : 1, "csr");
    XLA_FFI_Buffer *w = fr.arg(0), *idx = fr.arg(1), *ptr = fr.arg(2), *perm = has_perm ? fr.arg(3) : nullptr,
                   *ev = fr.arg(3 + (has_perm ? 1 : 0)), *out = fr.ret(0);
    const int wt = wt_of(w->dtype), sk = spk_of(ev->dtype);
    if (wt < 0 || sk < 0 || sk == BE_B200_SPK_PACKED) return fr.fail(BE_B200_EUNIMPL, "csr: unsupported weight / event dtype");
    if (idx->dtype != BE_XLA_S32) return fr.fail(BE_B200_EINVAL, "csr: indices must be int32");
    const int64_t m = ptr->dims[0] - 1, nse = numel(idx);
    int64_t k, n = 0;
    if (hybrid) {                       // scatter over active rows: events [m] or [m, n]
        if (mm) { n = out->dims[0]; k = out->dims[1]; } else k = out->dims[0];
        // the reference's task_begin / task_end / status scratch is aliased in -> out; this library needs none of it
        // (the active-row list replaces the task list), so it is passed through untouched
        for (int i = 0; i < 3; ++i) {
            XLA_FFI_Buffer *a = fr.arg(na - 3 + i), *r = fr.ret(1 + i);
            const size_t bytes = (size_t)numel(a) * xla_dtype_size(a->dtype);
            if (r && a->data != r->data && bytes &&
                cudaMemcpyAsync(r->data, a->data, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)fr.stream) != cudaSuccess)
                return fr.fail(BE_B200_EINTERNAL, "csr: workspace pass-through copy failed");
        }
    } else {                            // gather over all rows: events [k] or [k, n]
        k = ev->dims[0];
        if (mm) n = ev->dims[1];
    }
    const int transpose = hybrid ? 1 : 0;
    if (XLA_FFI_Error* e = fr.get_scratch(transpose ? m : k, mm ? n : 1)) return e;
    int rc;
    if (mm)
        rc = be_b200_binary_csrmm(w->data, wt, numel(w) == 1, (const int32_t*)idx->data, ptr->data, ptr->dtype == BE_XLA_S64,
                                  perm ? perm->data : nullptr, perm && perm->dtype == BE_XLA_S64, ev->data, nullptr, sk, m, k, nse,
                                  n, transpose, /*out_batch_major=*/transpose, out->data, fr.scratch, fr.scratch_bytes, fr.stream);
    else
        rc = be_b200_binary_csrmv(w->data, wt, numel(w) == 1, (const int32_t*)idx->data, ptr->data, ptr->dtype == BE_XLA_S64,
                                  perm ? perm->data : nullptr, perm && perm->dtype == BE_XLA_S64, ev->data, nullptr, sk, m, k, nse,
                                  transpose, out->data, fr.scratch, fr.scratch_bytes, fr.stream);
    return fr.finish(rc, "csr");
}

XLA_FFI_Error* ref_fcnmv_scatter(XLA_FFI_CallFrame* frame) {
    BE_XLA_PROLOGUE(3, 1, "fcnmv_scatter");
    XLA_FFI_Buffer *w = fr.arg(0), *idx = fr.arg(1), *ev = fr.arg(2), *out = fr.ret(0);
    const int wt = wt_of(w->dtype), sk = spk_of(ev->dtype);
    if (wt < 0 || sk < 0 || idx->dtype != BE_XLA_S32 || idx->rank != 2) return fr.fail(BE_B200_EUNIMPL, "fcnmv_scatter: unsupported operands");
    const int64_t n_rows = idx->dims[0], n_conn = idx->dims[1];
    if (XLA_FFI_Error* e = fr.get_scratch(n_rows, 1)) return e;
    return fr.finish(be_b200_binary_fcnmv(w->data, wt, numel(w) == 1, (const int32_t*)idx->data, n_rows, n_conn, ev->data, nullptr, sk,
                                          n_rows, out->dims[0], 1, out->data, fr.scratch, fr.scratch_bytes, fr.stream), "fcnmv_scatter");
}

// scatter ("sraw"): matrix [n_pre, n_batch] -> out [n_batch, n_post] (batch first, ref: _fcn/binary.py:1123-1124)
// gather: matrix [n_post, n_batch] or packed uint32 [n_post, ceil(n_batch/32)] -> out [n_pre, n_batch]
XLA_FFI_Error* ref_fcnmm(XLA_FFI_CallFrame* frame, bool scatter, bool packed) {
    BE_XLA_PROLOGUE(3, 1, "fcnmm");
    XLA_FFI_Buffer *w = fr.arg(0), *idx = fr.arg(1), *ev = fr.arg(2), *out = fr.ret(0);
    const int wt = wt_of(w->dtype), sk = packed ? BE_B200_SPK_PACKED : spk_of(ev->dtype);
    if (wt < 0 || sk < 0 || idx->dtype != BE_XLA_S32 || idx->rank != 2 || ev->rank != 2 || out->rank != 2)
        return fr.fail(BE_B200_EUNIMPL, "fcnmm: unsupported operands");
    const int64_t n_rows = idx->dims[0], n_conn = idx->dims[1];
    const int64_t n = scatter ? out->dims[0] : out->dims[1];
    const int64_t n_out = scatter ? out->dims[1] : n_rows, n_src = scatter ? n_rows : ev->dims[0];
    if (XLA_FFI_Error* e = fr.get_scratch(n_src, n)) return e;
    return fr.finish(be_b200_binary_fcnmm(w->data, wt, numel(w) == 1, (const int32_t*)idx->data, n_rows, n_conn, ev->data, nullptr, sk,
                                          n_src, n_out, n, scatter ? 1 : 0, scatter ? 1 : 0, out->data, fr.scratch, fr.scratch_bytes,
                                          fr.stream), "fcnmm");
}

XLA_FFI_Error* ref_fcnmm_pack(XLA_FFI_CallFrame* frame) {
    BE_XLA_PROLOGUE(1, 1, "fcnmm_pack");
    XLA_FFI_Buffer *ev = fr.arg(0), *packed = fr.ret(0);
    const int sk = spk_of(ev->dtype);
    if (sk < 0 || sk == BE_B200_SPK_PACKED || ev->rank != 2) return fr.fail(BE_B200_EUNIMPL, "fcnmm_pack: spikes must be bool / int8 / float32 [rows, batch]");
    const int64_t n = ev->dims[0], b = ev->dims[1];
    if (XLA_FFI_Error* e = fr.get_scratch(n, b)) return e;
    void* ids = nullptr;
    const size_t ids_bytes = (size_t)(n + 1) * 4 + 16;
    if (XLA_FFI_Error* e = fr.alloc_bytes(ids_bytes, &ids)) { fr.release(); return e; }
    const int rc = be_b200_compact_2d(ev->data, sk, n, b, /*nonzero_pred=*/0, (uint32_t*)packed->data, (int32_t*)ids, (int32_t*)ids + n,
                                      fr.scratch, fr.scratch_bytes, fr.stream);
    fr.free_bytes(ids_bytes, ids);
    return fr.finish(rc, "fcnmm_pack");
}

XLA_FFI_Error* ref_jit_pack_bool(XLA_FFI_CallFrame* frame) {
    BE_XLA_PROLOGUE(1, 1, "pack_bool");
    XLA_FFI_Buffer *ev = fr.arg(0), *packed = fr.ret(0);
    if (spk_of(ev->dtype) != BE_B200_SPK_BOOL) return fr.fail(BE_B200_EUNIMPL, "pack_bool: spikes must be int8 / bool");
    return fr.finish(be_b200_bitpack_1d(ev->data, BE_B200_SPK_BOOL, numel(ev), 1, (uint32_t*)packed->data, fr.stream), "pack_bool");
}

// args: scalar (weight, clen, seed, packed); normal (w_loc, w_scale, clen, seed, packed); uniform (w_low, w_high, ...)
// attrs vector_size, chunk_size (int32)   ref: _jit_scalar/binary.py:452-457
XLA_FFI_Error* ref_jitmv(XLA_FFI_CallFrame* frame, int kind, int corder) {
    const int np = kind == BE_B200_JIT_SCALAR ? 1 : 2;
    BE_XLA_PROLOGUE(np + 3, 1, "jitmv");
    XLA_FFI_Buffer *p0 = fr.arg(0), *p1 = np == 2 ? fr.arg(1) : nullptr, *clen = fr.arg(np), *seed = fr.arg(np + 1),
                   *packed = fr.arg(np + 2), *out = fr.ret(0);
    int64_t vec_len = 0, chunk = 0;
    if (!fr.attr_i64("vector_size", &vec_len) || !fr.attr_i64("chunk_size", &chunk))
        return fr.fail(BE_B200_EINVAL, "jitmv: attrs vector_size and chunk_size are required");
    const int wt = wt_of(p0->dtype);
    if (wt < 0 || packed->dtype != BE_XLA_U32) return fr.fail(BE_B200_EUNIMPL, "jitmv: unsupported operands");
    if (XLA_FFI_Error* e = fr.get_scratch(vec_len, 1)) return e;
    return fr.finish(be_b200_binary_jitmv(kind, p0->data, p1 ? p1->data : nullptr, wt, (const int32_t*)clen->data, (const int32_t*)seed->data,
                                          packed->data, nullptr, BE_B200_SPK_PACKED, vec_len, out->dims[0], chunk, corder, 0, -1, 0, -1,
                                          out->data, fr.scratch, fr.scratch_bytes, fr.stream), "jitmv");
}

XLA_FFI_Error* ref_jit_pack_mm(XLA_FFI_CallFrame* frame) {
    BE_XLA_PROLOGUE(1, 1, "pack");
    XLA_FFI_Buffer *B = fr.arg(0), *packed = fr.ret(0);
    int64_t k = 0, n = 0, n_words = 0;
    if (!fr.attr_i64("k", &k) || !fr.attr_i64("n", &n) || !fr.attr_i64("n_words", &n_words))
        return fr.fail(BE_B200_EINVAL, "pack: attrs k, n, n_words are required");
    if (spk_of(B->dtype) != BE_B200_SPK_BOOL) return fr.fail(BE_B200_EUNIMPL, "pack: spikes must be int8 / bool");
    const int64_t total = n * n_words;
    if (total > 0)
        pack_cols_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)fr.stream>>>((const int8_t*)B->data, (uint32_t*)packed->data,
                                                                                             k, n, n_words);
    return fr.finish(cudaGetLastError() == cudaSuccess ? BE_B200_OK : BE_B200_EINTERNAL, "pack");
}

// attrs m, k, n, n_words, chunk_size; notrans: out [m, n] walks k = packed rows; trans: events m, out [k, n]
// (ref: _jit_scalar/binary.py:978-997)
XLA_FFI_Error* ref_jitmm(XLA_FFI_CallFrame* frame, int kind, int corder) {
    const int np = kind == BE_B200_JIT_SCALAR ? 1 : 2;
    BE_XLA_PROLOGUE(np + 3, 1, "jitmm");
    XLA_FFI_Buffer *p0 = fr.arg(0), *p1 = np == 2 ? fr.arg(1) : nullptr, *clen = fr.arg(np), *seed = fr.arg(np + 1),
                   *packed = fr.arg(np + 2), *out = fr.ret(0);
    int64_t m = 0, k = 0, n = 0, n_words = 0, chunk = 0;
    if (!fr.attr_i64("m", &m) || !fr.attr_i64("k", &k) || !fr.attr_i64("n", &n) || !fr.attr_i64("n_words", &n_words) ||
        !fr.attr_i64("chunk_size", &chunk))
        return fr.fail(BE_B200_EINVAL, "jitmm: attrs m, k, n, n_words, chunk_size are required");
    const int wt = wt_of(p0->dtype);
    if (wt < 0 || packed->dtype != BE_XLA_U32) return fr.fail(BE_B200_EUNIMPL, "jitmm: unsupported operands");
    const int64_t vec_len = corder ? k : m, out_len = corder ? m : k;
    const int64_t wpr = (n + 31) / 32;
    void* rows = nullptr;
    const size_t rows_bytes = (size_t)std::max<int64_t>(vec_len * wpr, 4) * 4;
    if (XLA_FFI_Error* e = fr.alloc_bytes(rows_bytes, &rows)) return e;
    if (vec_len * wpr > 0)
        repack_rows_kernel<<<(unsigned)((vec_len * wpr + 255) / 256), 256, 0, (cudaStream_t)fr.stream>>>((const uint32_t*)packed->data,
                                                                                                        (uint32_t*)rows, vec_len, n, n_words);
    if (XLA_FFI_Error* e = fr.get_scratch(vec_len, n)) { fr.free_bytes(rows_bytes, rows); return e; }
    const int rc = be_b200_binary_jitmm(kind, p0->data, p1 ? p1->data : nullptr, wt, (const int32_t*)clen->data, (const int32_t*)seed->data,
                                        rows, nullptr, BE_B200_SPK_PACKED, vec_len, out_len, n, chunk, corder, out->data, fr.scratch,
                                        fr.scratch_bytes, fr.stream);
    fr.free_bytes(rows_bytes, rows);
    return fr.finish(rc, "jitmm");
}

// mv: (weights [rows, cols], spikes) -> out.  mm: spikes [k, n] -> PHYSICAL out [n, m] (batch-major; the reference's
// Python transposes it back, ref: _dense/binary.py:980-988)
XLA_FFI_Error* ref_dense(XLA_FFI_CallFrame* frame, bool mm, int transpose) {
    BE_XLA_PROLOGUE(2, 1, "dense");
    XLA_FFI_Buffer *w = fr.arg(0), *ev = fr.arg(1), *out = fr.ret(0);
    const int wt = wt_of(w->dtype), sk = spk_of(ev->dtype);
    if (wt < 0 || sk < 0 || sk == BE_B200_SPK_PACKED || w->rank != 2) return fr.fail(BE_B200_EUNIMPL, "dense: unsupported operands");
    const int64_t rows = w->dims[0], cols = w->dims[1], n_src = transpose ? rows : cols, m = transpose ? cols : rows;
    if (!mm) {
        if (XLA_FFI_Error* e = fr.get_scratch(n_src, 1)) return e;
        return fr.finish(be_b200_binary_densemv(w->data, wt, rows, cols, ev->data, nullptr, sk, transpose, out->data, fr.scratch,
                                                fr.scratch_bytes, fr.stream), "densemv");
    }
    const int64_t n = ev->dims[1];
    if (XLA_FFI_Error* e = fr.get_scratch(n_src, n)) return e;
    void* tmp = nullptr;
    const size_t elem = xla_dtype_size(w->dtype), tmp_bytes = (size_t)std::max<int64_t>(m * n, 4) * elem;
    if (XLA_FFI_Error* e = fr.alloc_bytes(tmp_bytes, &tmp)) { fr.release(); return e; }
    int rc = be_b200_binary_densemm(w->data, wt, rows, cols, ev->data, nullptr, sk, n, transpose, tmp, fr.scratch, fr.scratch_bytes, fr.stream);
    if (rc == BE_B200_OK) rc = launch_transpose(tmp, out->data, m, n, elem, (cudaStream_t)fr.stream);
    fr.free_bytes(tmp_bytes, tmp);
    return fr.finish(rc, "densemm");
}

}  // namespace

extern "C" {

XLA_FFI_Error* be_b200_xla_binary_csrmv(XLA_FFI_CallFrame* frame) { return csr_common(frame, false, false, "be_b200_xla_binary_csrmv"); }
XLA_FFI_Error* be_b200_xla_binary_csrmm(XLA_FFI_CallFrame* frame) { return csr_common(frame, true, false, "be_b200_xla_binary_csrmm"); }
XLA_FFI_Error* be_b200_xla_binary_csrmv_indexed(XLA_FFI_CallFrame* frame) { return csr_common(frame, false, true, "be_b200_xla_binary_csrmv_indexed"); }
XLA_FFI_Error* be_b200_xla_binary_csrmm_indexed(XLA_FFI_CallFrame* frame) { return csr_common(frame, true, true, "be_b200_xla_binary_csrmm_indexed"); }
XLA_FFI_Error* be_b200_xla_binary_fcnmv(XLA_FFI_CallFrame* frame) { return fcn_common(frame, false, "be_b200_xla_binary_fcnmv"); }
XLA_FFI_Error* be_b200_xla_binary_fcnmm(XLA_FFI_CallFrame* frame) { return fcn_common(frame, true, "be_b200_xla_binary_fcnmm"); }
XLA_FFI_Error* be_b200_xla_binary_jitmv(XLA_FFI_CallFrame* frame) { return jit_common(frame, false, "be_b200_xla_binary_jitmv"); }
XLA_FFI_Error* be_b200_xla_binary_jitmm(XLA_FFI_CallFrame* frame) { return jit_common(frame, true, "be_b200_xla_binary_jitmm"); }
XLA_FFI_Error* be_b200_xla_binary_densemv(XLA_FFI_CallFrame* frame) { return dense_common(frame, false, "be_b200_xla_binary_densemv"); }
XLA_FFI_Error* be_b200_xla_binary_densemm(XLA_FFI_CallFrame* frame) { return dense_common(frame, true, "be_b200_xla_binary_densemm"); }

XLA_FFI_Error* be_b200_xla_compact_1d(XLA_FFI_CallFrame* frame) {
    BE_XLA_PROLOGUE(1, 2, "compact_1d");
    XLA_FFI_Buffer *ev = fr.arg(0), *ids = fr.ret(0), *na = fr.ret(1);
    const int sk = spk_of(ev->dtype);
    if (sk < 0) return fr.fail(BE_B200_EUNIMPL, "compact_1d: unsupported spike dtype");
    const int64_t n = ids->dims[0];
    if (XLA_FFI_Error* e = fr.get_scratch(n, 1)) return e;
    return fr.finish(be_b200_compact_1d(ev->data, sk, n, 1, (int32_t*)ids->data, (int32_t*)na->data, nullptr, fr.scratch,
                                        fr.scratch_bytes, fr.stream), "be_b200_xla_compact_1d");
}

XLA_FFI_Error* be_b200_xla_compact_2d(XLA_FFI_CallFrame* frame) {
    BE_XLA_PROLOGUE(1, 3, "compact_2d");
    XLA_FFI_Buffer *ev = fr.arg(0), *packed = fr.ret(0), *ids = fr.ret(1), *na = fr.ret(2);
    const int sk = spk_of(ev->dtype);
    if (sk < 0 || ev->rank != 2) return fr.fail(BE_B200_EUNIMPL, "compact_2d: unsupported spike dtype/rank");
    const int64_t n = ev->dims[0], b = ev->dims[1];
    if (XLA_FFI_Error* e = fr.get_scratch(n, b)) return e;
    return fr.finish(be_b200_compact_2d(ev->data, sk, n, b, 1, (uint32_t*)packed->data, (int32_t*)ids->data, (int32_t*)na->data,
                                        fr.scratch, fr.scratch_bytes, fr.stream), "be_b200_xla_compact_2d");
}

XLA_FFI_Error* be_b200_xla_compact_2d_only(XLA_FFI_CallFrame* frame) {
    BE_XLA_PROLOGUE(1, 2, "compact_2d_only");
    XLA_FFI_Buffer *ev = fr.arg(0), *ids = fr.ret(0), *na = fr.ret(1);
    const int sk = spk_of(ev->dtype);
    if (sk < 0 || ev->rank != 2) return fr.fail(BE_B200_EUNIMPL, "compact_2d_only: unsupported spike dtype/rank");
    const int64_t n = ev->dims[0], b = ev->dims[1];
    if (XLA_FFI_Error* e = fr.get_scratch(n, b)) return e;
    return fr.finish(be_b200_compact_2d(ev->data, sk, n, b, 1, nullptr, (int32_t*)ids->data, (int32_t*)na->data, fr.scratch,
                                        fr.scratch_bytes, fr.stream), "be_b200_xla_compact_2d_only");
}

XLA_FFI_Error* be_b200_xla_bitpack_1d(XLA_FFI_CallFrame* frame) {
    BE_XLA_PROLOGUE(1, 1, "bitpack_1d");
    XLA_FFI_Buffer *ev = fr.arg(0), *packed = fr.ret(0);
    const int sk = spk_of(ev->dtype);
    if (sk < 0) return fr.fail(BE_B200_EUNIMPL, "bitpack_1d: unsupported spike dtype");
    return fr.finish(be_b200_bitpack_1d(ev->data, sk, ev->dims[0], 0, (uint32_t*)packed->data, fr.stream), "be_b200_xla_bitpack_1d");
}

// ---- SURVEY 8(f) rows -------------------------------------------------------------------------------
XLA_FFI_Error* be_b200_xla_csrmv(XLA_FFI_CallFrame* frame) { return csr_val_common(frame, false, "be_b200_xla_csrmv"); }
XLA_FFI_Error* be_b200_xla_csrmm(XLA_FFI_CallFrame* frame) { return csr_val_common(frame, true, "be_b200_xla_csrmm"); }
XLA_FFI_Error* be_b200_xla_fcnmv(XLA_FFI_CallFrame* frame) { return fcn_val_common(frame, false, "be_b200_xla_fcnmv"); }
XLA_FFI_Error* be_b200_xla_fcnmm(XLA_FFI_CallFrame* frame) { return fcn_val_common(frame, true, "be_b200_xla_fcnmm"); }
XLA_FFI_Error* be_b200_xla_jitmv(XLA_FFI_CallFrame* frame) { return jit_val_common(frame, false, "be_b200_xla_jitmv"); }
XLA_FFI_Error* be_b200_xla_jitmm(XLA_FFI_CallFrame* frame) { return jit_val_common(frame, true, "be_b200_xla_jitmm"); }

// args: weights, indices, [indptr if has_indptr], [perm if has_perm], spikes, trace; ret: updated weights
// (the reference aliases it onto arg 0, input_output_aliases={0: 0}; copied first when XLA did not).
XLA_FFI_Error* be_b200_xla_plasticity_rows(XLA_FFI_CallFrame* frame) {
    Frame fr(frame);
    if (fr.metadata()) return nullptr;
    if (frame->stage != 3) return nullptr;
    int64_t has_indptr = 1, has_perm = 0;
    fr.attr_i64("has_indptr", &has_indptr);
    fr.attr_i64("has_perm", &has_perm);
    double lo = -INFINITY, hi = INFINITY;
    fr.attr_f64("w_min", &lo);
    fr.attr_f64("w_max", &hi);
    const int na = 4 + (has_indptr ? 1 : 0) + (has_perm ? 1 : 0);
    if (frame->args.size != na || frame->rets.size < 1) return fr.fail(BE_B200_EINVAL, "plasticity_rows: wrong number of buffers");
    if (XLA_FFI_Error* e = fr.get_stream()) return e;
    int a = 0;
    XLA_FFI_Buffer *w = fr.arg(a++), *idx = fr.arg(a++), *ptr = has_indptr ? fr.arg(a++) : nullptr,
                   *perm = has_perm ? fr.arg(a++) : nullptr, *spk = fr.arg(a++), *trace = fr.arg(a++), *out = fr.ret(0);
    const int wt = wt_of(w->dtype), sk = spk_of(spk->dtype);
    if (wt < 0 || sk < 0 || trace->dtype != w->dtype) return fr.fail(BE_B200_EUNIMPL, "plasticity_rows: unsupported dtype");
    if (idx->dtype != BE_XLA_S32) return fr.fail(BE_B200_EINVAL, "plasticity_rows: indices must be int32");
    if (!has_indptr && idx->rank != 2) return fr.fail(BE_B200_EINVAL, "plasticity_rows: ELL indices must be [n_rows, n_conn]");
    static const size_t kWt[4] = {4, 8, 2, 2};
    if (out->data != w->data &&
        cudaMemcpyAsync(out->data, w->data, (size_t)numel(w) * kWt[wt], cudaMemcpyDeviceToDevice, (cudaStream_t)fr.stream) != cudaSuccess)
        return fr.fail(BE_B200_EINTERNAL, "plasticity_rows: weight copy failed");
    const int64_t n_rows = has_indptr ? ptr->dims[0] - 1 : idx->dims[0];
    const int64_t n_conn = has_indptr ? -1 : idx->dims[1];
    if (XLA_FFI_Error* e = fr.get_scratch(n_rows, 1)) return e;
    int rc = be_b200_plasticity_rows(out->data, wt, (const int32_t*)idx->data, ptr ? ptr->data : nullptr,
                                     ptr && ptr->dtype == BE_XLA_S64, n_conn, perm ? perm->data : nullptr,
                                     perm && perm->dtype == BE_XLA_S64, spk->data, nullptr, sk, trace->data, n_rows,
                                     trace->dims[0], numel(idx), lo, hi, fr.scratch, fr.scratch_bytes, fr.stream);
    if (rc == BE_B200_OK && (lo != -INFINITY || hi != INFINITY))
        rc = be_b200_clip(out->data, wt, numel(w), lo, hi, fr.stream);     // the reference clips the whole array
    return fr.finish(rc, "be_b200_xla_plasticity_rows");
}

// arg: spikes [n_src, n_batch]; rets per encoder as in brainevent_b200.h
XLA_FFI_Error* be_b200_xla_binary_2d_csr_encode(XLA_FFI_CallFrame* frame) {
    BE_XLA_PROLOGUE(1, 2, "csr_encode");
    XLA_FFI_Buffer *ev = fr.arg(0), *indices = fr.ret(0), *indptr = fr.ret(1);
    const int sk = spk_of(ev->dtype);
    if (sk < 0 || sk == BE_B200_SPK_PACKED || ev->rank != 2) return fr.fail(BE_B200_EUNIMPL, "csr_encode: unsupported spike dtype/rank");
    const int64_t n = ev->dims[0], b = ev->dims[1];
    if (XLA_FFI_Error* e = fr.get_scratch(n * b, 1)) return e;
    return fr.finish(be_b200_binary_2d_csr_encode(ev->data, sk, n, b, (int32_t*)indices->data, (int32_t*)indptr->data, fr.scratch,
                                                  fr.scratch_bytes, fr.stream), "be_b200_xla_binary_2d_csr_encode");
}

XLA_FFI_Error* be_b200_xla_binary_2d_pair_stream_encode(XLA_FFI_CallFrame* frame) {
    BE_XLA_PROLOGUE(1, 2, "pair_stream_encode");
    XLA_FFI_Buffer *ev = fr.arg(0), *pairs = fr.ret(0), *np_ = fr.ret(1);
    const int sk = spk_of(ev->dtype);
    if (sk < 0 || sk == BE_B200_SPK_PACKED || ev->rank != 2) return fr.fail(BE_B200_EUNIMPL, "pair_stream_encode: unsupported spike dtype/rank");
    const int64_t n = ev->dims[0], b = ev->dims[1];
    if (XLA_FFI_Error* e = fr.get_scratch(n * b, 1)) return e;
    return fr.finish(be_b200_binary_2d_pair_stream_encode(ev->data, sk, n, b, (int32_t*)pairs->data, (int32_t*)np_->data, fr.scratch,
                                                          fr.scratch_bytes, fr.stream), "be_b200_xla_binary_2d_pair_stream_encode");
}

XLA_FFI_Error* be_b200_xla_binary_2d_csc_encode(XLA_FFI_CallFrame* frame) {
    BE_XLA_PROLOGUE(1, 2, "csc_encode");
    XLA_FFI_Buffer *ev = fr.arg(0), *indices = fr.ret(0), *indptr = fr.ret(1);
    const int sk = spk_of(ev->dtype);
    if (sk < 0 || sk == BE_B200_SPK_PACKED || ev->rank != 2) return fr.fail(BE_B200_EUNIMPL, "csc_encode: unsupported spike dtype/rank");
    const int64_t n = ev->dims[0], b = ev->dims[1];
    if (XLA_FFI_Error* e = fr.get_scratch(n * b, 1)) return e;
    void* mask = nullptr;
    const size_t mask_bytes = (size_t)std::max<int64_t>(n * b, 16);
    if (XLA_FFI_Error* e = fr.alloc_bytes(mask_bytes, &mask)) { fr.release(); return e; }
    const int rc = be_b200_binary_2d_csc_encode(ev->data, sk, n, b, (int32_t*)indices->data, (int32_t*)indptr->data, mask,
                                                fr.scratch, fr.scratch_bytes, fr.stream);
    fr.free_bytes(mask_bytes, mask);
    return fr.finish(rc, "be_b200_xla_binary_2d_csc_encode");
}

// rets: spike_indices [n_src, row_size], max_row_nnz [1]
XLA_FFI_Error* be_b200_xla_binary_2d_row_sparse_encode(XLA_FFI_CallFrame* frame) {
    BE_XLA_PROLOGUE(1, 2, "row_sparse_encode");
    XLA_FFI_Buffer *ev = fr.arg(0), *out = fr.ret(0), *mx = fr.ret(1);
    const int sk = spk_of(ev->dtype);
    if (sk < 0 || sk == BE_B200_SPK_PACKED || ev->rank != 2 || out->rank != 2) return fr.fail(BE_B200_EUNIMPL, "row_sparse_encode: unsupported spike dtype/rank");
    return fr.finish(be_b200_binary_2d_row_sparse_encode(ev->data, sk, ev->dims[0], ev->dims[1], out->dims[1], (int32_t*)out->data,
                                                         (int32_t*)mx->data, fr.stream), "be_b200_xla_binary_2d_row_sparse_encode");
}

#include "xla_ref_aliases.inc"

}  // extern "C"
