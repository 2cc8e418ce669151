/*
 * brainevent_b200_xla.h -- XLA-FFI (jax.ffi) entry points of libbrainevent_b200.so.
 *
 * The reference reaches its CUDA kernels through `jax.ffi.ffi_call("<module>.<fn>", ...)` whose
 * target is an `extern "C" XLA_FFI_Error* be_<fn>(XLA_FFI_CallFrame*)` symbol resolved with
 * ctypes and registered by `jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(sym),
 * platform="CUDA")`  (ref: _op/kernix_runtime.py:161-189,308-314; generated handler
 * _op/kernix_codegen.py:617-736).  The symbols below are bound the same way; INTEGRATION.md shows
 * the kernel generators a maintainer registers with `XLACustomKernel.def_kernel('b200', 'gpu', ...)`.
 *
 * jaxlib's `xla/ffi/api/c_api.h` is not available in the build image, so the (stable, C) call-frame
 * ABI is restated here from the reference's own ctypes description of it
 * (ref: _op/numba_ffi.py:172-395).  Handlers are header-less: they decode the
 * frame, fetch the stream with `XLA_FFI_Stream_Get`, take scratch from
 * `XLA_FFI_DeviceMemory_Allocate`, enqueue on XLA's stream and never synchronise.  They answer the
 * metadata handshake (extension type 1) with the caller's API version and the
 * COMMAND_BUFFER_COMPATIBLE trait (ref: _op/numba_ffi.py:812-834).
 *
 * Buffers: dense row-major, dtype taken from the buffer (weights f32/f64/f16/bf16; events
 * PRED/S8/U8 -> bool, F32 -> "> 0", U32 -> bit-packed words).  Scalar attributes are int32/int64.
 */
#ifndef BRAINEVENT_B200_XLA_H_
#define BRAINEVENT_B200_XLA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- restated XLA FFI C ABI -------------------------------------------------------------- */
typedef struct XLA_FFI_Error XLA_FFI_Error;
typedef struct XLA_FFI_ExecutionContext XLA_FFI_ExecutionContext;

typedef struct XLA_FFI_Extension_Base {
    size_t struct_size;
    int type;   /* 1 = Metadata */
    struct XLA_FFI_Extension_Base* next;
} XLA_FFI_Extension_Base;

typedef struct XLA_FFI_Api_Version {
    size_t struct_size;
    XLA_FFI_Extension_Base* extension_start;
    int major_version;
    int minor_version;
} XLA_FFI_Api_Version;

typedef struct XLA_FFI_Metadata {
    size_t struct_size;
    XLA_FFI_Api_Version api_version;
    uint32_t traits;           /* bit 0: COMMAND_BUFFER_COMPATIBLE */
    int64_t state_type_id;     /* XLA_FFI_TypeId */
} XLA_FFI_Metadata;

typedef struct XLA_FFI_Metadata_Extension {
    XLA_FFI_Extension_Base extension_base;
    XLA_FFI_Metadata* metadata;
} XLA_FFI_Metadata_Extension;

typedef struct XLA_FFI_Buffer {
    size_t struct_size;
    XLA_FFI_Extension_Base* extension_start;
    int dtype;       /* XLA_FFI_DataType */
    void* data;
    int64_t rank;
    int64_t* dims;
} XLA_FFI_Buffer;

typedef struct XLA_FFI_ByteSpan { const char* ptr; size_t len; } XLA_FFI_ByteSpan;
typedef struct XLA_FFI_Scalar { int dtype; void* value; } XLA_FFI_Scalar;

typedef struct XLA_FFI_Args {
    size_t struct_size; XLA_FFI_Extension_Base* extension_start;
    int64_t size; int* types; void** args;
} XLA_FFI_Args;
typedef struct XLA_FFI_Rets {
    size_t struct_size; XLA_FFI_Extension_Base* extension_start;
    int64_t size; int* types; void** rets;
} XLA_FFI_Rets;
typedef struct XLA_FFI_Attrs {
    size_t struct_size; XLA_FFI_Extension_Base* extension_start;
    int64_t size; int* types; XLA_FFI_ByteSpan** names; void** attrs;
} XLA_FFI_Attrs;

struct XLA_FFI_Api;
typedef struct XLA_FFI_CallFrame {
    size_t struct_size;
    XLA_FFI_Extension_Base* extension_start;
    const struct XLA_FFI_Api* api;
    XLA_FFI_ExecutionContext* ctx;
    int stage;       /* 3 = EXECUTE */
    XLA_FFI_Args args;
    XLA_FFI_Rets rets;
    XLA_FFI_Attrs attrs;
    void* future;
} XLA_FFI_CallFrame;

typedef struct XLA_FFI_Error_Create_Args {
    size_t struct_size; XLA_FFI_Extension_Base* extension_start; const char* message; int errc;
} XLA_FFI_Error_Create_Args;
typedef struct XLA_FFI_Stream_Get_Args {
    size_t struct_size; XLA_FFI_Extension_Base* extension_start; XLA_FFI_ExecutionContext* ctx; void* stream;
} XLA_FFI_Stream_Get_Args;
typedef struct XLA_FFI_DeviceMemory_Allocate_Args {
    size_t struct_size; XLA_FFI_Extension_Base* extension_start; XLA_FFI_ExecutionContext* ctx;
    size_t size; size_t alignment; void* data;
} XLA_FFI_DeviceMemory_Allocate_Args;
typedef struct XLA_FFI_DeviceMemory_Free_Args {
    size_t struct_size; XLA_FFI_Extension_Base* extension_start; XLA_FFI_ExecutionContext* ctx;
    size_t size; void* data;
} XLA_FFI_DeviceMemory_Free_Args;

/* Function table; entries not used here are opaque pointers of the same width
 * (order: ref _op/numba_ffi.py:363-395). */
typedef struct XLA_FFI_Api {
    size_t struct_size;
    XLA_FFI_Extension_Base* extension_start;
    XLA_FFI_Api_Version api_version;
    void* internal_api;
    XLA_FFI_Error* (*XLA_FFI_Error_Create)(XLA_FFI_Error_Create_Args*);
    void* XLA_FFI_Error_GetMessage;
    void* XLA_FFI_Error_Destroy;
    void* XLA_FFI_Handler_Register;
    XLA_FFI_Error* (*XLA_FFI_Stream_Get)(XLA_FFI_Stream_Get_Args*);
    void* XLA_FFI_Type_Register;
    void* XLA_FFI_ExecutionContext_Get;
    void* XLA_FFI_State_Set;
    void* XLA_FFI_State_Get;
    XLA_FFI_Error* (*XLA_FFI_DeviceMemory_Allocate)(XLA_FFI_DeviceMemory_Allocate_Args*);
    XLA_FFI_Error* (*XLA_FFI_DeviceMemory_Free)(XLA_FFI_DeviceMemory_Free_Args*);
    void* XLA_FFI_ThreadPool_Schedule;
    void* XLA_FFI_ThreadPool_NumThreads;
    void* XLA_FFI_Future_Create;
    void* XLA_FFI_Future_SetAvailable;
    void* XLA_FFI_Future_SetError;
    void* XLA_FFI_RunId_Get;
    void* XLA_FFI_DeviceOrdinal_Get;
} XLA_FFI_Api;

/* XLA_FFI_DataType values used (ref: include/brainevent/dtypes.h:33-53) */
enum { BE_XLA_PRED = 1, BE_XLA_S8 = 2, BE_XLA_S32 = 4, BE_XLA_S64 = 5, BE_XLA_U8 = 6, BE_XLA_U32 = 8,
       BE_XLA_F16 = 10, BE_XLA_F32 = 11, BE_XLA_F64 = 12, BE_XLA_BF16 = 16 };

/* ---- handlers -------------------------------------------------------------------------------
 * Notation: args -> rets {attrs}.  Every handler replaces the reference FFI family cited. */

/* (weights, indices, indptr, events) -> (out) {transpose:i32}
 * replaces csr_binary_csrmv.binary_csrmv_nt_auto_* (_csr/binary_csrmv.cu:363-486) and
 * csr_binary_csrmv_hybrid_*.binary_csrmv_wat_hybrid_* (_csr/binary_csrmv_hybrid.cu:618-821) */
XLA_FFI_Error* be_b200_xla_binary_csrmv(XLA_FFI_CallFrame* frame);
/* (weights, indices, indptr, B[rows,n]) -> (out[.,n]) {transpose}
 * replaces csr_binary_csrmm.* / csr_binary_csrmm_hybrid.* (_csr/binary_csrmm.cu:336-453,
 * _csr/binary_csrmm_hybrid.cu:463-788); NOTE out is [k, n] for transpose=1, not the reference's
 * physical [n, k] that its Python side transposes back (_csr/binary.py:1263-1286). */
XLA_FFI_Error* be_b200_xla_binary_csrmm(XLA_FFI_CallFrame* frame);
/* (weights, indices, indptr, perm, events) -> (out) {transpose}
 * replaces binary_indexed_csrm{v,m}* (_csr/binary_indexed_csrmv{,_hybrid}.cu) */
XLA_FFI_Error* be_b200_xla_binary_csrmv_indexed(XLA_FFI_CallFrame* frame);
XLA_FFI_Error* be_b200_xla_binary_csrmm_indexed(XLA_FFI_CallFrame* frame);
/* (weights, indices[n_rows,n_conn], events) -> (out) {transpose}
 * replaces fcn_binary_mv.binary_fcnmv_scatter_* (_fcn/binary_fcnmv.cu:188-279) and
 * fcn_binary_mm.binary_fcnmm_* (_fcn/binary_fcnmm.cu:487-530,836-880) */
XLA_FFI_Error* be_b200_xla_binary_fcnmv(XLA_FFI_CallFrame* frame);
XLA_FFI_Error* be_b200_xla_binary_fcnmm(XLA_FFI_CallFrame* frame);
/* scalar: (weight, clen, seed, events) -> (out); normal/uniform: (p0, p1, clen, seed, events) -> (out)
 * {kind:i32 (0 scalar,1 normal,2 uniform), corder:i32, chunk_size:i32}
 * replaces jit_{scalar,normal,uniform}_binary_jit*mv.{pack_bool,notrans_*,trans_*}
 * (_jit_scalar/binary_jitsmv.cu:216-465 etc.): the events are packed inside the handler. */
XLA_FFI_Error* be_b200_xla_binary_jitmv(XLA_FFI_CallFrame* frame);
/* same with events [vec_len, n] -> out [out_len, n]; replaces jit_*_binary_jit*mm.{pack,notrans_*,trans_*} */
XLA_FFI_Error* be_b200_xla_binary_jitmm(XLA_FFI_CallFrame* frame);
/* (weights[rows,cols], events) -> (out) {transpose}
 * replaces dense_binary_mv.* / dense_binary_mm.* (_dense/binary_densemv.cu:114-186,
 * _dense/binary_densemm.cu:96-199) */
XLA_FFI_Error* be_b200_xla_binary_densemv(XLA_FFI_CallFrame* frame);
XLA_FFI_Error* be_b200_xla_binary_densemm(XLA_FFI_CallFrame* frame);
/* (spikes[n]) -> (active_ids i32[n], n_active i32[1]); replaces compact.compact_1d_{bool,float}
 * (_event/compact.cu:132-178) with ascending order */
XLA_FFI_Error* be_b200_xla_compact_1d(XLA_FFI_CallFrame* frame);
/* (spikes[n,b]) -> (packed u32[n,ceil(b/32)], active_ids, n_active);
 * replaces compact.fused_bitpack_compact_2d_* (_event/compact.cu:387-430) */
XLA_FFI_Error* be_b200_xla_compact_2d(XLA_FFI_CallFrame* frame);
/* (spikes[n,b]) -> (active_ids, n_active); replaces compact.compact_2d_only_* (:234-283) */
XLA_FFI_Error* be_b200_xla_compact_2d_only(XLA_FFI_CallFrame* frame);
/* (spikes[n]) -> (packed u32[ceil(n/32)]); replaces jit_*.pack_bool (_jit_scalar/binary_jitsmv.cu:216-236) */
XLA_FFI_Error* be_b200_xla_bitpack_1d(XLA_FFI_CallFrame* frame);

/* ---- SURVEY 8(f) rows (same conventions; operands/weights share one float dtype) ---------------
 * csrmv/csrmm: args (weights, indices, indptr, x) -> ret out {transpose}
 * fcnmv/fcnmm: args (weights, indices [n_rows, n_conn], x) -> ret out {transpose}
 * jitmv/jitmm: args (p0, [p1], clen, seed, x) -> ret out {kind, corder, chunk_size}
 * plasticity_rows: args (weights, indices, [indptr], [perm], spikes, trace) -> ret updated weights
 *   {has_indptr (default 1; 0 = ELL indices [n_rows, n_conn]), has_perm (default 0), w_min, w_max
 *   (float scalars, optional)}; the ret buffer may alias arg 0 (input_output_aliases={0: 0})
 * binary_2d_{csr,csc}_encode: arg spikes [n_src, n_batch] -> rets (indices [n_src*n_batch], indptr)
 * binary_2d_pair_stream_encode: arg spikes -> rets (pair_stream [n_src*n_batch, 2], n_pairs [1])
 * binary_2d_row_sparse_encode: arg spikes -> rets (spike_indices [n_src, row_size], max_row_nnz [1]) */
XLA_FFI_Error* be_b200_xla_csrmv(XLA_FFI_CallFrame* frame);
XLA_FFI_Error* be_b200_xla_csrmm(XLA_FFI_CallFrame* frame);
XLA_FFI_Error* be_b200_xla_fcnmv(XLA_FFI_CallFrame* frame);
XLA_FFI_Error* be_b200_xla_fcnmm(XLA_FFI_CallFrame* frame);
XLA_FFI_Error* be_b200_xla_jitmv(XLA_FFI_CallFrame* frame);
XLA_FFI_Error* be_b200_xla_jitmm(XLA_FFI_CallFrame* frame);
XLA_FFI_Error* be_b200_xla_plasticity_rows(XLA_FFI_CallFrame* frame);
XLA_FFI_Error* be_b200_xla_binary_2d_csr_encode(XLA_FFI_CallFrame* frame);
XLA_FFI_Error* be_b200_xla_binary_2d_csc_encode(XLA_FFI_CallFrame* frame);
XLA_FFI_Error* be_b200_xla_binary_2d_pair_stream_encode(XLA_FFI_CallFrame* frame);
XLA_FFI_Error* be_b200_xla_binary_2d_row_sparse_encode(XLA_FFI_CallFrame* frame);


#ifdef __cplusplus
}
#endif
#endif /* BRAINEVENT_B200_XLA_H_ */
