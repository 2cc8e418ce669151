// GPU-side reference arm -- TEST / BENCH INFRASTRUCTURE ONLY (never on the product path).
//
// the reference's own CUDA kernels compile stand-alone because their sources
// include only `cuda_common.h` + `brainevent/common.h`; the XLA dependency sits in the wrapper that
// `_op/kernix_codegen.py:617-736` generates.  This translation unit takes the place of that
// generated wrapper: it #includes ONE unmodified reference source from where it lies under
// /root/reference (selected with -DREF_SRC=..., nothing is copied into this repository) and
// exports its `// @BE` entry points behind a plain C ABI, building the `BE::Tensor` views by hand
// exactly as `docs/specs/cuda-kernel-audit.md:103-106` did.  Outputs go to oracle/_ref/ only.
//
// One .so per reference source: the JITC files reuse the names `pack_bool/notrans_f32/trans_f32`.
#include <cstdint>
#include <cstdio>
#include <exception>

#define REF_STR2(x) #x
#define REF_STR(x) REF_STR2(x)
#include REF_STR(REF_SRC)

extern "C" {

// dtype codes follow BE::DType (include/brainevent/tensor.h:42-59): 1 f32, 4 i8, 6 i32, 7 i64, 8 u8,
// 10 u32, 12 bool.
struct ref_tensor {
    void*   data;
    int64_t shape[4];
    int32_t ndim;
    int32_t dtype;
};

}  // extern "C"

static inline BE::Tensor T(const ref_tensor& t) {
    return BE::Tensor(t.data, t.shape, t.ndim, static_cast<BE::DType>(t.dtype));
}

#define REF_GUARD(call)                                                         \
    try { call; } catch (const std::exception& e) {                             \
        std::fprintf(stderr, "[ref_cuda] %s: %s\n", __func__, e.what());        \
        return -1;                                                              \
    }                                                                           \
    return 0;

extern "C" {

#if defined(REF_FCNMV)
// _fcn/binary_fcnmv.cu:188-252 -- (weights, indices, spikes) -> output
int ref_binary_fcnmv_scatter_homo_bool_f32(const ref_tensor* t, const int64_t*, int64_t s) {
    REF_GUARD(binary_fcnmv_scatter_homo_bool_f32(T(t[0]), T(t[1]), T(t[2]), T(t[3]), s)) }
int ref_binary_fcnmv_scatter_hetero_bool_f32(const ref_tensor* t, const int64_t*, int64_t s) {
    REF_GUARD(binary_fcnmv_scatter_hetero_bool_f32(T(t[0]), T(t[1]), T(t[2]), T(t[3]), s)) }
#endif

#if defined(REF_CSRMV)
// _csr/binary_csrmv.cu:362-434 -- (weights, indices, indptr, vector) -> output  (csr @ spk, NT gather)
int ref_binary_csrmv_nt_auto_homo_f32_bool(const ref_tensor* t, const int64_t*, int64_t s) {
    REF_GUARD(binary_csrmv_nt_auto_homo_f32_bool(T(t[0]), T(t[1]), T(t[2]), T(t[3]), T(t[4]), s)) }
int ref_binary_csrmv_nt_auto_hetero_f32_bool(const ref_tensor* t, const int64_t*, int64_t s) {
    REF_GUARD(binary_csrmv_nt_auto_hetero_f32_bool(T(t[0]), T(t[1]), T(t[2]), T(t[3]), T(t[4]), s)) }
#endif

#if defined(REF_CSRMV_HYBRID)
// _csr/binary_csrmv_hybrid.cu:618-787 -- (weights, indices, indptr, vector, task_begin, task_end,
// status) -> (output, task_begin, task_end, status); attrs[0] = task_capacity.  The scratch
// operands are aliased in->out by the reference's ffi_call (_csr/binary.py:597), so the same
// tensors are passed on both sides.  spk @ csr (T scatter), the GPU default.
int ref_binary_csrmv_wat_hybrid_homo_f32_bool(const ref_tensor* t, const int64_t* a, int64_t s) {
    REF_GUARD(binary_csrmv_wat_hybrid_homo_f32_bool(T(t[0]), T(t[1]), T(t[2]), T(t[3]), T(t[4]), T(t[5]), T(t[6]),
                                                    T(t[7]), T(t[4]), T(t[5]), T(t[6]), a[0], s)) }
int ref_binary_csrmv_wat_hybrid_hetero_f32_bool(const ref_tensor* t, const int64_t* a, int64_t s) {
    REF_GUARD(binary_csrmv_wat_hybrid_hetero_f32_bool(T(t[0]), T(t[1]), T(t[2]), T(t[3]), T(t[4]), T(t[5]), T(t[6]),
                                                      T(t[7]), T(t[4]), T(t[5]), T(t[6]), a[0], s)) }
#endif

#if defined(REF_CSRMM_HYBRID)
// _csr/binary_csrmm_hybrid.cu:552-754 -- matrix [n_pre, n_batch] int8, output physical [n_batch, n_post]
int ref_binary_csrmm_sraw_hybrid_homo_f32_bool(const ref_tensor* t, const int64_t* a, int64_t s) {
    REF_GUARD(binary_csrmm_sraw_hybrid_homo_f32_bool(T(t[0]), T(t[1]), T(t[2]), T(t[3]), T(t[4]), T(t[5]), T(t[6]),
                                                     T(t[7]), T(t[4]), T(t[5]), T(t[6]), a[0], s)) }
int ref_binary_csrmm_sraw_hybrid_hetero_f32_bool(const ref_tensor* t, const int64_t* a, int64_t s) {
    REF_GUARD(binary_csrmm_sraw_hybrid_hetero_f32_bool(T(t[0]), T(t[1]), T(t[2]), T(t[3]), T(t[4]), T(t[5]), T(t[6]),
                                                       T(t[7]), T(t[4]), T(t[5]), T(t[6]), a[0], s)) }
#endif

#if defined(REF_JITSMV)
// _jit_scalar/binary_jitsmv.cu:217-320 -- pack_bool(vector)->packed; (weight, clen, seed, packed) ->
// output with attrs vector_size, chunk_size
int ref_pack_bool(const ref_tensor* t, const int64_t*, int64_t s) { REF_GUARD(pack_bool(T(t[0]), T(t[1]), s)) }
int ref_notrans_f32(const ref_tensor* t, const int64_t* a, int64_t s) {
    REF_GUARD(notrans_f32(T(t[0]), T(t[1]), T(t[2]), T(t[3]), T(t[4]), (int)a[0], (int)a[1], s)) }
int ref_trans_f32(const ref_tensor* t, const int64_t* a, int64_t s) {
    REF_GUARD(trans_f32(T(t[0]), T(t[1]), T(t[2]), T(t[3]), T(t[4]), (int)a[0], (int)a[1], s)) }
#endif

#if defined(REF_JITNMV) || defined(REF_JITUMV)
// _jit_normal/binary_jitnmv.cu:316-430, _jit_uniform/binary_jitumv.cu:224-306 --
// (w_loc|w_low, w_scale|w_high, clen, seed, packed) -> output
int ref_pack_bool(const ref_tensor* t, const int64_t*, int64_t s) { REF_GUARD(pack_bool(T(t[0]), T(t[1]), s)) }
int ref_notrans_f32(const ref_tensor* t, const int64_t* a, int64_t s) {
    REF_GUARD(notrans_f32(T(t[0]), T(t[1]), T(t[2]), T(t[3]), T(t[4]), T(t[5]), (int)a[0], (int)a[1], s)) }
int ref_trans_f32(const ref_tensor* t, const int64_t* a, int64_t s) {
    REF_GUARD(trans_f32(T(t[0]), T(t[1]), T(t[2]), T(t[3]), T(t[4]), T(t[5]), (int)a[0], (int)a[1], s)) }
#endif

#if defined(REF_DENSEMM)
// _dense/binary_densemm.cu:119-164 -- (weights [k,m], spikes [k,n_batch]) -> output [n_batch,m]
int ref_binary_densemm_transpose_f32_bool(const ref_tensor* t, const int64_t*, int64_t s) {
    REF_GUARD(binary_densemm_transpose_f32_bool(T(t[0]), T(t[1]), T(t[2]), s)) }
#endif

#if defined(REF_COMPACT)
// _event/compact.cu:132-154 -- spikes -> (active_ids, n_active); unordered on the GPU
int ref_compact_1d_bool(const ref_tensor* t, const int64_t*, int64_t s) {
    REF_GUARD(compact_1d_bool(T(t[0]), T(t[1]), T(t[2]), s)) }
#endif

}  // extern "C"
