// conn.cuh -- pieces shared by the stored-connectivity translation units (csr_fcn.cu: scatter +
// C ABI; csr_fcn_gather.cu: gather kernels).
#pragma once
#include "common.cuh"

namespace be {

// ------------------------------------------------------------------------------------------
// row accessors
// ------------------------------------------------------------------------------------------
template <typename PT>
struct CsrRows {
    const PT* indptr;
    __device__ __forceinline__ void get(int64_t r, int64_t& s, int64_t& e) const {
        s = (int64_t)__ldg(indptr + r);
        e = (int64_t)__ldg(indptr + r + 1);
    }
};
struct FcnRows {
    int64_t n_conn;
    __device__ __forceinline__ void get(int64_t r, int64_t& s, int64_t& e) const {
        s = r * n_conn;
        e = s + n_conn;
    }
};

template <typename T, bool HOMO, int PERM>
__device__ __forceinline__ typename Acc<T>::type weight_at(const T* __restrict__ w,
                                                           const void* __restrict__ perm,
                                                           int64_t j, typename Acc<T>::type w0) {
    if (HOMO) return w0;
    if (PERM == 1) return to_acc<T>(__ldg(w + __ldg((const int32_t*)perm + j)));
    if (PERM == 2) return to_acc<T>(__ldg(w + __ldg((const int64_t*)perm + j)));
    return to_acc<T>(__ldg(w + j));
}


static inline bool al16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

struct ConnArgs {
    const void* weights;
    int w_dtype, homo;
    const int32_t* indices;
    const void* perm;
    int perm_kind;  // 0 none, 1 int32, 2 int64
    double avg_row;  // mean synapses per row (launch-shape heuristic only)
};

// type-erased row structure so the gather TU can be entered without templates
struct RowsDesc {
    int kind;            // 0: CSR int32 indptr, 1: CSR int64 indptr, 2: fixed n_conn (ELL)
    const void* indptr;
    int64_t n_conn;
};

// csr_fcn_gather.cu
int run_gather(const RowsDesc& rd, const ConnArgs& c, const Events& ev, int64_t n_src, int64_t n_rows,
               int64_t n_batch, void* out, cudaStream_t s);

}  // namespace be
