// conn.cuh -- pieces shared by the stored-connectivity translation units (csr_fcn.cu: scatter +
// C ABI; csr_fcn_gather.cu: gather kernels).
#pragma once
#include "common.cuh"

#include <type_traits>

namespace be {

// ------------------------------------------------------------------------------------------
// row accessors
// ------------------------------------------------------------------------------------------
template <typename PT>
struct CsrRows {
    using ptr_t = PT;
    const PT* indptr;
    __device__ __forceinline__ void get(int64_t r, int64_t& s, int64_t& e) const {
        s = (int64_t)__ldg(indptr + r);
        e = (int64_t)__ldg(indptr + r + 1);
    }
    // the extent in the stored width: callers that prefetch a row's extent keep the raw values and widen them
    // where they are used, so no instruction consumes the load right behind it
    __device__ __forceinline__ void get_raw(int64_t r, PT& s, PT& e) const {
        s = __ldg(indptr + r);
        e = __ldg(indptr + r + 1);
    }
    __device__ __forceinline__ int64_t total(int64_t n_rows) const { return (int64_t)__ldg(indptr + n_rows); }
};
struct FcnRows {
    using ptr_t = int64_t;
    int64_t n_conn;
    __device__ __forceinline__ void get_raw(int64_t r, int64_t& s, int64_t& e) const { get(r, s, e); }
    __device__ __forceinline__ void get(int64_t r, int64_t& s, int64_t& e) const {
        s = r * n_conn;
        e = s + n_conn;
    }
    __device__ __forceinline__ int64_t total(int64_t n_rows) const { return n_rows * n_conn; }
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


// Visit every synapse j in [s, e) of one row with a group of G lanes: f(column, weight).
// VEC: peel to a 4-element boundary, then one 128-bit index load (+ one 128-bit weight load) per
// lane per step, all issued before the first use so a row costs ~one memory round trip.
template <typename T, bool HOMO, int PERM, int VEC, int G, typename F>
__device__ __forceinline__ void for_each_synapse(int64_t s, int64_t e, int gl,
                                                 const int32_t* __restrict__ indices,
                                                 const T* __restrict__ weights,
                                                 const void* __restrict__ perm,
                                                 typename Acc<T>::type w0, F&& f) {
    using A = typename Acc<T>::type;
    if (VEC == 0) {
        for (int64_t j = s + gl; j < e; j += G)
            f(__ldg(indices + j), weight_at<T, HOMO, PERM>(weights, perm, j, w0));
        return;
    }
    // One loop covers head, body and tail: quads start at the 16-byte boundary at or below s, so
    // every lane issues aligned 128-bit loads (the first quad may reach up to 3 slots into the
    // previous row: masked, in bounds); only the single partial quad at the row's end falls back to
    // scalar loads.  No separate head/tail instruction streams, no extra partially-filled REDs.
    if (VEC == 2) {
        // rows are known to start and end on 4-element boundaries (ELL with n_conn % 4 == 0):
        // pure 128-bit streaming, no boundary handling at all
#pragma unroll 2
        for (int64_t j = s + (int64_t)gl * 4; j < e; j += (int64_t)G * 4) {
            const int4 ix = ldg_stream((const int4*)(indices + j));
            A wv[4];
            if (HOMO) { wv[0] = wv[1] = wv[2] = wv[3] = w0; }
            else if (PERM == 0) { Vec4<T> v = load4(weights + j); wv[0] = v.v[0]; wv[1] = v.v[1]; wv[2] = v.v[2]; wv[3] = v.v[3]; }
            else {
#pragma unroll
                for (int q = 0; q < 4; ++q) wv[q] = weight_at<T, HOMO, PERM>(weights, perm, j + q, w0);
            }
            f(ix.x, wv[0]);
            f(ix.y, wv[1]);
            f(ix.z, wv[2]);
            f(ix.w, wv[3]);
        }
        return;
    }
    const int64_t a0 = s & ~(int64_t)3;
#pragma unroll 2
    for (int64_t j = a0 + (int64_t)gl * 4; j < e; j += (int64_t)G * 4) {
        int ixv[4];
        A wv[4];
        if (j + 4 <= e) {
            const int4 ix = ldg_stream((const int4*)(indices + j));
            ixv[0] = ix.x; ixv[1] = ix.y; ixv[2] = ix.z; ixv[3] = ix.w;
            if (HOMO) { wv[0] = wv[1] = wv[2] = wv[3] = w0; }
            else if (PERM == 0) { Vec4<T> v = load4(weights + j); wv[0] = v.v[0]; wv[1] = v.v[1]; wv[2] = v.v[2]; wv[3] = v.v[3]; }
            else {
#pragma unroll
                for (int q = 0; q < 4; ++q) wv[q] = (j + q >= s) ? weight_at<T, HOMO, PERM>(weights, perm, j + q, w0) : A(0);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool ok = (j + q >= s) && (j + q < e);
                ixv[q] = ok ? __ldg(indices + j + q) : 0;
                wv[q] = ok ? weight_at<T, HOMO, PERM>(weights, perm, j + q, w0) : A(0);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (j + q >= s && j + q < e) f(ixv[q], wv[q]);
    }
}


static inline bool al16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

struct ConnArgs {
    const void* weights;
    int w_dtype, homo;
    const int32_t* indices;
    const void* perm;
    int perm_kind;  // 0 none, 1 int32, 2 int64
    double avg_row;  // mean synapses per row (launch-shape heuristic only)
    int64_t nse = 0; // stored entries (0 = unknown: bulk-copy paths that need the array's end stay off)
};

// type-erased row structure so the gather TU can be entered without templates
struct RowsDesc {
    int kind;            // 0: CSR int32 indptr, 1: CSR int64 indptr, 2: fixed n_conn (ELL)
    const void* indptr;
    int64_t n_conn;
};

// 0: scalar accesses (unaligned base pointers); 1: masked 128-bit quads (arbitrary row extents);
// 2: rows start/end on 4-element boundaries (ELL with n_conn % 4 == 0)
template <typename Rows>
static inline int vec_mode(const Rows& rows, const ConnArgs& c) {
    if (!(al16(c.indices) && (c.homo || al16(c.weights)))) return 0;
    if constexpr (std::is_same<Rows, FcnRows>::value) return (rows.n_conn % 4 == 0) ? 2 : 1;
    return 1;
}


// csr_fcn_gather.cu
int run_gather(const RowsDesc& rd, const ConnArgs& c, const Events& ev, int64_t n_src, int64_t n_rows,
               int64_t n_batch, void* out, cudaStream_t s);

}  // namespace be
