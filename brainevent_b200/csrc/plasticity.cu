// plasticity.cu -- spike-triggered additive weight updates over the same active-row traversal as
// the event scatter: for every active row r and every stored slot j of r
//     w[slot(j)] = clamp(w[slot(j)] + trace[indices[j]], lo, hi),   slot(j) = perm ? perm[j] : j
//
// Replaces ref: _csr/plasticity_binary.py (update_csr_on_binary_pre :45, numba :176-204;
// update_csr_on_binary_post :477, numba :621-654) + plasticity_binary_on_{pre,post}.cu;
// _fcn/plasticity_binary.py:81-109 + plasticity_row_driven.cu; _dense/plasticity_binary.py:144-164,
// :462-482 + plasticity_binary_on_{pre,post}.cu.
//
// In place, no atomics: a slot belongs to exactly one row (perm is a permutation), so the kernels
// are plain read-modify-write streams over the rows named by the compacted active list.  The
// reference clips the WHOLE weight array after the update (u.math.clip); be_b200_clip does that
// pass when asked, the row kernels clip what they touch.
#include "conn.cuh"

#include <algorithm>
#include <cmath>

namespace be {

template <typename A>
__device__ __forceinline__ A clampv(A v, A lo, A hi) { return v < lo ? lo : (v > hi ? hi : v); }

template <typename T, int PERM, int LPR, typename Rows>
__global__ void __launch_bounds__(256) plasticity_rows_kernel(Rows rows, const int32_t* __restrict__ indices,
                                                              const void* __restrict__ perm,
                                                              const int32_t* __restrict__ ids,
                                                              const int* __restrict__ n_active_ptr,
                                                              const T* __restrict__ trace, T* __restrict__ w,
                                                              typename Acc<T>::type lo, typename Acc<T>::type hi) {
    using A = typename Acc<T>::type;
    const int n_active = *n_active_ptr;
    constexpr int GROUPS = 256 / LPR;
    const int g = threadIdx.x / LPR, gl = threadIdx.x % LPR;
    for (int64_t a = (int64_t)blockIdx.x * GROUPS + g; a < n_active; a += (int64_t)gridDim.x * GROUPS) {
        const int r = __ldg(ids + a);
        int64_t s, e;
        rows.get(r, s, e);
        for (int64_t j = s + gl; j < e; j += LPR) {
            const A tr = to_acc<T>(__ldg(trace + __ldg(indices + j)));
            int64_t slot = j;
            if (PERM == 1) slot = __ldg((const int32_t*)perm + j);
            if (PERM == 2) slot = __ldg((const int64_t*)perm + j);
            w[slot] = from_acc<T>(clampv<A>(to_acc<T>(w[slot]) + tr, lo, hi));
        }
    }
}

template <typename T>
__global__ void clip_kernel(T* __restrict__ w, int64_t n, typename Acc<T>::type lo, typename Acc<T>::type hi) {
    using A = typename Acc<T>::type;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        w[i] = from_acc<T>(clampv<A>(to_acc<T>(w[i]), lo, hi));
}

// dense, pre-triggered: W[i, :] += trace[:] for active rows i (one CTA per active row)
template <typename T>
__global__ void __launch_bounds__(256) plasticity_dense_pre_kernel(T* __restrict__ W, int64_t cols,
                                                                   const int32_t* __restrict__ ids,
                                                                   const int* __restrict__ n_active_ptr,
                                                                   const T* __restrict__ trace,
                                                                   typename Acc<T>::type lo, typename Acc<T>::type hi) {
    using A = typename Acc<T>::type;
    const int n_active = *n_active_ptr;
    for (int a = blockIdx.x; a < n_active; a += gridDim.x) {
        T* row = W + (int64_t)__ldg(ids + a) * cols;
        for (int64_t c = threadIdx.x; c < cols; c += blockDim.x)
            row[c] = from_acc<T>(clampv<A>(to_acc<T>(row[c]) + to_acc<T>(__ldg(trace + c)), lo, hi));
    }
}

// dense, post-triggered: W[:, j] += trace[:] for active columns j; lanes run over the active list
template <typename T>
__global__ void __launch_bounds__(256) plasticity_dense_post_kernel(T* __restrict__ W, int64_t rows, int64_t cols,
                                                                    const int32_t* __restrict__ ids,
                                                                    const int* __restrict__ n_active_ptr,
                                                                    const T* __restrict__ trace,
                                                                    typename Acc<T>::type lo, typename Acc<T>::type hi) {
    using A = typename Acc<T>::type;
    const int n_active = *n_active_ptr;
    for (int64_t i = blockIdx.x; i < rows; i += gridDim.x) {
        const A tr = to_acc<T>(__ldg(trace + i));
        T* row = W + i * cols;
        for (int a = threadIdx.x; a < n_active; a += blockDim.x) {
            const int j = __ldg(ids + a);
            row[j] = from_acc<T>(clampv<A>(to_acc<T>(row[j]) + tr, lo, hi));
        }
    }
}

template <typename T, typename Rows>
static int run_rows(const Rows& rows, const int32_t* indices, const void* perm, int perm_kind, const Events& ev,
                    int64_t n_rows, double avg_row, const T* trace, T* w, double lo, double hi, cudaStream_t s) {
    using A = typename Acc<T>::type;
    const bool narrow = avg_row < 40.0;
    const int groups = narrow ? 32 : 8;
    const int blocks = (int)std::min<int64_t>(ceil_div(std::max<int64_t>(n_rows, 1), groups), (int64_t)kNumSM * 8);
#define BE_PL(PERM, LPR) plasticity_rows_kernel<T, PERM, LPR, Rows><<<blocks, 256, 0, s>>>(rows, indices, perm, ev.active_ids, ev.n_active, trace, w, (A)lo, (A)hi)
    {
        ProfileScope prof(s);
        if (perm_kind == 0) { if (narrow) BE_PL(0, 8); else BE_PL(0, 32); }
        else if (perm_kind == 1) { if (narrow) BE_PL(1, 8); else BE_PL(1, 32); }
        else { if (narrow) BE_PL(2, 8); else BE_PL(2, 32); }
    }
#undef BE_PL
    BE_LAUNCH_CHECK("plasticity_rows_kernel");
    return BE_B200_OK;
}

}  // namespace be

using namespace be;

extern "C" int be_b200_plasticity_rows(void* weights, int w_dtype, const int32_t* indices, const void* indptr,
                                       int indptr_is64, int64_t n_conn, const void* perm, int perm_is64,
                                       const void* spikes, const void* aux, int spk_kind, const void* trace,
                                       int64_t n_rows, int64_t n_trace, int64_t nse, double w_min, double w_max,
                                       void* workspace, size_t workspace_bytes, void* stream) { BE_NVTX_RANGE("be_b200_plasticity_rows");
    cudaStream_t s = (cudaStream_t)stream;
    BE_REQUIRE(n_rows >= 0 && n_trace >= 0 && nse >= 0 && n_rows < (1ll << 31) && n_trace < (1ll << 31), "plasticity_rows: shape out of range");
    BE_REQUIRE((weights && indices && trace) || nse == 0, "plasticity_rows: NULL pointer");
    BE_REQUIRE(n_conn >= 0 || indptr, "plasticity_rows: indptr is NULL");
    BE_REQUIRE(!(w_min > w_max), "plasticity_rows: w_min > w_max");
    if (n_rows == 0 || nse == 0) return BE_B200_OK;
    Events ev;
    int rc = prepare_events_1d(spikes, aux, spk_kind, n_rows, true, false, workspace, workspace_bytes, s, &ev,
                               nullptr, 0, /*ordered_ids=*/false, /*nonzero=*/true);
    if (rc) return rc;
    const int pk = perm ? (perm_is64 ? 2 : 1) : 0;
    const double avg = (double)nse / (double)n_rows;
    BE_DISPATCH_WT(w_dtype, T, {
        if (n_conn >= 0) return run_rows<T>(FcnRows{n_conn}, indices, perm, pk, ev, n_rows, avg, (const T*)trace, (T*)weights, w_min, w_max, s);
        if (indptr_is64) return run_rows<T>(CsrRows<int64_t>{(const int64_t*)indptr}, indices, perm, pk, ev, n_rows, avg, (const T*)trace, (T*)weights, w_min, w_max, s);
        return run_rows<T>(CsrRows<int32_t>{(const int32_t*)indptr}, indices, perm, pk, ev, n_rows, avg, (const T*)trace, (T*)weights, w_min, w_max, s);
    });
    return BE_B200_OK;
}

extern "C" int be_b200_plasticity_dense(void* W, int w_dtype, int64_t rows, int64_t cols, const void* spikes,
                                        const void* aux, int spk_kind, const void* trace, int on_post,
                                        double w_min, double w_max, void* workspace, size_t workspace_bytes,
                                        void* stream) { BE_NVTX_RANGE("be_b200_plasticity_dense");
    cudaStream_t s = (cudaStream_t)stream;
    BE_REQUIRE(rows >= 0 && cols >= 0 && rows < (1ll << 31) && cols < (1ll << 31), "plasticity_dense: shape out of range");
    BE_REQUIRE((W && trace) || rows * cols == 0, "plasticity_dense: NULL pointer");
    BE_REQUIRE(!(w_min > w_max), "plasticity_dense: w_min > w_max");
    if (rows == 0 || cols == 0) return BE_B200_OK;
    Events ev;
    int rc = prepare_events_1d(spikes, aux, spk_kind, on_post ? cols : rows, true, false, workspace, workspace_bytes, s,
                               &ev, nullptr, 0, /*ordered_ids=*/false, /*nonzero=*/true);
    if (rc) return rc;
    BE_DISPATCH_WT(w_dtype, T, {
        using A = typename Acc<T>::type;
        ProfileScope prof(s);
        if (on_post) plasticity_dense_post_kernel<T><<<(int)std::min<int64_t>(rows, (int64_t)kNumSM * 8), 256, 0, s>>>((T*)W, rows, cols, ev.active_ids, ev.n_active, (const T*)trace, (A)w_min, (A)w_max);
        else plasticity_dense_pre_kernel<T><<<(int)std::min<int64_t>(rows, (int64_t)kNumSM * 8), 256, 0, s>>>((T*)W, cols, ev.active_ids, ev.n_active, (const T*)trace, (A)w_min, (A)w_max);
    });
    BE_LAUNCH_CHECK("plasticity_dense_kernel");
    return BE_B200_OK;
}

extern "C" int be_b200_clip(void* w, int w_dtype, int64_t n, double w_min, double w_max, void* stream) { BE_NVTX_RANGE("be_b200_clip");
    cudaStream_t s = (cudaStream_t)stream;
    BE_REQUIRE(n >= 0 && (w || n == 0), "clip: bad arguments");
    BE_REQUIRE(!(w_min > w_max), "clip: w_min > w_max");
    if (n == 0) return BE_B200_OK;
    const int blocks = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)kNumSM * 16);
    BE_DISPATCH_WT(w_dtype, T, {
        using A = typename Acc<T>::type;
        clip_kernel<T><<<blocks, 256, 0, s>>>((T*)w, n, (A)w_min, (A)w_max);
    });
    BE_LAUNCH_CHECK("clip_kernel");
    return BE_B200_OK;
}
