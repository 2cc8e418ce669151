// float_ops.cu -- value-vector (non-event) products through the same connectivity formats:
// the callers either side of the event path.
//
// Replaces ref: _csr/float.py:49 csrmv / :559 csrmm (numba kernels :153-207, :670-746; CUDA
// _csr/float_csrmv.cu, float_csrmm.cu), _fcn/float.py:33 fcnmv / :136 fcnmm (+ _fcn/float_fcnmv.cu,
// float_fcnmm.cu).  JITC float products live in jitc.cu next to their event twins.
//
// Same building blocks as the event kernels (conn.cuh): a lane group streams a row's indices and
// weights with 128-bit L1-bypassing loads; the "transpose" orientation issues RED.ADD into an
// L2-resident output, the other orientation gathers x[indices[j]] and reduces in registers.
// No workspace: a dense value vector needs no compaction (zero rows are skipped in place).
#include "conn.cuh"

#include <algorithm>

namespace be {

// ------------------------------------------------------------------------------------------
// mv, transpose: out[indices[j]] += w[j] * x[i]
// ------------------------------------------------------------------------------------------
template <typename T, bool HOMO, int VEC, int LPR, typename Rows>
__global__ void __launch_bounds__(256) val_scatter_mv_kernel(Rows rows, const int32_t* __restrict__ indices,
                                                             const T* __restrict__ weights,
                                                             const T* __restrict__ x, int64_t m,
                                                             T* __restrict__ out) {
    using A = typename Acc<T>::type;
    const A w0 = HOMO ? to_acc<T>(__ldg(weights)) : A(0);
    constexpr int GROUPS = 256 / LPR;
    const int g = threadIdx.x / LPR, gl = threadIdx.x % LPR;
    for (int64_t r = (int64_t)blockIdx.x * GROUPS + g; r < m; r += (int64_t)gridDim.x * GROUPS) {
        const A xv = to_acc<T>(__ldg(x + r));
        if (xv == A(0)) continue;   // adds nothing (identical result for finite weights)
        int64_t s, e;
        rows.get(r, s, e);
        for_each_synapse<T, HOMO, 0, VEC, LPR>(s, e, gl, indices, weights, nullptr, w0,
                                               [&](int col, A w) { red_add(out + col, w * xv); });
    }
}

// ------------------------------------------------------------------------------------------
// mv, no transpose: out[i] = sum_j w[j] * x[indices[j]]   (deterministic, no atomics)
// ------------------------------------------------------------------------------------------
template <typename T, bool HOMO, int VEC, int LPR, typename Rows>
__global__ void __launch_bounds__(256) val_gather_mv_kernel(Rows rows, const int32_t* __restrict__ indices,
                                                            const T* __restrict__ weights,
                                                            const T* __restrict__ x, int64_t n_rows,
                                                            T* __restrict__ out) {
    using A = typename Acc<T>::type;
    const A w0 = HOMO ? to_acc<T>(__ldg(weights)) : A(0);
    constexpr int GROUPS = 256 / LPR;
    const int g = threadIdx.x / LPR, gl = threadIdx.x % LPR;
    // the loop bound is warp-uniform so the group reduction below can use the full mask
    for (int64_t base = (int64_t)blockIdx.x * GROUPS; base < n_rows; base += (int64_t)gridDim.x * GROUPS) {
        const int64_t r = base + g;
        A acc = A(0);
        if (r < n_rows) {
            int64_t s, e;
            rows.get(r, s, e);
            for_each_synapse<T, HOMO, 0, VEC, LPR>(s, e, gl, indices, weights, nullptr, w0, [&](int col, A w) {
                const A xv = to_acc<T>(__ldg(x + col));
                acc += HOMO ? xv : w * xv;
            });
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (r < n_rows && gl == 0) out[r] = from_acc<T>(HOMO ? acc * w0 : acc);
    }
}

// ------------------------------------------------------------------------------------------
// mm.  X is [rows, n] row-major, out is [rows_out, n] row-major: lanes = columns, so one synapse
// moves contiguous 128-byte segments.  A warp reads 32 synapses of its row at once and
// broadcasts them one by one.
// ------------------------------------------------------------------------------------------
template <typename T, bool HOMO, typename Rows>
__global__ void __launch_bounds__(256) val_scatter_mm_kernel(Rows rows, const int32_t* __restrict__ indices,
                                                             const T* __restrict__ weights,
                                                             const T* __restrict__ X, int64_t m, int64_t n,
                                                             T* __restrict__ out) {
    using A = typename Acc<T>::type;
    const A w0 = HOMO ? to_acc<T>(__ldg(weights)) : A(0);
    const int lane = threadIdx.x & 31;
    const int64_t n_ct = ceil_div(n, (int64_t)32);
    const int64_t n_tasks = m * n_ct;
    for (int64_t t = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); t < n_tasks; t += (int64_t)gridDim.x * 8) {
        const int64_t r = t / n_ct;
        const int64_t c = (t - r * n_ct) * 32 + lane;
        const A xv = c < n ? to_acc<T>(__ldg(X + r * n + c)) : A(0);
        int64_t s, e;
        rows.get(r, s, e);
        for (int64_t j0 = s; j0 < e; j0 += 32) {
            const int64_t j = j0 + lane;
            int my_col = 0;
            A my_w = A(0);
            if (j < e) { my_col = __ldg(indices + j); my_w = HOMO ? w0 : to_acc<T>(__ldg(weights + j)); }
            const int cnt = (int)min((int64_t)32, e - j0);
            for (int q = 0; q < cnt; ++q) {
                const int col = __shfl_sync(0xffffffffu, my_col, q);
                const A w = __shfl_sync(0xffffffffu, my_w, q);
                if (c < n && xv != A(0)) red_add(out + (int64_t)col * n + c, w * xv);
            }
        }
    }
}

template <typename T, bool HOMO, typename Rows>
__global__ void __launch_bounds__(256) val_gather_mm_kernel(Rows rows, const int32_t* __restrict__ indices,
                                                            const T* __restrict__ weights,
                                                            const T* __restrict__ X, int64_t n_rows, int64_t n,
                                                            T* __restrict__ out) {
    using A = typename Acc<T>::type;
    const A w0 = HOMO ? to_acc<T>(__ldg(weights)) : A(0);
    const int lane = threadIdx.x & 31;
    const int64_t n_ct = ceil_div(n, (int64_t)32);
    const int64_t n_tasks = n_rows * n_ct;
    for (int64_t t = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); t < n_tasks; t += (int64_t)gridDim.x * 8) {
        const int64_t r = t / n_ct;
        const int64_t c = (t - r * n_ct) * 32 + lane;
        int64_t s, e;
        rows.get(r, s, e);
        A acc = A(0);
        for (int64_t j0 = s; j0 < e; j0 += 32) {
            const int64_t j = j0 + lane;
            int my_col = 0;
            A my_w = A(0);
            if (j < e) { my_col = __ldg(indices + j); my_w = HOMO ? w0 : to_acc<T>(__ldg(weights + j)); }
            const int cnt = (int)min((int64_t)32, e - j0);
            int q = 0;
            for (; q + 4 <= cnt; q += 4) {     // four independent X rows in flight
                A xv[4], w[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int col = __shfl_sync(0xffffffffu, my_col, q + u);
                    w[u] = __shfl_sync(0xffffffffu, my_w, q + u);
                    xv[u] = c < n ? to_acc<T>(__ldg(X + (int64_t)col * n + c)) : A(0);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) acc += w[u] * xv[u];
            }
            for (; q < cnt; ++q) {
                const int col = __shfl_sync(0xffffffffu, my_col, q);
                const A w = __shfl_sync(0xffffffffu, my_w, q);
                if (c < n) acc += w * to_acc<T>(__ldg(X + (int64_t)col * n + c));
            }
        }
        if (c < n) out[r * n + c] = from_acc<T>(acc);
    }
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
template <typename T, bool HOMO, typename Rows>
static int run_val(const Rows& rows, const ConnArgs& c, const T* x, int64_t n_rows, int64_t n_other,
                   int64_t n, bool transpose, T* out, cudaStream_t s) {
    const int64_t out_rows = transpose ? n_other : n_rows;
    if (out_rows == 0) return BE_B200_OK;
    if (transpose) BE_CUDA(cudaMemsetAsync(out, 0, (size_t)out_rows * (size_t)std::max<int64_t>(n, 1) * sizeof(T), s));
    if (n_rows == 0) return BE_B200_OK;
    if (n == 0) {   // matvec
        const int vec = vec_mode(rows, c);
        const bool narrow = c.avg_row < 40.0;
        const int groups = narrow ? 32 : 8;
        const int blocks = (int)std::min<int64_t>(ceil_div(n_rows, groups), (int64_t)kNumSM * 16);
#define BE_VAL(K, VEC, LPR) K<T, HOMO, VEC, LPR, Rows><<<blocks, 256, 0, s>>>(rows, c.indices, (const T*)c.weights, x, n_rows, out)
#define BE_VAL_ALL(K)                                                                  \
        if (vec == 2) { if (narrow) BE_VAL(K, 2, 8); else BE_VAL(K, 2, 32); }           \
        else if (vec == 1) { if (narrow) BE_VAL(K, 1, 8); else BE_VAL(K, 1, 32); }      \
        else { if (narrow) BE_VAL(K, 0, 8); else BE_VAL(K, 0, 32); }
        {
            ProfileScope prof(s);
            if (transpose) { BE_VAL_ALL(val_scatter_mv_kernel) } else { BE_VAL_ALL(val_gather_mv_kernel) }
        }
#undef BE_VAL_ALL
#undef BE_VAL
        BE_LAUNCH_CHECK(transpose ? "val_scatter_mv_kernel" : "val_gather_mv_kernel");
        return BE_B200_OK;
    }
    const int64_t tasks = n_rows * ceil_div(n, (int64_t)32);
    const int blocks = (int)std::min<int64_t>(ceil_div(tasks, 8), (int64_t)kNumSM * 16);
    {
        ProfileScope prof(s);
        if (transpose) val_scatter_mm_kernel<T, HOMO, Rows><<<blocks, 256, 0, s>>>(rows, c.indices, (const T*)c.weights, x, n_rows, n, out);
        else val_gather_mm_kernel<T, HOMO, Rows><<<blocks, 256, 0, s>>>(rows, c.indices, (const T*)c.weights, x, n_rows, n, out);
    }
    BE_LAUNCH_CHECK(transpose ? "val_scatter_mm_kernel" : "val_gather_mm_kernel");
    return BE_B200_OK;
}

template <typename Rows>
static int dispatch_val(const Rows& rows, const ConnArgs& c, const void* x, int64_t n_rows, int64_t n_other,
                        int64_t n, bool transpose, void* out, cudaStream_t s) {
    BE_DISPATCH_WT(c.w_dtype, T, {
        return c.homo ? run_val<T, true>(rows, c, (const T*)x, n_rows, n_other, n, transpose, (T*)out, s)
                      : run_val<T, false>(rows, c, (const T*)x, n_rows, n_other, n, transpose, (T*)out, s);
    });
    return BE_B200_OK;
}

static int csr_val_entry(const void* weights, int w_dtype, int homo, const int32_t* indices,
                         const void* indptr, int indptr_is64, const void* x, int64_t m, int64_t k,
                         int64_t nse, int64_t n, int transpose, void* out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    BE_REQUIRE(m >= 0 && k >= 0 && nse >= 0 && m < (1ll << 31) && k < (1ll << 31), "csr: shape out of range");
    BE_REQUIRE(weights && indptr && out, "csr: NULL pointer");
    BE_REQUIRE(indices || nse == 0, "csr: indices is NULL");
    BE_REQUIRE(x || (transpose ? m : k) == 0, "csr: operand is NULL");
    ConnArgs c{weights, w_dtype, homo, indices, nullptr, 0, m > 0 ? (double)nse / (double)m : 0.0};
    if (indptr_is64) return dispatch_val(CsrRows<int64_t>{(const int64_t*)indptr}, c, x, m, k, n, transpose != 0, out, s);
    return dispatch_val(CsrRows<int32_t>{(const int32_t*)indptr}, c, x, m, k, n, transpose != 0, out, s);
}

static int fcn_val_entry(const void* weights, int w_dtype, int homo, const int32_t* indices,
                         int64_t n_rows, int64_t n_conn, const void* x, int64_t n_other, int64_t n,
                         int transpose, void* out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    BE_REQUIRE(n_rows >= 0 && n_conn >= 0 && n_other >= 0 && n_rows < (1ll << 31) && n_other < (1ll << 31), "fcn: shape out of range");
    BE_REQUIRE(weights && out, "fcn: NULL pointer");
    BE_REQUIRE(indices || n_rows * n_conn == 0, "fcn: indices is NULL");
    BE_REQUIRE(x || (transpose ? n_rows : n_other) == 0, "fcn: operand is NULL");
    ConnArgs c{weights, w_dtype, homo, indices, nullptr, 0, (double)n_conn};
    return dispatch_val(FcnRows{n_conn}, c, x, n_rows, n_other, n, transpose != 0, out, s);
}

}  // namespace be

using namespace be;

extern "C" int be_b200_csrmv(const void* weights, int w_dtype, int homo, const int32_t* indices,
                             const void* indptr, int indptr_is64, const void* x, int64_t m, int64_t k,
                             int64_t nse, int transpose, void* out, void* stream) { BE_NVTX_RANGE("be_b200_csrmv");
    return csr_val_entry(weights, w_dtype, homo, indices, indptr, indptr_is64, x, m, k, nse, 0, transpose, out, stream);
}

extern "C" int be_b200_csrmm(const void* weights, int w_dtype, int homo, const int32_t* indices,
                             const void* indptr, int indptr_is64, const void* X, int64_t m, int64_t k,
                             int64_t nse, int64_t n, int transpose, void* out, void* stream) { BE_NVTX_RANGE("be_b200_csrmm");
    BE_REQUIRE(n >= 1, "csrmm: n must be >= 1");
    return csr_val_entry(weights, w_dtype, homo, indices, indptr, indptr_is64, X, m, k, nse, n, transpose, out, stream);
}

extern "C" int be_b200_fcnmv(const void* weights, int w_dtype, int homo, const int32_t* indices,
                             int64_t n_rows, int64_t n_conn, const void* x, int64_t n_other,
                             int transpose, void* out, void* stream) { BE_NVTX_RANGE("be_b200_fcnmv");
    return fcn_val_entry(weights, w_dtype, homo, indices, n_rows, n_conn, x, n_other, 0, transpose, out, stream);
}

extern "C" int be_b200_fcnmm(const void* weights, int w_dtype, int homo, const int32_t* indices,
                             int64_t n_rows, int64_t n_conn, const void* X, int64_t n_other, int64_t n,
                             int transpose, void* out, void* stream) { BE_NVTX_RANGE("be_b200_fcnmm");
    BE_REQUIRE(n >= 1, "fcnmm: n must be >= 1");
    return fcn_val_entry(weights, w_dtype, homo, indices, n_rows, n_conn, X, n_other, n, transpose, out, stream);
}
