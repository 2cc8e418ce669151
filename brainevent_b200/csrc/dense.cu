// dense.cu -- event-driven products with a dense weight matrix.
//
// Replaces ref: _dense/binary_densemv.cu, _dense/binary_densemm.cu (and the optional
// _dense/binary_dense_cublas.cu).  W is row-major [rows, cols].
//   mv, transpose=1: out[cols] = sum over active i of W[i, :]      (reads only active rows)
//   mv, transpose=0: out[rows] = sum over active j of W[:, j]      (sparse: gathers W[r, ids];
//                                                                   dense: streams the row)
//   mm, transpose=1: W[k, m], events[k, n] -> out[m, n]
//   mm, transpose=0: W[m, k], events[k, n] -> out[m, n]
// The reference launches one warp per (row, batch column) and re-reads the weight row for every
// column; here each weight element is read once per call and applied to all columns.
#include "common.cuh"

#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace be {

// ---- mv, transpose=1 ---------------------------------------------------------------------
// thread = 4 consecutive output columns; gridDim.y slices the active list.
template <typename T>
__global__ void __launch_bounds__(256) dense_mv_t_kernel(const T* __restrict__ W, int64_t cols,
                                                         const int32_t* __restrict__ ids,
                                                         const int* __restrict__ n_active_ptr,
                                                         T* __restrict__ out, bool vec) {
    using A = typename Acc<T>::type;
    const int n_active = *n_active_ptr;
    const int64_t c0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (c0 >= cols) return;
    A acc[4] = {A(0), A(0), A(0), A(0)};
    const bool full = vec && (c0 + 4 <= cols);
    // U active rows per trip: their ids first, then U independent 128-bit row loads, then the adds -- with one
    // row per trip a thread had a single load in flight behind a dependent id load (65536 x 32768 at 1 %: 51 us for
    // 86 MB, 1.7 TB/s)
    constexpr int U = 8;
    for (int a0 = blockIdx.y * U; a0 < n_active; a0 += gridDim.y * U) {
        int rid[U];
#pragma unroll
        for (int u = 0; u < U; ++u) rid[u] = (a0 + u < n_active) ? __ldg(ids + a0 + u) : -1;
        if (full) {
            Vec4<T> v[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (rid[u] >= 0) v[u] = load4(W + (int64_t)rid[u] * cols + c0);
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (rid[u] >= 0) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[q] += v[u].v[q];
                }
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (rid[u] >= 0) {
                    const T* row = W + (int64_t)rid[u] * cols + c0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) if (c0 + q < cols) acc[q] += to_acc<T>(__ldg(row + q));
                }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (c0 + q >= cols) break;
        if (gridDim.y == 1) out[c0 + q] = from_acc<T>(acc[q]);
        else if (acc[q] != A(0)) red_add(out + c0 + q, acc[q]);
    }
}

// ---- mv, transpose=0 ---------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) dense_mv_nt_kernel(const T* __restrict__ W, int64_t rows,
                                                          int64_t cols, const uint32_t* __restrict__ packed,
                                                          const int32_t* __restrict__ ids,
                                                          const int* __restrict__ n_active_ptr,
                                                          T* __restrict__ out) {
    using A = typename Acc<T>::type;
    const int n_active = *n_active_ptr;
    const int lane = threadIdx.x & 31;
    const bool sparse = (int64_t)n_active * 8 < cols;   // one 32-B sector per gathered element
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * 8) {
        const T* row = W + r * cols;
        A acc = A(0);
        if (sparse) {
            for (int a = lane; a < n_active; a += 32) acc += to_acc<T>(__ldg(row + __ldg(ids + a)));
        } else {
            for (int64_t j = lane; j < cols; j += 32)
                if (bit_test(packed, j)) acc += to_acc<T>(__ldg(row + j));
        }
        acc = warp_sum(acc);
        if (lane == 0) out[r] = from_acc<T>(acc);
    }
}

// ---- mm, transpose=1: W[k, m], packed events [k, wpr] -> out[m, n] -------------------------
// thread = one output row r; per-pass accumulators acc[c][t] live in shared memory (padded so the
// transposed write-out is conflict-free); gridDim.y slices the active list.
constexpr int kDmmCols = 64;   // batch columns per pass
template <typename T>
__global__ void __launch_bounds__(256) dense_mm_t_kernel(const T* __restrict__ W, int64_t m,
                                                         const uint32_t* __restrict__ packed, int wpr,
                                                         const int32_t* __restrict__ ids,
                                                         const int* __restrict__ n_active_ptr,
                                                         int64_t n, T* __restrict__ out) {
    using A = typename Acc<T>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    A* acc = reinterpret_cast<A*>(smem_raw);   // [kDmmCols][257]
    const int n_active = *n_active_ptr;
    const int t = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * 256;
    const int64_t r = r0 + t;
    const int rows_here = (int)min((int64_t)256, m - r0);
    for (int64_t cb = 0; cb < n; cb += kDmmCols) {
        const int ncol = (int)min((int64_t)kDmmCols, n - cb);
        const int w0 = (int)(cb >> 5);
        for (int c = 0; c < ncol; ++c) acc[c * 257 + t] = A(0);
        for (int a = blockIdx.y; a < n_active; a += gridDim.y) {
            const int64_t i = __ldg(ids + a);
            uint32_t m0 = __ldg(packed + i * wpr + w0);
            uint32_t m1 = (w0 + 1 < wpr) ? __ldg(packed + i * wpr + w0 + 1) : 0u;
            if ((m0 | m1) == 0u) continue;
            const A wv = (r < m) ? to_acc<T>(__ldg(W + i * m + r)) : A(0);
            while (m0) { int c = __ffs(m0) - 1; m0 &= m0 - 1; acc[c * 257 + t] += wv; }
            while (m1) { int c = __ffs(m1) - 1; m1 &= m1 - 1; acc[(32 + c) * 257 + t] += wv; }
        }
        __syncthreads();
        for (int e = t; e < rows_here * ncol; e += 256) {
            const int rl = e / ncol, c = e - rl * ncol;
            const A v = acc[c * 257 + rl];
            T* o = out + (r0 + rl) * n + cb + c;
            if (gridDim.y == 1) *o = from_acc<T>(v);
            else if (v != A(0)) red_add(o, v);
        }
        __syncthreads();
    }
}

// ---- mm, transpose=0: W[m, k], packed events [k, wpr] -> out[m, n] -------------------------
// warp = output row, lanes = batch columns; W row streamed coalesced, zero-mask sources skipped.
template <typename T, int CPL>
__global__ void __launch_bounds__(256) dense_mm_nt_kernel(const T* __restrict__ W, int64_t m, int64_t k,
                                                          const uint32_t* __restrict__ packed, int wpr,
                                                          int64_t n, T* __restrict__ out) {
    using A = typename Acc<T>::type;
    const int lane = threadIdx.x & 31;
    const int n_pass = (wpr + CPL - 1) / CPL;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < m; r += (int64_t)gridDim.x * 8) {
        const T* row = W + r * k;
        for (int pass = 0; pass < n_pass; ++pass) {
            const int wb = pass * CPL;
            A acc[CPL];
#pragma unroll
            for (int q = 0; q < CPL; ++q) acc[q] = A(0);
            for (int64_t j0 = 0; j0 < k; j0 += 32) {
                const int64_t j = j0 + lane;
                uint32_t mw[CPL];
                bool nz = false;
#pragma unroll
                for (int q = 0; q < CPL; ++q) {
                    mw[q] = (j < k && wb + q < wpr) ? __ldg(packed + j * wpr + wb + q) : 0u;
                    nz |= (mw[q] != 0u);
                }
                const A my_w = nz ? to_acc<T>(__ldg(row + j)) : A(0);
                uint32_t live = __ballot_sync(0xffffffffu, nz);
                while (live) {
                    const int tt = __ffs(live) - 1;
                    live &= live - 1;
                    const A wv = __shfl_sync(0xffffffffu, my_w, tt);
#pragma unroll
                    for (int q = 0; q < CPL; ++q) {
                        const uint32_t bits = __shfl_sync(0xffffffffu, mw[q], tt);
                        if ((bits >> lane) & 1u) acc[q] += wv;
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
                const int64_t c = (int64_t)(wb + q) * 32 + lane;
                if (c < n) out[r * n + c] = from_acc<T>(acc[q]);
            }
        }
    }
}

template <typename T>
static int run_dense_mv(const T* W, int64_t rows, int64_t cols, const Events& ev, bool transpose,
                        T* out, cudaStream_t s) {
    if (transpose) {
        if (cols == 0) return BE_B200_OK;
        const int64_t col_blocks = ceil_div(ceil_div(cols, 4), 256);
        int slices = (int)std::max<int64_t>(1, std::min<int64_t>(64, (4 * kNumSM) / col_blocks));
        if (slices > 1) BE_CUDA(cudaMemsetAsync(out, 0, (size_t)cols * sizeof(T), s));
        const bool vec = (((uintptr_t)W & 15u) == 0) && ((cols * sizeof(T)) % 16 == 0);
        dim3 grid((unsigned)col_blocks, (unsigned)slices);
        { ProfileScope prof(s); dense_mv_t_kernel<T><<<grid, 256, 0, s>>>(W, cols, ev.active_ids, ev.n_active, out, vec); }
        BE_LAUNCH_CHECK("dense_mv_t_kernel");
    } else {
        if (rows == 0) return BE_B200_OK;
        int blocks = (int)std::min<int64_t>(ceil_div(rows, 8), (int64_t)kNumSM * 16);
        { ProfileScope prof(s); dense_mv_nt_kernel<T><<<blocks, 256, 0, s>>>(W, rows, cols, ev.packed, ev.active_ids, ev.n_active, out); }
        BE_LAUNCH_CHECK("dense_mv_nt_kernel");
    }
    return BE_B200_OK;
}

// dense_tc.cu
int launch_dense_mm_t_tc(const float* W, int64_t k, int64_t m, const uint32_t* packed, int wpr, int64_t n,
                         float* out, cudaStream_t s);

// BE_B200_DENSE_TC=0 keeps the SIMT kernel (A/B measurements); default: tensor cores at batch >= 64
static bool dense_tc_enabled() {
    const char* e = getenv("BE_B200_DENSE_TC");
    return !(e && e[0] == '0');
}

template <typename T>
static int run_dense_mm(const T* W, int64_t rows, int64_t cols, const Events& ev, int64_t n,
                        bool transpose, T* out, cudaStream_t s) {
    using A = typename Acc<T>::type;
    const int wpr = ev.words_per_row;
    if (transpose) {
        const int64_t m = cols;
        if (m == 0) return BE_B200_OK;
        if constexpr (std::is_same<T, float>::value) {
            // the one real contraction on the path: fp32-exact 3 x bf16 split on tcgen05
            if (n >= 64 && (((uintptr_t)W & 15u) == 0) && (m % 4 == 0) && dense_tc_enabled())
                return launch_dense_mm_t_tc(W, rows, m, ev.packed, wpr, n, out, s);
        }
        const int64_t row_blocks = ceil_div(m, 256);
        int slices = (int)std::max<int64_t>(1, std::min<int64_t>(32, (2 * kNumSM) / row_blocks));
        if (slices > 1) BE_CUDA(cudaMemsetAsync(out, 0, (size_t)(m * n) * sizeof(T), s));
        const size_t smem = (size_t)kDmmCols * 257 * sizeof(A);
        // the attribute is per device/context: set it on every call (cheap) rather than caching a process-wide flag
        BE_CUDA(cudaFuncSetAttribute(dense_mm_t_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((unsigned)row_blocks, (unsigned)slices);
        { ProfileScope prof(s); dense_mm_t_kernel<T><<<grid, 256, smem, s>>>(W, m, ev.packed, wpr, ev.active_ids, ev.n_active, n, out); }
        BE_LAUNCH_CHECK("dense_mm_t_kernel");
    } else {
        const int64_t m = rows, k = cols;
        if (m == 0) return BE_B200_OK;
        int blocks = (int)std::min<int64_t>(ceil_div(m, 8), (int64_t)kNumSM * 16);
#define BE_DN(CPL) dense_mm_nt_kernel<T, CPL><<<blocks, 256, 0, s>>>(W, m, k, ev.packed, wpr, n, out)
        { ProfileScope prof(s); if (wpr == 1) BE_DN(1); else if (wpr == 2) BE_DN(2); else BE_DN(4); }
#undef BE_DN
        BE_LAUNCH_CHECK("dense_mm_nt_kernel");
    }
    return BE_B200_OK;
}

}  // namespace be

using namespace be;

extern "C" int be_b200_binary_densemv(const void* W, int w_dtype, int64_t rows, int64_t cols,
                                      const void* spikes, const void* aux, int spk_kind,
                                      int transpose, void* out, void* workspace,
                                      size_t workspace_bytes, void* stream) { BE_NVTX_RANGE("be_b200_binary_densemv");
    cudaStream_t s = (cudaStream_t)stream;
    BE_REQUIRE(rows >= 0 && cols >= 0 && rows < (1ll << 31) && cols < (1ll << 31), "densemv: shape out of range");
    BE_REQUIRE((W || rows * cols == 0) && out, "densemv: NULL pointer");
    const int64_t n_src = transpose ? rows : cols;
    Events ev;
    int rc = prepare_events_1d(spikes, aux, spk_kind, n_src, true, !transpose, workspace, workspace_bytes, s, &ev);
    if (rc) return rc;
    BE_DISPATCH_WT(w_dtype, T, { return run_dense_mv<T>((const T*)W, rows, cols, ev, transpose != 0, (T*)out, s); });
    return BE_B200_OK;
}

extern "C" int be_b200_binary_densemm(const void* W, int w_dtype, int64_t rows, int64_t cols,
                                      const void* spikes, const void* aux, int spk_kind, int64_t n,
                                      int transpose, void* out, void* workspace,
                                      size_t workspace_bytes, void* stream) { BE_NVTX_RANGE("be_b200_binary_densemm");
    cudaStream_t s = (cudaStream_t)stream;
    BE_REQUIRE(rows >= 0 && cols >= 0 && rows < (1ll << 31) && cols < (1ll << 31) && n >= 1, "densemm: shape out of range");
    BE_REQUIRE((W || rows * cols == 0) && out, "densemm: NULL pointer");
    const int64_t n_src = transpose ? rows : cols;
    Events ev;
    int rc = prepare_events_2d(spikes, aux, spk_kind, n_src, n, transpose != 0, workspace, workspace_bytes, s, &ev);
    if (rc) return rc;
    BE_DISPATCH_WT(w_dtype, T, { return run_dense_mm<T>((const T*)W, rows, cols, ev, n, transpose != 0, (T*)out, s); });
    return BE_B200_OK;
}
