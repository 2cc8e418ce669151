// encode.cu -- dense 2-D spike matrix -> sparse event layouts: the spike encoders either side of
// the propagation path.
//
// Replaces ref: _event/compact.py binary_2d_pair_stream_encode (:634-670), binary_2d_row_sparse_encode
// (:792-826), binary_2d_csr_row_count / _csr_fill / _csr_encode (:962-1217), binary_2d_csc_encode
// (:1220-1262) and the encode kernels of _event/compact.cu.
//
// All of them are ORDERED compactions of the row-major matrix, so they sit on the 1-D compaction of
// events.cu (decoupled look-back, ascending flat positions) over the flattened n_src*n_batch
// elements: pair stream = (pos / n_batch, pos % n_batch), CSR indices = pos % n_batch, CSR indptr =
// lower bounds of the row starts in the sorted position list.  CSC is the same over a transposed
// byte mask.  Unlike the reference CUDA pair stream ("order not guaranteed") the output here is
// always in the numba kernels' row-major order.  Float spikes are active when != 0.
#include "common.cuh"

#include <algorithm>

namespace be {

__global__ void pair_stream_kernel(const int32_t* __restrict__ ids, const int* __restrict__ n_active_ptr,
                                   int n_batch, int2* __restrict__ pairs, int32_t* __restrict__ n_pairs) {
    const int n = *n_active_ptr;
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_pairs = n;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int id = __ldg(ids + p);
        pairs[p] = make_int2(id / n_batch, id % n_batch);
    }
}

// minor ids (id % inner) and the row-pointer array: ptr[r] = #{ids < r * inner}, r = 0..n_outer
__global__ void minor_and_ptr_kernel(const int32_t* __restrict__ ids, const int* __restrict__ n_active_ptr,
                                     int64_t n_outer, int inner, int32_t* __restrict__ indices,
                                     int32_t* __restrict__ indptr) {
    const int n = *n_active_ptr;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = tid; p < n; p += nth) indices[p] = __ldg(ids + p) % inner;
    for (int64_t r = tid; r <= n_outer; r += nth) {
        const int64_t key = r * inner;
        int lo = 0, hi = n;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((int64_t)__ldg(ids + mid) < key) lo = mid + 1; else hi = mid;
        }
        indptr[r] = lo;
    }
}

template <int KIND>
__device__ __forceinline__ bool spike_at(const void* spikes, int64_t i) {
    if (KIND == BE_B200_SPK_BOOL) return ((const uint8_t*)spikes)[i] != 0;
    return ((const float*)spikes)[i] != 0.0f;
}

// one warp per source row: active batch columns (1-based) compacted to the front of a fixed-width row
template <int KIND>
__global__ void __launch_bounds__(256) row_sparse_kernel(const void* __restrict__ spikes, int64_t n_src, int n_batch,
                                                         int row_size, int32_t* __restrict__ out,
                                                         int* __restrict__ max_count) {
    const int lane = threadIdx.x & 31;
    int worst = 0;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < n_src; r += (int64_t)gridDim.x * 8) {
        int pos = 0;
        int32_t* orow = out + r * row_size;
        for (int c0 = 0; c0 < n_batch; c0 += 32) {
            const int c = c0 + lane;
            const bool on = c < n_batch && spike_at<KIND>(spikes, r * n_batch + c);
            const unsigned m = __ballot_sync(0xffffffffu, on);
            const int my = pos + __popc(m & ((1u << lane) - 1u));
            if (on && my < row_size) orow[my] = c + 1;
            pos += __popc(m);
        }
        for (int k = min(pos, row_size) + lane; k < row_size; k += 32) orow[k] = 0;
        worst = max(worst, pos);
    }
    if (lane == 0 && worst > 0) atomicMax(max_count, worst);
}

// maskT[b, r] = spikes[r, b] != 0 through a 32x32 shared tile (coalesced both ways)
template <int KIND>
__global__ void transpose_mask_kernel(const void* __restrict__ spikes, int64_t n_src, int64_t n_batch,
                                      uint8_t* __restrict__ maskT) {
    __shared__ uint8_t tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32, b0 = (int64_t)blockIdx.y * 32;   // rows on x: n_src can be large
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int64_t r = r0 + dy, b = b0 + threadIdx.x;
        tile[dy][threadIdx.x] = (r < n_src && b < n_batch) ? (uint8_t)spike_at<KIND>(spikes, r * n_batch + b) : 0;
    }
    __syncthreads();
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int64_t b = b0 + dy, r = r0 + threadIdx.x;
        if (b < n_batch && r < n_src) maskT[b * n_src + r] = tile[threadIdx.x][dy];
    }
}

static int check_2d(const void* spikes, int spk_kind, int64_t n_src, int64_t n_batch, const char* who) {
    BE_REQUIRE(spk_kind == BE_B200_SPK_BOOL || spk_kind == BE_B200_SPK_F32, "%s: spikes must be bool/uint8 or float32", who);
    BE_REQUIRE(n_src >= 0 && n_batch >= 0 && n_src * n_batch < (1ll << 31), "%s: n_src * n_batch must be < 2^31", who);
    BE_REQUIRE(spikes || n_src * n_batch == 0, "%s: spikes is NULL", who);
    return BE_B200_OK;
}

static int flat_ids(const void* spikes, int spk_kind, int64_t n, void* ws, size_t ws_bytes, cudaStream_t s, Events* ev) {
    return prepare_events_1d(spikes, nullptr, spk_kind, n, true, false, ws, ws_bytes, s, ev, nullptr, 0,
                             /*ordered_ids=*/true, /*nonzero=*/true);
}

static int grid_for(int64_t n) { return (int)std::min<int64_t>(ceil_div(std::max<int64_t>(n, 1), 256), (int64_t)kNumSM * 8); }

}  // namespace be

using namespace be;

extern "C" int be_b200_binary_2d_pair_stream_encode(const void* spikes, int spk_kind, int64_t n_src, int64_t n_batch,
                                                    int32_t* pair_stream, int32_t* n_pairs, void* workspace,
                                                    size_t workspace_bytes, void* stream) { BE_NVTX_RANGE("be_b200_binary_2d_pair_stream_encode");
    cudaStream_t s = (cudaStream_t)stream;
    if (int rc = check_2d(spikes, spk_kind, n_src, n_batch, "pair_stream_encode")) return rc;
    BE_REQUIRE(n_pairs && (pair_stream || n_src * n_batch == 0), "pair_stream_encode: NULL output");
    const int64_t n = n_src * n_batch;
    if (n > 0) BE_CUDA(cudaMemsetAsync(pair_stream, 0, (size_t)n * 8, s));
    Events ev;
    if (int rc = flat_ids(spikes, spk_kind, n, workspace, workspace_bytes, s, &ev)) return rc;
    pair_stream_kernel<<<grid_for(n), 256, 0, s>>>(ev.active_ids, ev.n_active, (int)std::max<int64_t>(n_batch, 1),
                                                   (int2*)pair_stream, n_pairs);
    BE_LAUNCH_CHECK("pair_stream_kernel");
    return BE_B200_OK;
}

extern "C" int be_b200_binary_2d_csr_encode(const void* spikes, int spk_kind, int64_t n_src, int64_t n_batch,
                                            int32_t* indices, int32_t* indptr, void* workspace,
                                            size_t workspace_bytes, void* stream) { BE_NVTX_RANGE("be_b200_binary_2d_csr_encode");
    cudaStream_t s = (cudaStream_t)stream;
    if (int rc = check_2d(spikes, spk_kind, n_src, n_batch, "csr_encode")) return rc;
    BE_REQUIRE(indptr && (indices || n_src * n_batch == 0), "csr_encode: NULL output");
    const int64_t n = n_src * n_batch;
    if (n > 0) BE_CUDA(cudaMemsetAsync(indices, 0, (size_t)n * 4, s));
    Events ev;
    if (int rc = flat_ids(spikes, spk_kind, n, workspace, workspace_bytes, s, &ev)) return rc;
    minor_and_ptr_kernel<<<grid_for(std::max(n, n_src + 1)), 256, 0, s>>>(ev.active_ids, ev.n_active, n_src,
                                                                         (int)std::max<int64_t>(n_batch, 1), indices, indptr);
    BE_LAUNCH_CHECK("minor_and_ptr_kernel");
    return BE_B200_OK;
}

extern "C" int be_b200_binary_2d_csc_encode(const void* spikes, int spk_kind, int64_t n_src, int64_t n_batch,
                                            int32_t* indices, int32_t* indptr, void* scratch, void* workspace,
                                            size_t workspace_bytes, void* stream) { BE_NVTX_RANGE("be_b200_binary_2d_csc_encode");
    cudaStream_t s = (cudaStream_t)stream;
    if (int rc = check_2d(spikes, spk_kind, n_src, n_batch, "csc_encode")) return rc;
    BE_REQUIRE(indptr && (indices || n_src * n_batch == 0), "csc_encode: NULL output");
    const int64_t n = n_src * n_batch;
    BE_REQUIRE(scratch || n == 0, "csc_encode: scratch (n_src*n_batch bytes) is NULL");
    if (n > 0) {
        BE_CUDA(cudaMemsetAsync(indices, 0, (size_t)n * 4, s));
        dim3 grid((unsigned)ceil_div(n_src, 32), (unsigned)ceil_div(n_batch, 32)), block(32, 8);
        BE_REQUIRE(grid.y <= 65535u, "csc_encode: n_batch > 2,097,120 is not supported");
        if (spk_kind == BE_B200_SPK_BOOL) transpose_mask_kernel<BE_B200_SPK_BOOL><<<grid, block, 0, s>>>(spikes, n_src, n_batch, (uint8_t*)scratch);
        else transpose_mask_kernel<BE_B200_SPK_F32><<<grid, block, 0, s>>>(spikes, n_src, n_batch, (uint8_t*)scratch);
        BE_LAUNCH_CHECK("transpose_mask_kernel");
    }
    Events ev;
    if (int rc = flat_ids(scratch, BE_B200_SPK_BOOL, n, workspace, workspace_bytes, s, &ev)) return rc;
    minor_and_ptr_kernel<<<grid_for(std::max(n, n_batch + 1)), 256, 0, s>>>(ev.active_ids, ev.n_active, n_batch,
                                                                           (int)std::max<int64_t>(n_src, 1), indices, indptr);
    BE_LAUNCH_CHECK("minor_and_ptr_kernel");
    return BE_B200_OK;
}

extern "C" int be_b200_binary_2d_row_sparse_encode(const void* spikes, int spk_kind, int64_t n_src, int64_t n_batch,
                                                   int64_t row_size, int32_t* spike_indices, int32_t* max_row_nnz,
                                                   void* stream) { BE_NVTX_RANGE("be_b200_binary_2d_row_sparse_encode");
    cudaStream_t s = (cudaStream_t)stream;
    if (int rc = check_2d(spikes, spk_kind, n_src, n_batch, "row_sparse_encode")) return rc;
    BE_REQUIRE(row_size > 0 && row_size <= std::max<int64_t>(n_batch, 1) && row_size < (1ll << 31), "row_sparse_encode: row_size must be in [1, n_batch]");
    BE_REQUIRE(max_row_nnz && (spike_indices || n_src == 0), "row_sparse_encode: NULL output");
    BE_CUDA(cudaMemsetAsync(max_row_nnz, 0, 4, s));
    if (n_src == 0) return BE_B200_OK;
    const int blocks = (int)std::min<int64_t>(ceil_div(n_src, 8), (int64_t)kNumSM * 16);
    if (spk_kind == BE_B200_SPK_BOOL) row_sparse_kernel<BE_B200_SPK_BOOL><<<blocks, 256, 0, s>>>(spikes, n_src, (int)n_batch, (int)row_size, spike_indices, max_row_nnz);
    else row_sparse_kernel<BE_B200_SPK_F32><<<blocks, 256, 0, s>>>(spikes, n_src, (int)n_batch, (int)row_size, spike_indices, max_row_nnz);
    BE_LAUNCH_CHECK("row_sparse_kernel");
    return BE_B200_OK;
}
