// jitc_csr.cu -- materialise a just-in-time connectivity matrix as CSR with the same walk the
// product kernels use.
//
// Replaces ref: _jit_scalar/csr.{py,cu} (jits_csr_count / jits_csr_fill / jits_to_csr :488-570),
// _jit_normal/csr.{py,cu}, _jit_uniform/csr.{py,cu}.
//
// Three device passes, no host round trip inside a pass and no host-side sort (the reference
// reads the filled arrays back and lexsorts them on the host):
//   count : one warp per (generator row, chunk) walks the connections and accumulates per-CSR-row
//           counts; an exclusive scan turns them into int64 indptr (indptr[n_rows] = nnz).
//   fill  : the same walk appends each column id to its row's slice through a per-row cursor
//           (warp-aggregated when the whole warp feeds one row), then CUB's segmented sort puts
//           every row in ascending column order -- the canonical, deterministic CSR the reference
//           returns.  Values are computed AFTER the sort from (row, column) alone: the weight
//           is a pure hash of the pair, so nothing but 4-byte keys goes through the sort.
// corder=1: generator row == CSR row, walk over n_cols.  corder=0: generator rows are the n_cols
// columns and the walk runs over the n_rows CSR rows.  chunk_size derives from n_cols either way.
#include "jitc.cuh"

#include <algorithm>

#include <cub/cub.cuh>

namespace be {

template <int S, bool CORDER, bool FILL>
__global__ void __launch_bounds__(256) jit_csr_walk_kernel(const int32_t* __restrict__ clen_p,
                                                           const int32_t* __restrict__ seed_p,
                                                           int64_t n_gen, int64_t walk_len, int64_t chunk_size,
                                                           int32_t* __restrict__ counts,
                                                           const int64_t* __restrict__ indptr,
                                                           int32_t* __restrict__ indices) {
    uint32_t cl = (uint32_t)__ldg(clen_p);
    if (cl < 2U) cl = 2U;
    const uint32_t seed = (uint32_t)__ldg(seed_p);
    const uint32_t lane = threadIdx.x & 31;
    const int n_chunks = (int)ceil_div(walk_len, chunk_size);
    const int64_t n_tasks = n_gen * n_chunks;
    for (int64_t t = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); t < n_tasks; t += (int64_t)gridDim.x * 8) {
        const int64_t r = t / n_chunks;
        const int c = (int)(t - r * n_chunks);
        const int64_t cs = (int64_t)c * chunk_size;
        const uint32_t width = (uint32_t)min(chunk_size, walk_len - cs);
        uint32_t s = 0, q = 0, qmax = 0;
        if (lane < (uint32_t)S) {
            qmax = lane_qmax<S>(width, lane);
            s = rng_init(seed, (uint32_t)r, (uint32_t)c, lane);
            q = rng_initial_q(s, cl);
        }
        if (CORDER) {
            // every connection of this warp lands in CSR row r: one cursor bump per warp round
            if (!FILL) {
                uint32_t n = 0;
                while (q < qmax) { ++n; s = rng_next(s); q += 1U + __umulhi(s, cl - 1U); }
                n = __reduce_add_sync(0xffffffffu, n);
                if (lane == 0 && n) atomicAdd(counts + r, (int32_t)n);
            } else {
                const int64_t base = __ldg(indptr + r);
                uint32_t live = __ballot_sync(0xffffffffu, q < qmax);
                while (live) {
                    int off = 0;
                    if (lane == (uint32_t)(__ffs(live) - 1)) off = atomicAdd(counts + r, __popc(live));
                    off = __shfl_sync(0xffffffffu, off, __ffs(live) - 1);
                    if (q < qmax) {
                        indices[base + off + __popc(live & ((1u << lane) - 1u))] = (int32_t)(cs + lane + (int64_t)S * q);
                        s = rng_next(s);
                        q += 1U + __umulhi(s, cl - 1U);
                    }
                    live = __ballot_sync(0xffffffffu, q < qmax);
                }
            }
        } else {
            while (q < qmax) {
                const int64_t j = cs + lane + (int64_t)S * q;     // CSR row; the generator row is the column
                const int off = atomicAdd(counts + j, 1);
                if (FILL) indices[__ldg(indptr + j) + off] = (int32_t)r;
                s = rng_next(s);
                q += 1U + __umulhi(s, cl - 1U);
            }
        }
    }
}

__global__ void counts_to_i64_kernel(const int32_t* __restrict__ counts, int64_t n, int64_t* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = i < n ? (int64_t)counts[i] : 0;
}

// data[pos] = w(generator row, walk position) for the sorted structure; one warp per CSR row
template <typename T, int KIND, bool CORDER>
__global__ void __launch_bounds__(256) jit_csr_values_kernel(JitParams jp, const int64_t* __restrict__ indptr,
                                                             const int32_t* __restrict__ indices,
                                                             int64_t n_rows, T* __restrict__ data) {
    using A = typename Acc<T>::type;
    A p0, p1; uint32_t cl, seed;
    load_params<T, KIND>(jp, p0, p1, cl, seed);
    const int lane = threadIdx.x & 31;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < n_rows; r += (int64_t)gridDim.x * 8) {
        const int64_t b = __ldg(indptr + r), e = __ldg(indptr + r + 1);
        for (int64_t p = b + lane; p < e; p += 32) {
            const uint32_t col = (uint32_t)__ldg(indices + p);
            data[p] = from_acc<T>(CORDER ? jit_weight<KIND, A>(p0, p1, seed, (uint32_t)r, col)
                                         : jit_weight<KIND, A>(p0, p1, seed, col, (uint32_t)r));
        }
    }
}

static size_t scan_temp_bytes(int64_t n) {
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, (const int64_t*)nullptr, (int64_t*)nullptr, n + 1);
    return tb;
}
static size_t segsort_temp_bytes(int64_t nnz, int64_t n_rows) {
    size_t tb = 0;
    cub::DeviceSegmentedSort::SortKeys(nullptr, tb, (const int32_t*)nullptr, (int32_t*)nullptr, nnz, n_rows,
                                       (const int64_t*)nullptr, (const int64_t*)nullptr);
    return tb;
}

template <int S>
static void launch_walk(bool corder, bool fill, int blocks, cudaStream_t s, const int32_t* clen, const int32_t* seed,
                        int64_t n_gen, int64_t walk_len, int64_t chunk, int32_t* counts, const int64_t* indptr,
                        int32_t* indices) {
    if (corder) {
        if (fill) jit_csr_walk_kernel<S, true, true><<<blocks, 256, 0, s>>>(clen, seed, n_gen, walk_len, chunk, counts, indptr, indices);
        else jit_csr_walk_kernel<S, true, false><<<blocks, 256, 0, s>>>(clen, seed, n_gen, walk_len, chunk, counts, indptr, indices);
    } else {
        if (fill) jit_csr_walk_kernel<S, false, true><<<blocks, 256, 0, s>>>(clen, seed, n_gen, walk_len, chunk, counts, indptr, indices);
        else jit_csr_walk_kernel<S, false, false><<<blocks, 256, 0, s>>>(clen, seed, n_gen, walk_len, chunk, counts, indptr, indices);
    }
}

static int walk(bool corder, bool fill, int stride, const int32_t* clen, const int32_t* seed, int64_t n_rows,
                int64_t n_cols, int64_t chunk, int32_t* counts, const int64_t* indptr, int32_t* indices,
                cudaStream_t s) {
    const int64_t n_gen = corder ? n_rows : n_cols;
    const int64_t walk_len = corder ? n_cols : n_rows;
    const int64_t tasks = n_gen * ceil_div(walk_len, chunk);
    const int blocks = (int)std::min<int64_t>(ceil_div(std::max<int64_t>(tasks, 1), 8), (int64_t)kNumSM * 16);
    if (stride == 32) launch_walk<32>(corder, fill, blocks, s, clen, seed, n_gen, walk_len, chunk, counts, indptr, indices);
    else launch_walk<4>(corder, fill, blocks, s, clen, seed, n_gen, walk_len, chunk, counts, indptr, indices);
    BE_LAUNCH_CHECK("jit_csr_walk_kernel");
    return BE_B200_OK;
}

}  // namespace be

using namespace be;

extern "C" size_t be_b200_jitc_csr_workspace_bytes(int64_t n_rows, int64_t nnz) {
    if (n_rows < 0) n_rows = 0;
    if (nnz < 0) nnz = 0;
    const size_t cnt = align_up((size_t)(n_rows + 1) * 4, 256);
    const size_t i64 = align_up((size_t)(n_rows + 1) * 8, 256);
    const size_t count_phase = cnt + i64 + align_up(scan_temp_bytes(n_rows), 256);
    const size_t fill_phase = cnt + align_up((size_t)nnz * 4, 256) + align_up(segsort_temp_bytes(nnz, n_rows), 256);
    return std::max(count_phase, fill_phase) + 256;
}

extern "C" int be_b200_jitc_csr_count(const int32_t* clen, const int32_t* seed, int64_t n_rows,
                                      int64_t n_cols, int64_t chunk_size, int stride, int corder,
                                      int64_t* indptr, void* workspace, size_t workspace_bytes,
                                      void* stream) { BE_NVTX_RANGE("be_b200_jitc_csr_count");
    cudaStream_t s = (cudaStream_t)stream;
    BE_REQUIRE(clen && seed && indptr && workspace, "jitc_csr_count: NULL pointer");
    BE_REQUIRE(stride == 32 || stride == 4, "jitc_csr_count: stride must be 32 (mv) or 4 (mm)");
    BE_REQUIRE(n_rows >= 0 && n_cols >= 0 && n_rows < (1ll << 31) && n_cols < (1ll << 31) && chunk_size >= 1, "jitc_csr_count: bad sizes");
    BE_REQUIRE(((uintptr_t)workspace & 255u) == 0, "jitc_csr_count: workspace must be 256-byte aligned");
    BE_REQUIRE(workspace_bytes >= be_b200_jitc_csr_workspace_bytes(n_rows, 0), "jitc_csr_count: workspace too small");
    char* p = (char*)workspace;
    int32_t* counts = (int32_t*)p; p += align_up((size_t)(n_rows + 1) * 4, 256);
    int64_t* wide = (int64_t*)p; p += align_up((size_t)(n_rows + 1) * 8, 256);
    BE_CUDA(cudaMemsetAsync(counts, 0, (size_t)(n_rows + 1) * 4, s));
    if (n_rows > 0 && n_cols > 0) {
        int rc = walk(corder != 0, false, stride, clen, seed, n_rows, n_cols, chunk_size, counts, nullptr, nullptr, s);
        if (rc) return rc;
    }
    counts_to_i64_kernel<<<(int)std::min<int64_t>(ceil_div(n_rows + 1, 256), 148 * 8), 256, 0, s>>>(counts, n_rows, wide);
    BE_LAUNCH_CHECK("counts_to_i64_kernel");
    size_t tb = scan_temp_bytes(n_rows);
    BE_CUDA(cub::DeviceScan::ExclusiveSum(p, tb, (const int64_t*)wide, indptr, n_rows + 1, s));
    return BE_B200_OK;
}

extern "C" int be_b200_jitc_csr_fill(int kind, const void* p0, const void* p1, int w_dtype,
                                     const int32_t* clen, const int32_t* seed, int64_t n_rows,
                                     int64_t n_cols, int64_t chunk_size, int stride, int corder,
                                     const int64_t* indptr, int64_t nnz, int32_t* indices, void* data,
                                     void* workspace, size_t workspace_bytes, void* stream) { BE_NVTX_RANGE("be_b200_jitc_csr_fill");
    cudaStream_t s = (cudaStream_t)stream;
    BE_REQUIRE(p0 && clen && seed && indptr && workspace, "jitc_csr_fill: NULL pointer");
    BE_REQUIRE(kind == BE_B200_JIT_SCALAR || p1, "jitc_csr_fill: p1 is NULL");
    BE_REQUIRE(stride == 32 || stride == 4, "jitc_csr_fill: stride must be 32 (mv) or 4 (mm)");
    BE_REQUIRE(n_rows >= 0 && n_cols >= 0 && n_rows < (1ll << 31) && n_cols < (1ll << 31) && chunk_size >= 1 && nnz >= 0, "jitc_csr_fill: bad sizes");
    if (nnz == 0) return BE_B200_OK;
    BE_REQUIRE(indices && data, "jitc_csr_fill: NULL output");
    BE_REQUIRE(((uintptr_t)workspace & 255u) == 0, "jitc_csr_fill: workspace must be 256-byte aligned");
    BE_REQUIRE(workspace_bytes >= be_b200_jitc_csr_workspace_bytes(n_rows, nnz), "jitc_csr_fill: workspace too small");
    char* p = (char*)workspace;
    int32_t* cursor = (int32_t*)p; p += align_up((size_t)(n_rows + 1) * 4, 256);
    int32_t* unsorted = (int32_t*)p; p += align_up((size_t)nnz * 4, 256);
    BE_CUDA(cudaMemsetAsync(cursor, 0, (size_t)(n_rows + 1) * 4, s));
    int rc = walk(corder != 0, true, stride, clen, seed, n_rows, n_cols, chunk_size, cursor, indptr, unsorted, s);
    if (rc) return rc;
    size_t tb = segsort_temp_bytes(nnz, n_rows);
    BE_CUDA(cub::DeviceSegmentedSort::SortKeys(p, tb, (const int32_t*)unsorted, indices, nnz, n_rows, indptr, indptr + 1, s));
    JitParams jp{p0, p1, clen, seed};
    const int blocks = (int)std::min<int64_t>(ceil_div(n_rows, 8), (int64_t)kNumSM * 16);
    BE_DISPATCH_WT(w_dtype, T, {
        BE_DISPATCH_KIND(kind, K, {
            if (corder) jit_csr_values_kernel<T, K, true><<<blocks, 256, 0, s>>>(jp, indptr, indices, n_rows, (T*)data);
            else jit_csr_values_kernel<T, K, false><<<blocks, 256, 0, s>>>(jp, indptr, indices, n_rows, (T*)data);
        });
    });
    BE_LAUNCH_CHECK("jit_csr_values_kernel");
    return BE_B200_OK;
}
