// csc.cu -- CSC mirror of a CSR structure (setup time, not on the per-step path).
//
// Replaces ref: _misc.py:1341-1384 (_csr_to_csc_index_numpy) and _csr/csr_to_csc.cu.
// perm = stable argsort of the column indices; csc_indices[p] = row owning slot perm[p];
// csc_indptr[c] = number of slots with column < c.  The stable sort is CUB's LSD radix sort
// (library code, used here only because this runs once per matrix, never per timestep).
#include "common.cuh"

#include <cub/cub.cuh>

namespace be {

template <typename OT>
__global__ void iota_kernel(OT* v, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] = (OT)i;
}

// csc_indptr[c] = lower_bound(sorted_cols, c)
template <typename OT>
__global__ void col_ptr_kernel(const int32_t* __restrict__ sorted_cols, int64_t nse, int64_t n_post,
                               OT* __restrict__ csc_indptr) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c <= n_post; c += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = nse;
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if ((int64_t)sorted_cols[mid] < c) lo = mid + 1; else hi = mid;
        }
        csc_indptr[c] = (OT)lo;
    }
}

// csc_indices[p] = (upper_bound(indptr, perm[p]) - 1)
template <typename PT, typename OT>
__global__ void row_of_slot_kernel(const PT* __restrict__ indptr, int64_t m, const OT* __restrict__ perm,
                                   int64_t nse, int32_t* __restrict__ csc_indices) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nse; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = (int64_t)perm[p];
        int64_t lo = 0, hi = m + 1;
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if ((int64_t)indptr[mid] <= j) lo = mid + 1; else hi = mid;
        }
        csc_indices[p] = (int32_t)(lo - 1);
    }
}

template <typename OT>
static size_t sort_temp_bytes(int64_t nse, int end_bit) {
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, (const int32_t*)nullptr, (int32_t*)nullptr,
                                    (const OT*)nullptr, (OT*)nullptr, nse, 0, end_bit);
    return tb;
}

static int bits_for(int64_t n_post) {
    int b = 1;
    while (b < 32 && (1ll << b) < n_post) ++b;
    return b;
}

template <typename PT, typename OT>
static int run_csc(const PT* indptr, const int32_t* indices, int64_t m, int64_t n_post, int64_t nse,
                   OT* csc_indptr, int32_t* csc_indices, OT* perm, void* workspace,
                   size_t workspace_bytes, cudaStream_t s) {
    const int end_bit = bits_for(n_post);
    char* p = (char*)workspace;
    int32_t* sorted_cols = (int32_t*)p; p += align_up((size_t)nse * 4, 256);
    OT* iota = (OT*)p; p += align_up((size_t)nse * sizeof(OT), 256);
    size_t tb = sort_temp_bytes<OT>(nse, end_bit);
    BE_REQUIRE((size_t)(p - (char*)workspace) + tb <= workspace_bytes, "csr_to_csc: workspace too small");
    const int blocks = (int)std::min<int64_t>(ceil_div(std::max<int64_t>(nse, n_post + 1), 256), 148 * 16);
    if (nse > 0) {
        iota_kernel<OT><<<blocks, 256, 0, s>>>(iota, nse);
        BE_LAUNCH_CHECK("iota_kernel");
        BE_CUDA(cub::DeviceRadixSort::SortPairs(p, tb, indices, sorted_cols, (const OT*)iota, perm, nse, 0, end_bit, s));
        row_of_slot_kernel<PT, OT><<<blocks, 256, 0, s>>>(indptr, m, perm, nse, csc_indices);
        BE_LAUNCH_CHECK("row_of_slot_kernel");
    }
    col_ptr_kernel<OT><<<blocks, 256, 0, s>>>(sorted_cols, nse, n_post, csc_indptr);
    BE_LAUNCH_CHECK("col_ptr_kernel");
    return BE_B200_OK;
}

}  // namespace be

using namespace be;

extern "C" size_t be_b200_csc_workspace_bytes(int64_t m, int64_t n_post, int64_t nse) {
    (void)m;
    if (nse < 0) nse = 0;
    size_t tb = std::max(sort_temp_bytes<int64_t>(nse, 32), sort_temp_bytes<int32_t>(nse, 32));
    return align_up((size_t)nse * 4, 256) + align_up((size_t)nse * 8, 256) + align_up(tb, 256) + 256;
}

extern "C" int be_b200_csr_to_csc(const void* indptr, int is64, const int32_t* indices, int64_t m,
                                  int64_t n_post, int64_t nse, void* csc_indptr,
                                  int32_t* csc_indices, void* perm, void* workspace,
                                  size_t workspace_bytes, void* stream) { BE_NVTX_RANGE("be_b200_csr_to_csc");
    cudaStream_t s = (cudaStream_t)stream;
    BE_REQUIRE(m >= 0 && n_post >= 0 && nse >= 0, "csr_to_csc: negative size");
    BE_REQUIRE(indptr && csc_indptr && workspace, "csr_to_csc: NULL pointer");
    BE_REQUIRE(nse == 0 || (indices && csc_indices && perm), "csr_to_csc: NULL pointer");
    BE_REQUIRE(((uintptr_t)workspace & 255u) == 0, "csr_to_csc: workspace must be 256-byte aligned");
    if (is64) return run_csc<int64_t, int64_t>((const int64_t*)indptr, indices, m, n_post, nse, (int64_t*)csc_indptr, csc_indices, (int64_t*)perm, workspace, workspace_bytes, s);
    return run_csc<int32_t, int32_t>((const int32_t*)indptr, indices, m, n_post, nse, (int32_t*)csc_indptr, csc_indices, (int32_t*)perm, workspace, workspace_bytes, s);
}
