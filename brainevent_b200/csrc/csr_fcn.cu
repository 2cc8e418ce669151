// csr_fcn.cu -- event-driven propagation through stored connectivity (CSR/CSC and fixed-number/ELL).
//
// Replaces ref: _csr/binary_csrmv.cu (NT gather), _csr/binary_csrmv_hybrid.cu (T scatter),
// _csr/binary_csrmm{,_hybrid}.cu, _csr/binary_indexed_csrm{v,m}{,_hybrid}.cu,
// _fcn/binary_fcnmv.cu, _fcn/binary_fcnmm.cu.
//
// Design (B200-first, not a port):
//  * scatter ("transpose"): driven by the ordered active-row list from events.cu, so only active
//    rows are touched; a group of lanes streams a row's indices/weights with 128-bit
//    L1-bypassing loads and issues fire-and-forget RED.ADD to the L2-resident output.  Rows
//    longer than kHeavyRow are deferred and processed by the whole CTA.
//  * gather ("no transpose"): deterministic, no atomics; every output row tests its sources
//    against the bit-packed spike vector (8x smaller than a byte mask, L1/L2 resident) and loads
//    a 16-byte weight quad only when one of its four sources fired.
//  * mm: the connectivity row is read ONCE per call and applied to every batch column
//    (the reference loops over columns on the host); outputs are [rows, batch] with
//    batch innermost so one synapse touches one contiguous segment.
#include "conn.cuh"

#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace be {

constexpr int kHeavyRow = 8192;   // rows longer than this are processed by the whole CTA
constexpr int kHeavyCap = 32;     // deferred heavy rows per CTA before falling back to the warp

// ------------------------------------------------------------------------------------------
// scatter, matvec:  out[indices[j]] += w[j]  for j in rows of active sources
// ------------------------------------------------------------------------------------------
template <typename T, bool HOMO, int PERM, int VEC, int G>
__device__ __forceinline__ void scatter_segment(int64_t s, int64_t e, int gl,
                                                const int32_t* __restrict__ indices,
                                                const T* __restrict__ weights,
                                                const void* __restrict__ perm,
                                                typename Acc<T>::type w0, T* __restrict__ out) {
    using A = typename Acc<T>::type;
    for_each_synapse<T, HOMO, PERM, VEC, G>(s, e, gl, indices, weights, perm, w0,
                                            [&](int col, A w) { red_add(out + col, w); });
}

template <typename T, bool HOMO, int PERM, int VEC, int LPR, typename Rows>
__global__ void __launch_bounds__(256) scatter_mv_kernel(Rows rows, const int32_t* __restrict__ indices,
                                                         const T* __restrict__ weights,
                                                         const void* __restrict__ perm,
                                                         const int32_t* __restrict__ ids,
                                                         const int* __restrict__ n_active_ptr,
                                                         T* __restrict__ out, int sel_thr) {
    using A = typename Acc<T>::type;
    if (sel_thr >= 0 && *n_active_ptr > sel_thr) return;   // the gather variant owns this call
    __shared__ int s_heavy[kHeavyCap];
    __shared__ int s_nheavy;
    if (threadIdx.x == 0) s_nheavy = 0;
    __syncthreads();
    const int n_active = *n_active_ptr;
    const A w0 = HOMO ? to_acc<T>(__ldg(weights)) : A(0);
    constexpr int GROUPS = 256 / LPR;
    const int g = threadIdx.x / LPR, gl = threadIdx.x % LPR;
    const int64_t stride = (int64_t)gridDim.x * GROUPS;
    int64_t a = (int64_t)blockIdx.x * GROUPS + g;
    int r_n = 0;
    int64_t s_n = 0, e_n = 0;
    if (a < n_active) { r_n = __ldg(ids + a); rows.get(r_n, s_n, e_n); }
    for (; a < n_active; a += stride) {
        const int r = r_n;
        const int64_t s = s_n, e = e_n;
        if (a + stride < n_active) { r_n = __ldg(ids + a + stride); rows.get(r_n, s_n, e_n); }   // next row's chain in flight
        if (e - s > kHeavyRow) {
            int slot = -1;
            if (gl == 0) slot = atomicAdd(&s_nheavy, 1);
            const int leader = ((threadIdx.x & 31) / LPR) * LPR;
            const unsigned gmask = (LPR >= 32) ? 0xffffffffu : (((1u << LPR) - 1u) << leader);
            slot = __shfl_sync(gmask, slot, leader);
            if (slot < kHeavyCap) { if (gl == 0) s_heavy[slot] = r; continue; }
        }
        scatter_segment<T, HOMO, PERM, VEC, LPR>(s, e, gl, indices, weights, perm, w0, out);
    }
    __syncthreads();
    const int nh = min(s_nheavy, kHeavyCap);
    for (int h = 0; h < nh; ++h) {
        int64_t s, e;
        rows.get(s_heavy[h], s, e);
        scatter_segment<T, HOMO, PERM, VEC, 256>(s, e, threadIdx.x, indices, weights, perm, w0, out);
    }
}

// ------------------------------------------------------------------------------------------
// scatter, matmat: out[indices[j], c] += w[j] for every active column c of an active row.
// One warp per active row; the row's connectivity is read once.  Sparse column masks use
// lanes = synapses (32 REDs per instruction); dense masks use lanes = columns (contiguous REDs).
// ------------------------------------------------------------------------------------------

template <typename T, bool HOMO, int PERM, int VEC, int MW, typename Rows>
__global__ void __launch_bounds__(256) scatter_mm_kernel(Rows rows, const int32_t* __restrict__ indices,
                                                         const T* __restrict__ weights,
                                                         const void* __restrict__ perm,
                                                         const int32_t* __restrict__ ids,
                                                         const int* __restrict__ n_active_ptr,
                                                         const uint32_t* __restrict__ packed, int wpr,
                                                         int64_t n, T* __restrict__ out,
                                                         int post_lo, int post_hi) {
    // [post_lo, post_hi): the output rows this launch owns.  When [n_out, batch] does not fit in
    // L2 the host splits the call into passes over L2-sized output tiles so that the reductions
    // stay L2-resident instead of becoming DRAM read-modify-writes (the matrix rows are re-read
    // per pass, which is the cheaper side: 8 B/synapse streamed vs 64 B/event of DRAM RMW).
    using A = typename Acc<T>::type;
    const int n_active = *n_active_ptr;
    const A w0 = HOMO ? to_acc<T>(__ldg(weights)) : A(0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // the (row id -> row extent) chain of the NEXT row is fetched while the current row streams,
    // so a row visit costs one memory round trip instead of three dependent ones
    const int64_t stride = (int64_t)gridDim.x * 8;
    int64_t a = (int64_t)blockIdx.x * 8 + warp;
    int r_n = 0;
    int64_t s_n = 0, e_n = 0;
    if (a < n_active) { r_n = __ldg(ids + a); rows.get(r_n, s_n, e_n); }
    for (; a < n_active; a += stride) {
        const int r = r_n;
        const int64_t s = s_n, e = e_n;
        if (a + stride < n_active) { r_n = __ldg(ids + a + stride); rows.get(r_n, s_n, e_n); }
        const uint32_t* mrow = packed + (int64_t)r * wpr;
        for (int w_base = 0; w_base < wpr; w_base += MW) {
            uint32_t mask[MW];
            int pc = 0;
#pragma unroll
            for (int q = 0; q < MW; ++q) {
                mask[q] = (w_base + q < wpr) ? __ldg(mrow + w_base + q) : 0u;
                pc += __popc(mask[q]);
            }
            if (pc == 0) continue;
            if (pc <= 8) {
                // lanes = synapses
                for_each_synapse<T, HOMO, PERM, VEC, 32>(s, e, lane, indices, weights, perm, w0, [&](int col, A wv) {
                    if (col < post_lo || col >= post_hi) return;
                    T* o = out + (int64_t)col * n + (int64_t)w_base * 32;
#pragma unroll
                    for (int q = 0; q < MW; ++q) {
                        uint32_t m = mask[q];
                        while (m) {
                            const int c = __ffs(m) - 1;
                            m &= m - 1;
                            red_add(o + q * 32 + c, wv);
                        }
                    }
                });
            } else {
                // lanes = columns: broadcast one synapse at a time
                for (int64_t j0 = s; j0 < e; j0 += 32) {
                    const int64_t j = j0 + lane;
                    int64_t my_o = 0;
                    A my_w = A(0);
                    bool mine = false;
                    if (j < e) {
                        const int col = __ldg(indices + j);
                        mine = col >= post_lo && col < post_hi;
                        if (mine) {
                            my_o = (int64_t)col * n + (int64_t)w_base * 32;
                            my_w = weight_at<T, HOMO, PERM>(weights, perm, j, w0);
                        }
                    }
                    uint32_t todo = __ballot_sync(0xffffffffu, mine);
                    while (todo) {
                        const int t = __ffs(todo) - 1;
                        todo &= todo - 1;
                        const int64_t o = __shfl_sync(0xffffffffu, my_o, t);
                        const A wv = __shfl_sync(0xffffffffu, my_w, t);
#pragma unroll
                        for (int q = 0; q < MW; ++q)
                            if ((mask[q] >> lane) & 1u) red_add(out + o + q * 32 + lane, wv);
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// scatter, matmat, BATCH-MAJOR output out[c, :] (physical [batch, n_out]): every (active row,
// active column) pair is one matvec-style row scatter into that column's contiguous output
// vector.  The host walks the columns in groups whose output vectors fit in L2 together
// ([col_lo, col_hi) per launch): reductions stay L2-resident, a row is re-read only for the
// groups it spikes in (~1.3x at 1 % rates instead of once per output tile), and no synapse is
// filtered away after being loaded.
// ------------------------------------------------------------------------------------------
template <typename T, bool HOMO, int PERM, int VEC, typename Rows>
__global__ void __launch_bounds__(256) scatter_mm_bm_kernel(Rows rows, const int32_t* __restrict__ indices,
                                                            const T* __restrict__ weights,
                                                            const void* __restrict__ perm,
                                                            const int32_t* __restrict__ ids,
                                                            const int* __restrict__ n_active_ptr,
                                                            const uint32_t* __restrict__ packed, int wpr,
                                                            int64_t n_out, T* __restrict__ out,
                                                            int col_lo, int col_hi) {
    using A = typename Acc<T>::type;
    const int n_active = *n_active_ptr;
    const A w0 = HOMO ? to_acc<T>(__ldg(weights)) : A(0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w_lo = col_lo >> 5, w_hi = (col_hi + 31) >> 5;
    auto group_mask = [&](const uint32_t* mrow, int w) -> uint32_t {     // bits of word w that belong to this column group
        uint32_t m = __ldg(mrow + w);
        const int b0 = w * 32;
        if (col_lo > b0) m &= ~0u << (col_lo - b0);
        if (col_hi < b0 + 32) m &= (1u << (col_hi - b0)) - 1u;
        return m;
    };
    // The rows active in ANY column are ~50 % of all rows at batch 64 / 1 %, but only ~11 % of them fire inside one
    // column group: the group test is therefore done by LANE (32 candidate rows per warp trip: id, mask words and
    // row extent in one round of independent loads) and the warp then streams only the rows that have work.
    // (One row per warp trip made every silent row cost two dependent memory round trips: 2.8 M tests per call.)
    const int64_t stride = (int64_t)gridDim.x * 8 * 32;
    for (int64_t a0 = ((int64_t)blockIdx.x * 8 + warp) * 32; a0 < n_active; a0 += stride) {
        int r = -1;
        bool work = false;
        int64_t s = 0, e = 0;
        if (a0 + lane < n_active) {
            r = __ldg(ids + a0 + lane);
            const uint32_t* mrow = packed + (int64_t)r * wpr;
            uint32_t any = 0;
            for (int w = w_lo; w < w_hi; ++w) any |= group_mask(mrow, w);
            work = any != 0;
            if (work) rows.get(r, s, e);
        }
        uint32_t todo = __ballot_sync(0xffffffffu, work);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int rr = __shfl_sync(0xffffffffu, r, src);
            const int64_t ss = __shfl_sync(0xffffffffu, s, src), ee = __shfl_sync(0xffffffffu, e, src);
            const uint32_t* mrow = packed + (int64_t)rr * wpr;
            for (int w = w_lo; w < w_hi; ++w) {
                uint32_t m = group_mask(mrow, w);                 // warp-uniform, L1 hit
                while (m) {
                    const int c = w * 32 + __ffs(m) - 1;
                    m &= m - 1;
                    scatter_segment<T, HOMO, PERM, VEC, 32>(ss, ee, lane, indices, weights, perm, w0, out + (int64_t)c * n_out);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------
template <typename T, bool HOMO, int PERM, typename Rows>
static int launch_scatter_mv(const Rows& rows, const ConnArgs& c, const Events& ev, int64_t n_src,
                             T* out, cudaStream_t s) {
    const int vec = vec_mode(rows, c);
    const bool narrow = c.avg_row < 40.0;
    const int groups = narrow ? 32 : 8;
    int blocks = (int)std::min<int64_t>(ceil_div(std::max<int64_t>(n_src, 1), groups), kNumSM * 8);
#define BE_SC(VEC, LPR)                                                                      \
    scatter_mv_kernel<T, HOMO, PERM, VEC, LPR, Rows><<<blocks, 256, 0, s>>>(                 \
        rows, c.indices, (const T*)c.weights, c.perm, ev.active_ids, ev.n_active, out, ev.sel_thr)
    {
        ProfileScope prof(s);
        if (vec == 2) { if (narrow) BE_SC(2, 8); else BE_SC(2, 32); }
        else if (vec == 1) { if (narrow) BE_SC(1, 8); else BE_SC(1, 32); }
        else { if (narrow) BE_SC(0, 8); else BE_SC(0, 32); }
    }
#undef BE_SC
    BE_LAUNCH_CHECK("scatter_mv_kernel");
    return BE_B200_OK;
}


// ------------------------------------------------------------------------------------------
// scatter, matvec, SMALL networks: ONE launch of ONE thread-block cluster (8 CTAs x 1024 threads).
// A few-thousand-neuron projection (the reference's own examples: COBA / CUBA at 4 000 neurons) spends its time in
// launch latency -- compaction kernel + scatter kernel, ~5 us for a few thousand reductions.  Here every CTA clears
// its share of the output and compacts its share of the spike vector into shared memory, the cluster barrier
// (hardware, release / acquire) orders the clear before any reduction, and every CTA then scatters the rows it found.
// No workspace, no global id list, no second launch.
// ------------------------------------------------------------------------------------------
constexpr int kSmallCtas = 8, kSmallThreads = 1024;
constexpr int64_t kSmallMaxSrc = (int64_t)kSmallCtas * 4096;     // sources (4096 list entries per CTA)
constexpr int64_t kSmallMaxOut = 1 << 18;                        // outputs cleared inside the kernel

template <typename T, bool HOMO, int PERM, int VEC, int KIND, typename Rows>
__global__ void __cluster_dims__(kSmallCtas, 1, 1) __launch_bounds__(kSmallThreads)
scatter_mv_small_kernel(Rows rows, const int32_t* __restrict__ indices, const T* __restrict__ weights,
                        const void* __restrict__ perm, const void* __restrict__ spikes, int64_t n_src,
                        T* __restrict__ out, int64_t n_out) {
    using A = typename Acc<T>::type;
    __shared__ int s_list[4096];
    __shared__ int s_n;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const A w0 = HOMO ? to_acc<T>(__ldg(weights)) : A(0);
    if (tid == 0) s_n = 0;
    for (int64_t i = (int64_t)blockIdx.x * kSmallThreads + tid; i < n_out; i += (int64_t)kSmallCtas * kSmallThreads) out[i] = T(0);
    __syncthreads();
    const int64_t per = (n_src + kSmallCtas - 1) / kSmallCtas;
    const int64_t lo = (int64_t)blockIdx.x * per, hi = min(n_src, lo + per);
    for (int64_t base = lo; base < hi; base += kSmallThreads) {
        const int64_t i = base + tid;
        bool on = false;
        if (i < hi) {
            if (KIND == BE_B200_SPK_BOOL) on = __ldg((const uint8_t*)spikes + i) != 0;
            else if (KIND == BE_B200_SPK_F32) on = __ldg((const float*)spikes + i) > 0.0f;
            else on = (__ldg((const uint32_t*)spikes + (i >> 5)) >> (i & 31)) & 1u;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, on);
        int pos = 0;
        if (lane == 0 && bal) pos = atomicAdd(&s_n, __popc(bal));
        pos = __shfl_sync(0xffffffffu, pos, 0);
        if (on) s_list[pos + __popc(bal & ((1u << lane) - 1u))] = (int)i;
    }
    __threadfence();        // the cleared outputs are visible device-wide before any CTA of the cluster reduces into them
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const int n = s_n;
    for (int a = warp; a < n; a += kSmallThreads / 32) {
        int64_t s, e;
        rows.get(s_list[a], s, e);
        scatter_segment<T, HOMO, PERM, VEC, 32>(s, e, lane, indices, weights, perm, w0, out);
    }
}

template <typename T, bool HOMO, int PERM, typename Rows>
static int launch_scatter_mv_small(const Rows& rows, const ConnArgs& c, const void* spikes, int spk_kind, int64_t n_src,
                                   T* out, int64_t n_out, cudaStream_t s) {
    const int vec = vec_mode(rows, c);
#define BE_SS(VEC, KIND)                                                                                   \
    scatter_mv_small_kernel<T, HOMO, PERM, VEC, KIND, Rows><<<kSmallCtas, kSmallThreads, 0, s>>>(          \
        rows, c.indices, (const T*)c.weights, c.perm, spikes, n_src, out, n_out)
#define BE_SS_K(VEC)                                                                                       \
    do { if (spk_kind == BE_B200_SPK_BOOL) BE_SS(VEC, BE_B200_SPK_BOOL);                                   \
         else if (spk_kind == BE_B200_SPK_F32) BE_SS(VEC, BE_B200_SPK_F32); else BE_SS(VEC, BE_B200_SPK_PACKED); } while (0)
    {
        ProfileScope prof(s);
        if (vec == 2) BE_SS_K(2); else if (vec == 1) BE_SS_K(1); else BE_SS_K(0);
    }
#undef BE_SS_K
#undef BE_SS
    BE_LAUNCH_CHECK("scatter_mv_small_kernel");
    return BE_B200_OK;
}

constexpr size_t kL2TileBytesDefault = 48u << 20;   // output tile kept L2-resident per scatter pass (L2 = 126 MB)
static size_t l2_tile_bytes() {
    if (const char* e = getenv("BE_B200_SPMM_TILE_MB")) { const long v = atol(e); if (v >= 1 && v <= 4096) return (size_t)v << 20; }   // A/B switch
    return kL2TileBytesDefault;
}
#define kL2TileBytes (l2_tile_bytes())

template <typename T, bool HOMO, int PERM, typename Rows>
static int launch_scatter_mm(const Rows& rows, const ConnArgs& c, const Events& ev, int64_t n_src,
                             int64_t n_out, int64_t n, T* out, cudaStream_t s) {
    int blocks = (int)std::min<int64_t>(ceil_div(std::max<int64_t>(n_src, 1), 8), kNumSM * 8);
    const size_t row_bytes = (size_t)n * sizeof(T);
    const int64_t rows_per_pass = std::max<int64_t>(1, (int64_t)(kL2TileBytes / row_bytes));
    const bool vec = vec_mode(rows, c) != 0;
    ProfileScope prof(s);   // one record for all passes of the call
    for (int64_t lo = 0; lo < n_out; lo += rows_per_pass) {
        const int64_t hi = std::min(n_out, lo + rows_per_pass);
        // clear the tile right before its pass: the zeroed lines are still L2-resident when the
        // reductions arrive, so the tile reaches DRAM once (write-back), never as RMW traffic
        BE_CUDA(cudaMemsetAsync(out + lo * n, 0, (size_t)(hi - lo) * row_bytes, s));
#define BE_SMM(VEC, MW)                                                                         \
        scatter_mm_kernel<T, HOMO, PERM, VEC, MW, Rows><<<blocks, 256, 0, s>>>(                     \
            rows, c.indices, (const T*)c.weights, c.perm, ev.active_ids, ev.n_active, ev.packed,    \
            ev.words_per_row, n, out, (int)lo, (int)hi)
        if (ev.words_per_row <= 2) { if (vec) BE_SMM(1, 2); else BE_SMM(0, 2); }
        else { if (vec) BE_SMM(1, 8); else BE_SMM(0, 8); }
#undef BE_SMM
    }
    BE_LAUNCH_CHECK("scatter_mm_kernel");
    return BE_B200_OK;
}

template <typename T, bool HOMO, int PERM, typename Rows>
static int launch_scatter_mm_bm(const Rows& rows, const ConnArgs& c, const Events& ev, int64_t n_src,
                                int64_t n_out, int64_t n, T* out, cudaStream_t s) {
    int blocks = (int)std::min<int64_t>(ceil_div(std::max<int64_t>(n_src, 1), 8 * 32), kNumSM * 8);
    const size_t col_bytes = (size_t)n_out * sizeof(T);
    const int64_t cols_per_pass = std::max<int64_t>(1, (int64_t)(kL2TileBytes / std::max<size_t>(col_bytes, 1)));
    const int vec = vec_mode(rows, c);
    ProfileScope prof(s);
    for (int64_t lo = 0; lo < n; lo += cols_per_pass) {
        const int64_t hi = std::min(n, lo + cols_per_pass);
        BE_CUDA(cudaMemsetAsync(out + lo * n_out, 0, (size_t)(hi - lo) * col_bytes, s));
#define BE_SBM(VEC)                                                                              \
        scatter_mm_bm_kernel<T, HOMO, PERM, VEC, Rows><<<blocks, 256, 0, s>>>(                       \
            rows, c.indices, (const T*)c.weights, c.perm, ev.active_ids, ev.n_active, ev.packed,     \
            ev.words_per_row, n_out, out, (int)lo, (int)hi)
        if (vec == 2) BE_SBM(2); else if (vec == 1) BE_SBM(1); else BE_SBM(0);
#undef BE_SBM
    }
    BE_LAUNCH_CHECK("scatter_mm_bm_kernel");
    return BE_B200_OK;
}

static bool small_path_enabled() {            // BE_B200_SMALL_SCATTER=0: always the two-kernel path (A/B)
    const char* e = getenv("BE_B200_SMALL_SCATTER");
    return !(e && e[0] == '0');
}

// One entry for all four shapes.  n == 0 means matvec.
template <typename T, typename Rows>
static int run_conn(const Rows& rows, const RowsDesc& rd, const ConnArgs& c, const void* spikes, const void* aux,
                    int spk_kind, int64_t n_src, int64_t n_rows, int64_t n_out, int64_t n_batch,
                    bool scatter, void* out_v, void* workspace, size_t workspace_bytes,
                    cudaStream_t s, bool batch_major = false) {
    T* out = (T*)out_v;
    Events ev;
    int rc;
    const bool mm = n_batch > 0;
    const int64_t out_elems = n_out * (mm ? n_batch : 1);
    // small networks: one single-cluster launch does compaction, output clear and scatter (latency, not bandwidth)
    if (scatter && !mm && n_src > 0 && n_src <= kSmallMaxSrc && n_out > 0 && n_out <= kSmallMaxOut &&
        (spk_kind == BE_B200_SPK_BOOL || spk_kind == BE_B200_SPK_F32 || spk_kind == BE_B200_SPK_PACKED) &&
        (double)n_src * c.avg_row <= 4.0e6 && small_path_enabled()) {
#define BE_RUN_SMALL(HOMO, PERM) return launch_scatter_mv_small<T, HOMO, PERM, Rows>(rows, c, spikes, spk_kind, n_src, out, n_out, s)
        if (c.homo) BE_RUN_SMALL(true, 0);
        if constexpr (!std::is_same<Rows, FcnRows>::value) {
            if (c.perm_kind == 1) BE_RUN_SMALL(false, 1);
            if (c.perm_kind == 2) BE_RUN_SMALL(false, 2);
        }
        BE_RUN_SMALL(false, 0);
#undef BE_RUN_SMALL
    }
    ev.defer_pack_ok = !mm && !scatter;     // the matvec gather can pack small bool vectors itself
    if (mm) rc = prepare_events_2d(spikes, aux, spk_kind, n_src, n_batch, scatter, workspace, workspace_bytes, s, &ev);
    else rc = prepare_events_1d(spikes, aux, spk_kind, n_src, scatter, !scatter, workspace, workspace_bytes, s, &ev,
                                scatter ? (void*)out : nullptr, scatter ? (size_t)out_elems * sizeof(T) : 0,
                                /*ordered_ids=*/false);   // the output clear rides on the compaction kernel
    if (rc) return rc;
    if (out_elems == 0) return BE_B200_OK;
    if (n_src == 0) {
        if (!scatter) BE_CUDA(cudaMemsetAsync(out, 0, (size_t)out_elems * sizeof(T), s));
        return BE_B200_OK;
    }
#define BE_RUN(HOMO, PERM)                                                                              \
    do {                                                                                                \
        if (scatter && mm && batch_major) return launch_scatter_mm_bm<T, HOMO, PERM, Rows>(rows, c, ev, n_src, n_out, n_batch, out, s); \
        return mm ? launch_scatter_mm<T, HOMO, PERM, Rows>(rows, c, ev, n_src, n_out, n_batch, out, s)       \
                  : launch_scatter_mv<T, HOMO, PERM, Rows>(rows, c, ev, n_src, out, s);                      \
    } while (0)
    if (!scatter) return run_gather(rd, c, ev, n_src, n_rows, n_batch, out, s);
    if (c.homo) BE_RUN(true, 0);
    if constexpr (!std::is_same<Rows, FcnRows>::value) {
        if (c.perm_kind == 1) BE_RUN(false, 1);
        if (c.perm_kind == 2) BE_RUN(false, 2);
    }
    BE_RUN(false, 0);
#undef BE_RUN
}

}  // namespace be

using namespace be;

static int csr_entry(const void* weights, int w_dtype, int homo, const int32_t* indices,
                     const void* indptr, int indptr_is64, const void* perm, int perm_is64,
                     const void* spikes, const void* aux, int spk_kind, int64_t m, int64_t k,
                     int64_t nse, int64_t n, int transpose, void* out, void* workspace,
                     size_t workspace_bytes, void* stream, int out_batch_major = 0) {
    cudaStream_t s = (cudaStream_t)stream;
    BE_REQUIRE(!out_batch_major || (transpose && n > 0), "batch-major output is only defined for the scatter matmat");
    BE_REQUIRE(m >= 0 && k >= 0 && m < (1ll << 31) && k < (1ll << 31), "csr: shape out of range");
    BE_REQUIRE(weights && indptr && out, "csr: NULL pointer");
    BE_REQUIRE(indices || nse == 0, "csr: indices is NULL");
    ConnArgs c{weights, w_dtype, homo, indices, homo ? nullptr : perm,
               (homo || !perm) ? 0 : (perm_is64 ? 2 : 1), m > 0 ? (double)nse / (double)m : 0.0, nse};
    const int64_t n_src = transpose ? m : k, n_out = transpose ? k : m;
    BE_DISPATCH_WT(w_dtype, T, {
        if (indptr_is64) {
            CsrRows<int64_t> rows{(const int64_t*)indptr};
            return run_conn<T>(rows, RowsDesc{1, indptr, 0}, c, spikes, aux, spk_kind, n_src, m, n_out, n, transpose != 0, out, workspace, workspace_bytes, s, out_batch_major != 0);
        } else {
            CsrRows<int32_t> rows{(const int32_t*)indptr};
            return run_conn<T>(rows, RowsDesc{0, indptr, 0}, c, spikes, aux, spk_kind, n_src, m, n_out, n, transpose != 0, out, workspace, workspace_bytes, s, out_batch_major != 0);
        }
    });
    return BE_B200_OK;
}

extern "C" int be_b200_binary_csrmv(const void* weights, int w_dtype, int homo,
                                    const int32_t* indices, const void* indptr, int indptr_is64,
                                    const void* perm, int perm_is64, const void* spikes,
                                    const void* aux, int spk_kind, int64_t m, int64_t k,
                                    int64_t nse, int transpose, void* out, void* workspace,
                                    size_t workspace_bytes, void* stream) { BE_NVTX_RANGE("be_b200_binary_csrmv");
    return csr_entry(weights, w_dtype, homo, indices, indptr, indptr_is64, perm, perm_is64, spikes,
                     aux, spk_kind, m, k, nse, 0, transpose, out, workspace, workspace_bytes, stream);
}

// spk @ csr with the variant chosen by spike rate on the device (north star: "CSC/post-major
// gather (deterministic) next to an atomic CSR scatter variant chosen by spike rate").
extern "C" int be_b200_binary_csrmv_auto(const void* weights, int w_dtype, int homo,
                                         const int32_t* indices, const void* indptr, int indptr_is64,
                                         const void* csc_indptr, const int32_t* csc_indices,
                                         const void* perm, const void* spikes, const void* aux,
                                         int spk_kind, int64_t m, int64_t k, int64_t nse,
                                         float rate_threshold, void* out, void* workspace,
                                         size_t workspace_bytes, void* stream) { BE_NVTX_RANGE("be_b200_binary_csrmv_auto");
    cudaStream_t s = (cudaStream_t)stream;
    BE_REQUIRE(m >= 0 && k >= 0 && m < (1ll << 31) && k < (1ll << 31), "csrmv_auto: shape out of range");
    BE_REQUIRE(weights && indptr && csc_indptr && out, "csrmv_auto: NULL pointer");
    BE_REQUIRE((indices && csc_indices && (homo || perm)) || nse == 0, "csrmv_auto: NULL structure pointer");
    BE_REQUIRE(w_dtype >= 0 && w_dtype < 4, "csrmv_auto: unsupported weight dtype %d", w_dtype);
    static const size_t kWt[4] = {4, 8, 2, 2};
    if (k == 0) return BE_B200_OK;
    Events ev;
    int rc = prepare_events_1d(spikes, aux, spk_kind, m, true, true, workspace, workspace_bytes, s, &ev, out,
                               (size_t)k * kWt[w_dtype], /*ordered_ids=*/false);
    if (rc) return rc;
    if (m == 0) return BE_B200_OK;
    ev.sel_thr = (int)std::min<double>((double)rate_threshold * (double)m, 2147483647.0);
    ConnArgs cs{weights, w_dtype, homo, indices, nullptr, 0, m > 0 ? (double)nse / (double)m : 0.0};
    ConnArgs cg{weights, w_dtype, homo, csc_indices, homo ? nullptr : perm, homo ? 0 : (indptr_is64 ? 2 : 1),
                k > 0 ? (double)nse / (double)k : 0.0, nse};
    RowsDesc rg{indptr_is64 ? 1 : 0, csc_indptr, 0};
    BE_DISPATCH_WT(w_dtype, T, {
        if (indptr_is64) {
            CsrRows<int64_t> rows{(const int64_t*)indptr};
            rc = homo ? launch_scatter_mv<T, true, 0>(rows, cs, ev, m, (T*)out, s) : launch_scatter_mv<T, false, 0>(rows, cs, ev, m, (T*)out, s);
        } else {
            CsrRows<int32_t> rows{(const int32_t*)indptr};
            rc = homo ? launch_scatter_mv<T, true, 0>(rows, cs, ev, m, (T*)out, s) : launch_scatter_mv<T, false, 0>(rows, cs, ev, m, (T*)out, s);
        }
    });
    if (rc) return rc;
    return run_gather(rg, cg, ev, m, k, 0, out, s);
}

extern "C" int be_b200_binary_csrmm(const void* weights, int w_dtype, int homo,
                                    const int32_t* indices, const void* indptr, int indptr_is64,
                                    const void* perm, int perm_is64, const void* spikes,
                                    const void* aux, int spk_kind, int64_t m, int64_t k,
                                    int64_t nse, int64_t n, int transpose, int out_batch_major,
                                    void* out, void* workspace, size_t workspace_bytes, void* stream) { BE_NVTX_RANGE("be_b200_binary_csrmm");
    BE_REQUIRE(n >= 1, "csrmm: batch must be >= 1");
    return csr_entry(weights, w_dtype, homo, indices, indptr, indptr_is64, perm, perm_is64, spikes,
                     aux, spk_kind, m, k, nse, n, transpose, out, workspace, workspace_bytes, stream, out_batch_major);
}

static int fcn_entry(const void* weights, int w_dtype, int homo, const int32_t* indices,
                     int64_t n_rows, int64_t n_conn, const void* spikes, const void* aux,
                     int spk_kind, int64_t n_src, int64_t n_out, int64_t n, int transpose,
                     void* out, void* workspace, size_t workspace_bytes, void* stream,
                     int out_batch_major = 0) {
    cudaStream_t s = (cudaStream_t)stream;
    BE_REQUIRE(!out_batch_major || (transpose && n > 0), "batch-major output is only defined for the scatter matmat");
    BE_REQUIRE(n_rows >= 0 && n_conn >= 0 && n_src >= 0 && n_out >= 0, "fcn: negative size");
    BE_REQUIRE(n_rows < (1ll << 31) && n_src < (1ll << 31) && n_out < (1ll << 31), "fcn: size out of range");
    BE_REQUIRE(weights && out, "fcn: NULL pointer");
    BE_REQUIRE(indices || n_rows * n_conn == 0, "fcn: indices is NULL");
    if (transpose) BE_REQUIRE(n_src == n_rows, "fcn scatter: events length %lld != n_rows %lld", (long long)n_src, (long long)n_rows);
    else BE_REQUIRE(n_out == n_rows, "fcn gather: out length %lld != n_rows %lld", (long long)n_out, (long long)n_rows);
    ConnArgs c{weights, w_dtype, homo, indices, nullptr, 0, (double)n_conn};
    FcnRows rows{n_conn};
    BE_DISPATCH_WT(w_dtype, T, {
        return run_conn<T>(rows, RowsDesc{2, nullptr, n_conn}, c, spikes, aux, spk_kind, n_src, n_rows, n_out, n, transpose != 0, out, workspace, workspace_bytes, s, out_batch_major != 0);
    });
    return BE_B200_OK;
}

extern "C" int be_b200_binary_fcnmv(const void* weights, int w_dtype, int homo,
                                    const int32_t* indices, int64_t n_rows, int64_t n_conn,
                                    const void* spikes, const void* aux, int spk_kind,
                                    int64_t n_src, int64_t n_out, int transpose, void* out,
                                    void* workspace, size_t workspace_bytes, void* stream) { BE_NVTX_RANGE("be_b200_binary_fcnmv");
    return fcn_entry(weights, w_dtype, homo, indices, n_rows, n_conn, spikes, aux, spk_kind, n_src,
                     n_out, 0, transpose, out, workspace, workspace_bytes, stream);
}

extern "C" int be_b200_binary_fcnmm(const void* weights, int w_dtype, int homo,
                                    const int32_t* indices, int64_t n_rows, int64_t n_conn,
                                    const void* spikes, const void* aux, int spk_kind,
                                    int64_t n_src, int64_t n_out, int64_t n, int transpose,
                                    int out_batch_major, void* out, void* workspace,
                                    size_t workspace_bytes, void* stream) { BE_NVTX_RANGE("be_b200_binary_fcnmm");
    BE_REQUIRE(n >= 1, "fcnmm: batch must be >= 1");
    return fcn_entry(weights, w_dtype, homo, indices, n_rows, n_conn, spikes, aux, spk_kind, n_src,
                     n_out, n, transpose, out, workspace, workspace_bytes, stream, out_batch_major);
}
