// csr_fcn_gather.cu -- deterministic post-major gather through stored connectivity
// (ref: _csr/binary_csrmv.cu NT, _csr/binary_csrmm.cu NT, FCN gather).  Split from csr_fcn.cu so
// the two translation units compile in parallel.
#include "conn.cuh"

#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace be {

// ------------------------------------------------------------------------------------------
// gather, matvec:  out[i] = sum_j w[j] * bit(indices[j])   (all rows, deterministic)
//
// The kernel streams the whole index array once (HBM-bound by construction); what decides
// whether it reaches the HBM roofline is the cost of the per-synapse random lookup into the
// bit-packed spike vector.  Through L1 a warp's 32 random addresses cost up to 32 wavefronts, which
// caps the chip at ~5e11 lookups/s (2 TB/s of int32 indices).  SMEM=true copies the packed vector
// (n/8 bytes: 125 KB for 1 M neurons) into shared memory once per persistent CTA, where 32 random
// 4-byte reads cost ~3-4 bank cycles, and runs 1024 threads per SM to keep the streams full.
// ------------------------------------------------------------------------------------------
template <typename T, bool HOMO, int PERM, bool VEC, int LPR, bool SMEM, typename Rows, int UQ = 0>
__global__ void __launch_bounds__(SMEM ? 1024 : 256) gather_mv_kernel(Rows rows, const int32_t* __restrict__ indices,
                                                                      const T* __restrict__ weights,
                                                                      const void* __restrict__ perm,
                                                                      const uint32_t* __restrict__ packed,
                                                                      int64_t n_words, int64_t n_rows,
                                                                      T* __restrict__ out,
                                                                      const int* __restrict__ sel_n_active, int sel_thr,
                                                                      const uint8_t* __restrict__ raw_bool, int64_t raw_n,
                                                                      int pend_off_words) {
    using A = typename Acc<T>::type;
    if (sel_thr >= 0 && *sel_n_active <= sel_thr) return;   // the scatter variant owns this call
    extern __shared__ __align__(16) uint32_t s_bits[];
    // fp32 weights of the hits travel global -> shared memory by 4-byte cp.async into per-lane slots and are summed
    // when the lane's NEXT row is done: ncu (r2, source page) had 36 % of the hetero kernel's samples waiting on the
    // weight quads it had just requested.  pend_off_words == 0: no room beside the spike bits, load in place.
    constexpr int NP = 4;                                   // deferred hits per lane and row; more are summed in place
    constexpr bool ASYNCW_T = SMEM && VEC && !HOMO && PERM == 0 && std::is_same<T, float>::value;
    const bool asyncw = ASYNCW_T && pend_off_words > 0;
    float* const s_pend = reinterpret_cast<float*>(s_bits + pend_off_words);    // [2][blockDim.x][NP]
    A pend_acc = A(0);
    int pend_n = 0, pend_buf = 0;
    int64_t pend_row = -1;
    // the first row's extent (and the homogeneous weight) are requested before the spike vector is
    // staged, so their latency overlaps the staging instead of following it
    const A w0 = HOMO ? to_acc<T>(__ldg(weights)) : A(0);
    const int GROUPS = blockDim.x / LPR;
    const int g = threadIdx.x / LPR, gl = threadIdx.x % LPR;
    const int64_t row_stride = (int64_t)gridDim.x * GROUPS;
    const int64_t row_iters = ceil_div(n_rows, row_stride);
    int64_t r = (int64_t)blockIdx.x * GROUPS + g;
    typename Rows::ptr_t s_n = 0, e_n = 0;      // raw (stored-width) extent of the next row, widened where it is used:
    if (r < n_rows) rows.get_raw(r, s_n, e_n);  // no instruction consumes the load right behind it (ncu: 16 % of the samples)
    if (SMEM && raw_bool != nullptr) {
        // small bool vector: every CTA packs the raw bytes itself (no separate pack kernel)
        for (int64_t w = threadIdx.x; w < n_words; w += blockDim.x) {
            const int64_t base = w * 32;
            uint32_t bits = 0;
            if ((((uintptr_t)raw_bool) & 15u) == 0 && base + 32 <= raw_n) {
                const int4 v0 = __ldg((const int4*)(raw_bool + base)), v1 = __ldg((const int4*)(raw_bool + base) + 1);
                const uint32_t q[8] = {(uint32_t)v0.x, (uint32_t)v0.y, (uint32_t)v0.z, (uint32_t)v0.w,
                                       (uint32_t)v1.x, (uint32_t)v1.y, (uint32_t)v1.z, (uint32_t)v1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if ((q[i] >> (8 * b)) & 0xffu) bits |= 1u << (i * 4 + b);
            } else {
                for (int b = 0; b < 32 && base + b < raw_n; ++b)
                    if (__ldg(raw_bool + base + b)) bits |= 1u << b;
            }
            s_bits[w] = bits;
        }
        __syncthreads();
    } else if (SMEM) {
        // TMA bulk copy (cp.async.bulk, SASS UBLKCP) of the packed spike vector into shared memory:
        // one thread issues <= 32 KB chunks against an mbarrier, everybody waits on its phase
        __shared__ uint64_t s_bar;
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
        const int64_t n16 = n_words >> 2;
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t total = (uint32_t)(n16 * 16);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(total) : "memory");
            const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(s_bits);
            for (uint32_t off = 0; off < total; off += 32768u) {
                const uint32_t sz = min(32768u, total - off);
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                    ::"r"(dst0 + off), "l"((const char*)packed + off), "r"(sz), "r"(bar)
                    : "memory");
            }
        }
        for (int64_t i = (n16 << 2) + threadIdx.x; i < n_words; i += blockDim.x) s_bits[i] = __ldg(packed + i);
        {
            uint32_t done = 0;
            while (!done)
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done) : "r"(bar) : "memory");
        }
        __syncthreads();
    }
    auto bit = [&](int idx) -> bool {
        const uint32_t w = SMEM ? s_bits[idx >> 5] : __ldg(packed + (idx >> 5));
        return (w >> (idx & 31)) & 1u;
    };
    for (int64_t it = 0; it < row_iters; ++it, r += row_stride) {
        A acc = A(0);
        int cnt = 0, nh = 0;
        const int64_t s = (int64_t)s_n, e = (int64_t)e_n;
        const bool live = r < n_rows;
        if (r + row_stride < n_rows) rows.get_raw(r + row_stride, s_n, e_n);   // next row's extent in flight
        if (live) {
            if (VEC) {
                int64_t a = (s + 3) & ~(int64_t)3;
                if (a > e) a = e;
                const int64_t aend = a + ((e - a) & ~(int64_t)3);
                if (gl < a - s && bit(__ldg(indices + s + gl))) {
                    if (HOMO) ++cnt; else acc += weight_at<T, HOMO, PERM>(weights, perm, s + gl, w0);
                }
                // 4 independent 128-bit index loads per lane are issued before the first lookup
                // narrow groups: more independent quads per lane.  UQ (launch-time choice, 4-lane groups only): just enough
                // quads for the expected row -- 7 instead of 8 for 100-synapse rows, whose 8th slot is always masked
                constexpr int U = UQ > 0 ? UQ : ((LPR <= 4) ? 8 : 4);
                for (int64_t j0 = a + (int64_t)gl * 4; j0 < aend; j0 += (int64_t)LPR * 4 * U) {
                    int4 ix[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int64_t j = j0 + (int64_t)u * LPR * 4;
                        ix[u] = (j < aend) ? ldg_stream((const int4*)(indices + j)) : make_int4(-1, -1, -1, -1);
                    }
                    if (!HOMO && PERM == 0 && std::is_same<T, float>::value) {
                        // fp32 weights: first all lookups (4-bit hit mask per quad), then ALL the weight quads
                        // that were hit are requested with predicated 128-bit loads before the first one is
                        // used -- one exposed latency per trip instead of one per hit quad, and no divergence
                        uint32_t hm = 0;
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            if (ix[u].x < 0) continue;
                            hm |= ((uint32_t)bit(ix[u].x) | ((uint32_t)bit(ix[u].y) << 1) | ((uint32_t)bit(ix[u].z) << 2) |
                                   ((uint32_t)bit(ix[u].w) << 3)) << (4 * u);
                        }
                        if (hm && asyncw) {
                            // one 4-byte async copy per hit into this lane's slots of the row's buffer
                            const uint32_t pdst = (uint32_t)__cvta_generic_to_shared(
                                s_pend + ((size_t)(it & 1) * blockDim.x + threadIdx.x) * NP);
                            while (hm) {
                                const int b = __ffs(hm) - 1;
                                hm &= hm - 1;
                                const float* wp = (const float*)weights + j0 + (int64_t)(b >> 2) * LPR * 4 + (b & 3);
                                if (nh < NP) {
                                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(pdst + 4u * (uint32_t)nh), "l"(wp) : "memory");
                                    ++nh;
                                } else {
                                    acc += (A)__ldg(wp);
                                }
                            }
                        } else if (hm) {
                            int4 wq[U];
#pragma unroll
                            for (int u = 0; u < U; ++u) {
                                wq[u] = make_int4(0, 0, 0, 0);
                                if ((hm >> (4 * u)) & 15u)
                                    wq[u] = ldg_stream((const int4*)((const float*)weights + j0 + (int64_t)u * LPR * 4));
                            }
#pragma unroll
                            for (int u = 0; u < U; ++u) {
                                const uint32_t q = (hm >> (4 * u)) & 15u;
                                acc += (A)((q & 1u ? __int_as_float(wq[u].x) : 0.0f) + (q & 2u ? __int_as_float(wq[u].y) : 0.0f) +
                                           (q & 4u ? __int_as_float(wq[u].z) : 0.0f) + (q & 8u ? __int_as_float(wq[u].w) : 0.0f));
                            }
                        }
                        continue;
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (ix[u].x < 0) continue;
                        const int64_t j = j0 + (int64_t)u * LPR * 4;
                        const bool b0 = bit(ix[u].x), b1 = bit(ix[u].y), b2 = bit(ix[u].z), b3 = bit(ix[u].w);
                        if (HOMO) cnt += (int)b0 + (int)b1 + (int)b2 + (int)b3;
                        else if (b0 | b1 | b2 | b3) {
                            if (PERM == 0) {
                                Vec4<T> v = load4(weights + j);
                                acc += (b0 ? v.v[0] : A(0)) + (b1 ? v.v[1] : A(0)) + (b2 ? v.v[2] : A(0)) + (b3 ? v.v[3] : A(0));
                            } else {
                                if (b0) acc += weight_at<T, HOMO, PERM>(weights, perm, j, w0);
                                if (b1) acc += weight_at<T, HOMO, PERM>(weights, perm, j + 1, w0);
                                if (b2) acc += weight_at<T, HOMO, PERM>(weights, perm, j + 2, w0);
                                if (b3) acc += weight_at<T, HOMO, PERM>(weights, perm, j + 3, w0);
                            }
                        }
                    }
                }
                if (gl < e - aend && bit(__ldg(indices + aend + gl))) {
                    if (HOMO) ++cnt; else acc += weight_at<T, HOMO, PERM>(weights, perm, aend + gl, w0);
                }
            } else {
                for (int64_t j = s + gl; j < e; j += LPR)
                    if (bit(__ldg(indices + j))) {
                        if (HOMO) ++cnt; else acc += weight_at<T, HOMO, PERM>(weights, perm, j, w0);
                    }
            }
        }
        if (ASYNCW_T && asyncw) {
            // this row's copies form one group; the PREVIOUS row's group is complete once at most one is pending
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            A tot = pend_acc;
            const float* pw = s_pend + ((size_t)pend_buf * blockDim.x + threadIdx.x) * NP;
#pragma unroll
            for (int q = 0; q < NP; ++q)
                if (q < pend_n) tot += (A)pw[q];
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            if (gl == 0 && pend_row >= 0) out[pend_row] = from_acc<T>(tot);
            pend_acc = acc; pend_n = nh; pend_buf = (int)(it & 1); pend_row = live ? r : -1;
            continue;
        }
        if (HOMO) {
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            acc = w0 * (A)cnt;
        } else {
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        }
        if (gl == 0 && live) out[r] = from_acc<T>(acc);
    }
    if (ASYNCW_T && asyncw) {                      // the last row of every lane group
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        A tot = pend_acc;
        const float* pw = s_pend + ((size_t)pend_buf * blockDim.x + threadIdx.x) * NP;
#pragma unroll
        for (int q = 0; q < NP; ++q)
            if (q < pend_n) tot += (A)pw[q];
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (gl == 0 && pend_row >= 0) out[pend_row] = from_acc<T>(tot);
    }
}

// ------------------------------------------------------------------------------------------
// gather, matmat: out[i, c] = sum_j w[j] * bit(B[indices[j], c]); lanes = columns,
// CPL columns per lane (batch tile of 32*CPL columns per pass).
// ------------------------------------------------------------------------------------------
template <typename T, bool HOMO, int PERM, int CPL, typename Rows>
__global__ void __launch_bounds__(256) gather_mm_kernel(Rows rows, const int32_t* __restrict__ indices,
                                                        const T* __restrict__ weights,
                                                        const void* __restrict__ perm,
                                                        const uint32_t* __restrict__ packed, int wpr,
                                                        int64_t n_rows, int64_t n, T* __restrict__ out) {
    using A = typename Acc<T>::type;
    const A w0 = HOMO ? to_acc<T>(__ldg(weights)) : A(0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_pass = (wpr + CPL - 1) / CPL;
    for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < n_rows; r += (int64_t)gridDim.x * 8) {
        int64_t s, e;
        rows.get(r, s, e);
        for (int pass = 0; pass < n_pass; ++pass) {
            const int wb = pass * CPL;
            A acc[CPL];
#pragma unroll
            for (int q = 0; q < CPL; ++q) acc[q] = A(0);
            for (int64_t j0 = s; j0 < e; j0 += 32) {
                const int64_t j = j0 + lane;
                uint32_t mw[CPL];
                A my_w = A(0);
                bool nz = false;
#pragma unroll
                for (int q = 0; q < CPL; ++q) mw[q] = 0u;
                if (j < e) {
                    const uint32_t* brow = packed + (int64_t)__ldg(indices + j) * wpr + wb;
#pragma unroll
                    for (int q = 0; q < CPL; ++q) {
                        mw[q] = (wb + q < wpr) ? __ldg(brow + q) : 0u;
                        nz |= (mw[q] != 0u);
                    }
                    if (nz) my_w = weight_at<T, HOMO, PERM>(weights, perm, j, w0);
                }
                uint32_t live = __ballot_sync(0xffffffffu, nz);
                while (live) {
                    const int t = __ffs(live) - 1;
                    live &= live - 1;
                    const A wv = __shfl_sync(0xffffffffu, my_w, t);
#pragma unroll
                    for (int q = 0; q < CPL; ++q) {
                        const uint32_t m = __shfl_sync(0xffffffffu, mw[q], t);
                        if ((m >> lane) & 1u) acc[q] += wv;
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

constexpr size_t kGatherSmemMax = 200u << 10;   // packed spikes up to 200 KB (1.6 M sources) go to shared memory

template <typename T, bool HOMO, int PERM, typename Rows>
static int launch_gather_mv(const Rows& rows, const ConnArgs& c, const Events& ev, int64_t n_src, int64_t n_rows,
                            T* out, cudaStream_t s) {
    if (n_rows == 0) return BE_B200_OK;
    const bool vec = al16(c.indices) && (c.homo || al16(c.weights));
    // narrow groups keep more rows (and 8 independent 128-bit quads per lane) in flight: 0.69 vs 0.61
    // of the HBM peak at 100 synapses/row
    int lpr = c.avg_row <= 128 ? 4 : (c.avg_row <= 256 ? 8 : (c.avg_row <= 640 ? 16 : 32));
    if (const char* e = getenv("BE_B200_GATHER_LPR")) { const int v = atoi(e); if (v == 4 || v == 8 || v == 16 || v == 32) lpr = v; }   // A/B switch
    const int64_t n_words = ceil_div(n_src, 32);
    const size_t smem = (size_t)n_words * 4;
    // the shared-memory variant pays for itself when every CTA has enough rows to amortise the copy
    bool use_smem = vec && al16(ev.packed) && smem <= kGatherSmemMax &&
                    (double)n_rows * c.avg_row >= 4.0 * (double)n_words * kNumSM;
    if (const char* e = getenv("BE_B200_GATHER_SMEM")) use_smem = use_smem && e[0] != '0';   // A/B switch
    Events evl = ev;
    if (evl.raw_bool != nullptr && !use_smem) {       // only the shared-memory variant packs in-kernel
        const int rc = pack_deferred(&evl, s);
        if (rc) return rc;
    }
    ProfileScope prof(s);
    if (use_smem) {
        // small problems (a few thousand rows): 256-thread CTAs spread over more SMs instead of a
        // handful of 1024-thread ones -- the call is launch/latency-bound there
        const int threads = (n_rows * lpr <= (int64_t)kNumSM * 256) ? 256 : 1024;
        const int blocks = (int)std::min<int64_t>(kNumSM, ceil_div(n_rows, threads / lpr));
        // deferred-weight slots (fp32 hetero): 2 buffers x threads x 4 floats behind the spike bits, if they fit
        const bool defer_ok = !HOMO && PERM == 0 && std::is_same<T, float>::value && getenv("BE_B200_GATHER_DEFER_W") == nullptr;
        const size_t bits_pad = ((size_t)n_words + 3) & ~(size_t)3;
        size_t pend_bytes = defer_ok ? (bits_pad - (size_t)n_words) * 4 + (size_t)threads * 2 * 4 * 4 : 0;
        if (smem + pend_bytes > (227u << 10) - 1024) pend_bytes = 0;
        const int pend_off = pend_bytes ? (int)bits_pad : 0;
#define BE_GAS2(LPR, UQ)                                                                                      \
        do {                                                                                                  \
            BE_CUDA(cudaFuncSetAttribute(gather_mv_kernel<T, HOMO, PERM, true, LPR, true, Rows, UQ>,          \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227u << 10) - 1024));  \
            gather_mv_kernel<T, HOMO, PERM, true, LPR, true, Rows, UQ><<<blocks, threads, smem + pend_bytes, s>>>( \
                rows, c.indices, (const T*)c.weights, c.perm, evl.packed, n_words, n_rows, out, evl.n_active, evl.sel_thr, \
                evl.raw_bool, evl.raw_n, pend_off);                                                           \
        } while (0)
#define BE_GAS(LPR) BE_GAS2(LPR, 0)
        // quads per lane of a 4-lane group: ceil(row / 16), at least 5 (rows that are longer simply take a second trip)
        int uq = 8;
        if (const char* e = getenv("BE_B200_GATHER_UQ")) uq = atoi(e);
        else if (lpr == 4) uq = std::max(5, std::min(8, (int)((c.avg_row + 15.0) / 16.0)));
        if (lpr == 4 && uq >= 5 && uq < 8) {
            if (uq == 5) BE_GAS2(4, 5); else if (uq == 6) BE_GAS2(4, 6); else BE_GAS2(4, 7);
        } else {
            switch (lpr) { case 4: BE_GAS(4); break; case 8: BE_GAS(8); break; case 16: BE_GAS(16); break; default: BE_GAS(32); }
        }
#undef BE_GAS
#undef BE_GAS2
    } else {
        int blocks = (int)std::min<int64_t>(ceil_div(n_rows, 256 / lpr), (int64_t)kNumSM * 16);
#define BE_GA(VEC, LPR)                                                                      \
        gather_mv_kernel<T, HOMO, PERM, VEC, LPR, false, Rows><<<blocks, 256, 0, s>>>(       \
            rows, c.indices, (const T*)c.weights, c.perm, evl.packed, n_words, n_rows, out, evl.n_active, evl.sel_thr, \
            nullptr, 0, 0)
#define BE_GA2(VEC) switch (lpr) { case 4: BE_GA(VEC, 4); break; case 8: BE_GA(VEC, 8); break; \
                                   case 16: BE_GA(VEC, 16); break; default: BE_GA(VEC, 32); }
        if (vec) { BE_GA2(true) } else { BE_GA2(false) }
#undef BE_GA2
#undef BE_GA
    }
    BE_LAUNCH_CHECK("gather_mv_kernel");
    return BE_B200_OK;
}

template <typename T, bool HOMO, int PERM, typename Rows>
static int launch_gather_mm(const Rows& rows, const ConnArgs& c, const Events& ev, int64_t n_rows,
                            int64_t n, T* out, cudaStream_t s) {
    if (n_rows == 0) return BE_B200_OK;
    int blocks = (int)std::min<int64_t>(ceil_div(n_rows, 8), (int64_t)kNumSM * 16);
    const int wpr = ev.words_per_row;
#define BE_GM(CPL)                                                                           \
    gather_mm_kernel<T, HOMO, PERM, CPL, Rows><<<blocks, 256, 0, s>>>(                       \
        rows, c.indices, (const T*)c.weights, c.perm, ev.packed, wpr, n_rows, n, out)
    {
        ProfileScope prof(s);
        if (wpr == 1) BE_GM(1); else if (wpr == 2) BE_GM(2); else BE_GM(4);
    }
#undef BE_GM
    BE_LAUNCH_CHECK("gather_mm_kernel");
    return BE_B200_OK;
}


template <typename T, typename Rows>
static int run_gather_t(const Rows& rows, const ConnArgs& c, const Events& ev, int64_t n_src, int64_t n_rows,
                        int64_t n_batch, T* out, cudaStream_t s) {
    const bool mm = n_batch > 0;
#define BE_RUNG(HOMO, PERM)                                                                         \
    return mm ? launch_gather_mm<T, HOMO, PERM, Rows>(rows, c, ev, n_rows, n_batch, out, s)          \
              : launch_gather_mv<T, HOMO, PERM, Rows>(rows, c, ev, n_src, n_rows, out, s)
    if (c.homo) { BE_RUNG(true, 0); }
    if constexpr (!std::is_same<Rows, FcnRows>::value) {
        if (c.perm_kind == 1) { BE_RUNG(false, 1); }
        if (c.perm_kind == 2) { BE_RUNG(false, 2); }
    }
    BE_RUNG(false, 0);
#undef BE_RUNG
}

int run_gather(const RowsDesc& rd, const ConnArgs& c, const Events& ev, int64_t n_src, int64_t n_rows,
               int64_t n_batch, void* out, cudaStream_t s) {
    BE_DISPATCH_WT(c.w_dtype, T, {
        if (rd.kind == 0) { CsrRows<int32_t> rows{(const int32_t*)rd.indptr}; return run_gather_t<T>(rows, c, ev, n_src, n_rows, n_batch, (T*)out, s); }
        if (rd.kind == 1) { CsrRows<int64_t> rows{(const int64_t*)rd.indptr}; return run_gather_t<T>(rows, c, ev, n_src, n_rows, n_batch, (T*)out, s); }
        FcnRows rows{rd.n_conn};
        return run_gather_t<T>(rows, c, ev, n_src, n_rows, n_batch, (T*)out, s);
    });
    return BE_B200_OK;
}

}  // namespace be
