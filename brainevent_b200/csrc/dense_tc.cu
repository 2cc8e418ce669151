// dense_tc.cu -- tcgen05 (5th-gen tensor core) path for the dense event matmul at batch >= 64.
//
//   out[m, n] = sum_k W[k, m] * e(S[k, n])          (binary_densemm, transpose=True, fp32 weights)
//
// This is the one place on the path that really is a contraction (north star: "tensor cores only
// for the dense BinaryArray@dense path at batch >= 64").  To stay inside the fp32 tolerance
// (rtol 1e-5) the fp32 weight is split on the fly into three bf16 planes, w = hi + mid + lo
// (8+8+8 significand bits: an error-free split up to 2^-24 relative); the event operand is an
// exact 0/1 bf16 matrix expanded from the bit-packed spikes, so every product is exact and the
// tensor core only performs fp32 accumulation (3 MMAs per K-step into the same TMEM tile).
//
// Mapping: one CTA per (128 output rows) x (64 batch columns) x (split of the K range).
//   * all 256 threads stream a [64 k, 128 m] fp32 tile of W with 128-bit loads, split it into the
//     three bf16 planes and write them straight into the canonical MN-major (no-swizzle) UMMA
//     shared-memory layout; the spike bits of the same 64 sources are expanded to bf16 0/1;
//   * one elected thread issues tcgen05.mma.kind::f16 (M=128, N=64, K=16) with the accumulator
//     in TMEM; tcgen05.commit on an mbarrier hands the shared-memory stage back (2 stages);
//   * event-driven at ROW granularity: the CTA compacts the sources of its K range that fired in at
//     least one batch column of its tile into a shared-memory list (ascending) and builds its
//     k-blocks from 32 consecutive LIST entries, so a silent source row of W is never read (the
//     CUDA cores place the operands in the UMMA layout themselves, so gathering rows is free);
//   * epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> registers -> out (plain stores, or
//     RED when the K range is split across CTAs).
#include "common.cuh"

#include <algorithm>
#include <cstdlib>

namespace be {

constexpr int TC_BM = 128, TC_BK = 32, TC_THREADS = 256, TC_STAGES = 2;   // BN (batch tile) is a template parameter: 64 / 128 / 256
constexpr int TC_A_PLANE = TC_BM * TC_BK * 2;            // 8 KB: one bf16 plane of the A tile
constexpr int TC_CHUNK = 3072;                             // source rows compacted per pass (u16 list in shared memory; with
                                                           // 4096 the static arrays push 3 CTAs/SM past the 196 KB carve-out and
                                                           // leave only 28 KB of L1 for the W loads in flight)
constexpr int TC_CHUNK_PASSES = TC_CHUNK / TC_THREADS;    // rows per thread per pass
static_assert((TC_CHUNK_PASSES * (TC_THREADS / 32)) % 32 == 0, "the ballot words are scanned by one warp, an equal share per lane");
constexpr int TC_ITEMS = (TC_BK / 8) * 8 / 8;            // [8 k x 16 m] load items per warp per k-block
__host__ __device__ constexpr int tc_stage_bytes(int bn) { return 3 * TC_A_PLANE + bn * TC_BK * 2; }
__host__ __device__ constexpr int tc_smem_bytes(int bn) { return TC_STAGES * tc_stage_bytes(bn) + 1024; }
// three CTAs per SM (batch tile 64) must fit the 196 KB shared-memory carve-out, which leaves 60 KB of
// L1 for the ~48 KB of W loads in flight per SM; one step up (228 KB, 28 KB of L1) measured 7 % slower
static_assert(3 * (tc_smem_bytes(64) + (TC_CHUNK + TC_BK) * 2 + TC_CHUNK_PASSES * 8 * 6 + 64 + 1024) <= 196 * 1024,
              "static + dynamic shared memory of dense_mm_t_tc_kernel<64> exceeds the 196 KB carve-out at 3 CTAs/SM");

// canonical MN-major, no swizzle: core matrix = 8 (K) rows x 16 B (8 bf16 along MN), 128 B each.
// A: core (mi, kk) at (kk * 16 + mi) * 128  -> SBO (MN step) = 128 B, LBO (K step) = 2048 B
// B: core (ni, kk) at (kk * BN/8 + ni) * 128  -> SBO = 128 B, LBO = BN/8 * 128 B
constexpr uint32_t TC_A_SBO = 128, TC_A_LBO = 16 * 128, TC_B_SBO = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
    return d;                 // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}

// kind::f16 instruction descriptor: D=f32, A=B=bf16, both MN-major, M=128, N=bn
__host__ __device__ constexpr uint32_t tc_idesc(int bn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(bn >> 3) << 17) |
           ((uint32_t)(TC_BM >> 4) << 24);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    const uint32_t addr = smem_u32(bar);
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc),
        "r"(idesc), "r"(accumulate)
        : "memory");
}

// (a, b) -> packed bf16x2 planes hi / mid / lo with a = hi.x + mid.x + lo.x (to 2^-24 relative)
__device__ __forceinline__ uint32_t cvt_bf16x2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));   // upper half <- first source operand
    return r;
}
__device__ __forceinline__ void split3(float a, float b, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
    hi = cvt_bf16x2(a, b);
    const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xffff0000u);
    mid = cvt_bf16x2(ra, rb);
    const float sa = ra - __uint_as_float(mid << 16), sb = rb - __uint_as_float(mid & 0xffff0000u);
    lo = cvt_bf16x2(sa, sb);
}

template <int TC_BN>
__global__ void __launch_bounds__(TC_THREADS, (TC_BN == 64) ? 3 : 2)
dense_mm_t_tc_kernel(const float* __restrict__ W, int64_t k, int64_t m, const uint32_t* __restrict__ packed, int wpr,
                     int64_t n, float* __restrict__ out, int kb_per_split, int split_k, int lane_map, int pf_dist) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t empty_bar[TC_STAGES];     // MMAs of a stage have finished reading it (tcgen05.commit)
    __shared__ uint64_t full_bar[TC_STAGES];      // all 8 warps have stored their part of a stage
    __shared__ uint32_t tmem_base_s;

    constexpr int TC_STAGE_BYTES = tc_stage_bytes(TC_BN);
    constexpr uint32_t TC_B_LBO = (TC_BN / 8) * 128;
    constexpr int WPT = TC_BN / 32;                    // spike words per source row in this tile
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t m0 = (int64_t)blockIdx.x * TC_BM;
    const int64_t n0 = (int64_t)blockIdx.z * TC_BN;
    const int total_kb = (int)((k + TC_BK - 1) / TC_BK);
    const int kb_begin = blockIdx.y * kb_per_split;
    const int kb_end = min(total_kb, kb_begin + kb_per_split);

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&empty_bar[s], 1); mbar_init(&full_bar[s], TC_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "n"(TC_BN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;

    // ---- loop-invariant per-thread coordinates ---------------------------------------------------
    // each warp streams TC_ITEMS [8 k x 16 m] patches of the tile: lane = (k8 = lane/4, quad = lane%4);
    // the spike words this thread will expand into the B tile are fetched with the same batch of loads.
    __shared__ uint16_t s_rows[TC_CHUNK + TC_BK];                   // active rows of the chunk (chunk-relative), 0xffff padded
    __shared__ uint32_t s_ballot[TC_CHUNK_PASSES * (TC_THREADS / 32)];
    __shared__ uint16_t s_off[TC_CHUNK_PASSES * (TC_THREADS / 32)];
    __shared__ int s_total;
    constexpr int B_CELLS = TC_BK * (TC_BN / 8) / TC_THREADS;      // (k, n-octet) cells per thread
    // lane -> (source slot k8, column quad) of a patch.  Map 0 (default): k8 = lane / 4, quad = lane % 4.  Map 2: bit 0 = quad
    // parity, bits 1-3 = k8, bit 4 = quad / 2 -- a half-warp then owns all 8 slots x 2 quads of ONE core-matrix column pair, so its 64-bit plane stores cover
    // 128 distinct bytes (16 bank pairs) -- with map 0 (k8 = lane / 4) a half-warp held slots 0-3 of two core matrices,
    // which sit on the same banks: every plane store took two wavefronts (ncu r2_dense: 13.9 M store bank conflicts
    // in 28.3 M store wavefronts, LSU data pipe 71 % busy).  Global side: a quarter-warp reads 4 rows x 32 B = 4 full
    // sectors, as before (2 rows x 64 B).  Measured 3 % SLOWER than map 0 (profiles/r2_experiments.txt): kept as a switch only.
    const int k8_l = lane_map == 2 ? ((lane >> 1) & 7) : (lane_map == 1 ? (lane & 7) : (lane >> 2));
    const int quad_l = lane_map == 2 ? ((lane & 1) | ((lane >> 4) << 1)) : (lane_map == 1 ? (lane >> 3) : (lane & 3));
    // with TC_ITEMS == 4 the four patches of a warp share their source slot (kk = warp / 2) and sit
    // 16 columns apart, so one list lookup and one 64-bit multiply address all four 128-bit loads
    static_assert(TC_ITEMS == 4 && TC_BK == 32 && TC_THREADS == 256, "patch mapping below assumes 4 items per warp");
    const int a_k = (warp >> 1) * 8 + k8_l;                        // my source slot inside a k-block
    const int64_t a_col0 = m0 + (warp & 1) * 64 + quad_l * 4;     // my first output column (item i: + 16 i)
    uint32_t a_soff[TC_ITEMS];    // byte offset of my 4 elements inside an A plane
#pragma unroll
    for (int i = 0; i < TC_ITEMS; ++i) {
        const int m_local = ((warp & 1) * 4 + i) * 16 + quad_l * 4;
        a_soff[i] = (uint32_t)(((warp >> 1) * 16 + (m_local >> 3)) * 128 + k8_l * 16 + (m_local & 7) * 2);
    }
    int b_k[B_CELLS], b_sh[B_CELLS], b_w[B_CELLS];
    int64_t b_c0[B_CELLS];
    uint32_t b_soff[B_CELLS];
#pragma unroll
    for (int i = 0; i < B_CELLS; ++i) {
        const int cell = tid + i * TC_THREADS;              // k8 fastest: 8 lanes fill one 128-byte core matrix
        const int k_local = (cell / (8 * (TC_BN / 8))) * 8 + (cell & 7), ni = (cell >> 3) % (TC_BN / 8);
        const int64_t c0 = n0 + ni * 8;
        b_k[i] = k_local;
        b_c0[i] = c0;
        b_sh[i] = (int)(c0 & 31);
        b_w[i] = (int)(c0 >> 5);
        b_soff[i] = (uint32_t)(((k_local >> 3) * (TC_BN / 8) + ni) * 128 + (k_local & 7) * 16);
    }
    const bool mn_interior = (m0 + TC_BM <= m) && (n0 + TC_BN <= n);
    const int64_t row_begin = (int64_t)kb_begin * TC_BK;
    const int64_t row_end = min(k, (int64_t)kb_end * TC_BK);

    // L2 prefetch of the W row segments `pf_dist` k-blocks ahead (cp.async.bulk.prefetch.L2, one 512-byte row segment
    // per lane of one warp): the bytes in flight towards HBM then no longer depend on the registers of the 256
    // converter threads -- their 128-bit loads find the rows in L2 (~1/3 of the DRAM latency)
    const bool pf_ok = pf_dist > 0 && (m & 3) == 0 && ((reinterpret_cast<uintptr_t>(W) & 15u) == 0) && m0 < m;
    const uint32_t pf_bytes = (uint32_t)(min((int64_t)TC_BM, m - m0) * 4);

    int n_issued = 0;      // k-blocks that actually reached the tensor core (uniform across the CTA)
    for (int64_t c_begin = row_begin; c_begin < row_end; c_begin += TC_CHUNK) {
        const int c_rows = (int)min((int64_t)TC_CHUNK, row_end - c_begin);
        const float* Wc = W + c_begin * m;
        const uint32_t* pc = packed + c_begin * wpr;
        // ---- ordered compaction of the chunk's active rows (active = fired in a column of this tile);
        // all spike words are requested before the first ballot so the passes share one round trip
        uint32_t anyv[TC_CHUNK_PASSES];
#pragma unroll
        for (int j = 0; j < TC_CHUNK_PASSES; ++j) {
            const int r = j * TC_THREADS + tid;
            uint32_t any = 0;
            if (r < c_rows) {
#pragma unroll
                for (int w = 0; w < WPT; ++w) {
                    const int wi = (int)(n0 >> 5) + w;
                    if (wi < wpr) {
                        uint32_t word = __ldg(pc + (int64_t)r * wpr + wi);
                        const int64_t cbase = (int64_t)wi * 32;
                        if (cbase + 32 > n) word &= (n > cbase) ? ((1u << (n - cbase)) - 1u) : 0u;
                        any |= word;
                    }
                }
            }
            anyv[j] = any;
        }
#pragma unroll
        for (int j = 0; j < TC_CHUNK_PASSES; ++j) {
            const uint32_t bal = __ballot_sync(0xffffffffu, anyv[j] != 0);
            if (lane == 0) s_ballot[j * (TC_THREADS / 32) + warp] = bal;
        }
        __syncthreads();
        if (warp == 0) {
            constexpr int PER = TC_CHUNK_PASSES * (TC_THREADS / 32) / 32;     // ballot words per lane
            int c[PER], sum = 0;
#pragma unroll
            for (int q = 0; q < PER; ++q) { c[q] = __popc(s_ballot[lane * PER + q]); sum += c[q]; }
            int incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            int run = incl - sum;
#pragma unroll
            for (int q = 0; q < PER; ++q) { s_off[lane * PER + q] = (uint16_t)run; run += c[q]; }
            if (lane == 31) s_total = incl;
        }
        __syncthreads();
        const int total = s_total;
#pragma unroll
        for (int j = 0; j < TC_CHUNK_PASSES; ++j) {
            const uint32_t bal = s_ballot[j * (TC_THREADS / 32) + warp];
            if ((bal >> lane) & 1u)
                s_rows[s_off[j * (TC_THREADS / 32) + warp] + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)(j * TC_THREADS + tid);
        }
        if (tid < TC_BK) s_rows[total + tid] = 0xffffu;
        __syncthreads();
        const int n_vkb = (total + TC_BK - 1) / TC_BK;

        auto prefetch_tile = [&](int vkb) {      // called by ONE warp: lane = source slot of k-block vkb
            const uint32_t rid = s_rows[vkb * TC_BK + lane];
            if (rid != 0xffffu)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(Wc + (int64_t)rid * m + m0), "r"(pf_bytes) : "memory");
        };
        if (pf_ok)
            for (int v = 1 + warp; v < min(n_vkb, pf_dist); v += TC_THREADS / 32) prefetch_tile(v);

        auto load_tile = [&](int vkb, float4 (&wv)[TC_ITEMS], uint32_t (&bw)[B_CELLS]) {
            const uint16_t* rows = s_rows + vkb * TC_BK;
            if (mn_interior && (vkb + 1) * TC_BK <= total) {
                const float4* wrow = reinterpret_cast<const float4*>(Wc + (int64_t)rows[a_k] * m + a_col0);
                // plain __ldg: ld.global.nc.L1::no_allocate measured 4 % slower here
#pragma unroll
                for (int i = 0; i < TC_ITEMS; ++i) wv[i] = __ldg(wrow + i * 4);
#pragma unroll
                for (int i = 0; i < B_CELLS; ++i)      // raw word: shifted where it is expanded, so nothing waits on this load here
                    bw[i] = __ldg(pc + (int64_t)rows[b_k[i]] * wpr + b_w[i]);
                return;
            }
            const uint32_t rid_a = rows[a_k];
#pragma unroll
            for (int i = 0; i < TC_ITEMS; ++i) {
                const int64_t col = a_col0 + i * 16;
                float t[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    t[q] = (rid_a != 0xffffu && col + q < m) ? __ldg(Wc + (int64_t)rid_a * m + col + q) : 0.0f;
                wv[i] = make_float4(t[0], t[1], t[2], t[3]);
            }
#pragma unroll
            for (int i = 0; i < B_CELLS; ++i) {
                const uint32_t rid = rows[b_k[i]];
                uint32_t bits = 0;
                if (rid != 0xffffu && b_c0[i] < n) {
                    bits = (__ldg(pc + (int64_t)rid * wpr + b_w[i]) >> b_sh[i]) & 0xffu;
                    if (b_c0[i] + 8 > n) bits &= (1u << (n - b_c0[i])) - 1u;
                }
                bw[i] = bits << b_sh[i];               // same form as the fast path (the consumer shifts back)
            }
        };

        // one k-block: request block vkb + 1 into (nxt, nxt_b), then convert / stage / multiply block
        // vkb from (cur, cur_b).  The caller alternates two explicit register sets, so the loads of
        // the next block are always issued BEFORE the wait on the current one (two blocks in flight)
        auto step = [&](int vkb, float4 (&cur)[TC_ITEMS], uint32_t (&cur_b)[B_CELLS], float4 (&nxt)[TC_ITEMS],
                        uint32_t (&nxt_b)[B_CELLS]) {
            if (pf_ok && vkb + pf_dist < n_vkb && warp == (vkb & 7)) prefetch_tile(vkb + pf_dist);
            if (vkb + 1 < n_vkb) load_tile(vkb + 1, nxt, nxt_b);

            const int stage = n_issued % TC_STAGES;
            uint8_t* sA = smem + stage * TC_STAGE_BYTES;
            uint8_t* sB = sA + 3 * TC_A_PLANE;
            if (n_issued >= TC_STAGES) mbar_wait(&empty_bar[stage], ((n_issued / TC_STAGES) - 1) & 1);
            // ---- A: split into bf16 hi/mid/lo planes, MN-major core matrices --------------------
#pragma unroll
            for (int i = 0; i < TC_ITEMS; ++i) {
                // error-free 3-way split, two elements per cvt.rn.bf16x2.f32
                uint32_t hi[2], mid[2], lo[2];
                split3(cur[i].x, cur[i].y, hi[0], mid[0], lo[0]);
                split3(cur[i].z, cur[i].w, hi[1], mid[1], lo[1]);
                *reinterpret_cast<uint2*>(sA + a_soff[i]) = make_uint2(hi[0], hi[1]);
                *reinterpret_cast<uint2*>(sA + TC_A_PLANE + a_soff[i]) = make_uint2(mid[0], mid[1]);
                *reinterpret_cast<uint2*>(sA + 2 * TC_A_PLANE + a_soff[i]) = make_uint2(lo[0], lo[1]);
            }
            // ---- B: 0/1 bf16 from the prefetched spike bits -----------------------------------------
#pragma unroll
            for (int i = 0; i < B_CELLS; ++i) {
                const uint32_t bits = (cur_b[i] >> b_sh[i]) & 0xffu;
                uint32_t v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    v[q] = ((bits >> (2 * q)) & 1u ? 0x3f80u : 0u) | ((bits >> (2 * q + 1)) & 1u ? 0x3f800000u : 0u);
                *reinterpret_cast<uint4*>(sB + b_soff[i]) = make_uint4(v[0], v[1], v[2], v[3]);
            }
            // generic-proxy writes -> visible to the tensor core (async proxy); then the warp reports its part of the
            // stage on the stage's mbarrier.  No CTA barrier: only the warp whose turn it is to issue the MMAs
            // (round robin, step % 8) waits for the other seven -- the rest go straight on to the next k-block
            // (ncu r2: 13 % of the samples sat at the per-k-block __syncthreads, with thread 0's MMA issue behind it)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[stage]);
            if (warp == (n_issued & 7)) {
                mbar_wait(&full_bar[stage], (n_issued / TC_STAGES) & 1);
                if (lane == 0) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
#pragma unroll
                    for (int ks = 0; ks < TC_BK / 16; ++ks) {
                        const uint64_t bd = umma_desc(b0 + ks * 2 * TC_B_LBO, TC_B_LBO, TC_B_SBO);
#pragma unroll
                        for (int p = 0; p < 3; ++p) {
                            const uint64_t ad = umma_desc(a0 + p * TC_A_PLANE + ks * 2 * TC_A_LBO, TC_A_LBO, TC_A_SBO);
                            umma_bf16(tmem_d, ad, bd, tc_idesc(TC_BN), (n_issued > 0 || ks > 0 || p > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty_bar[stage]);     // arrives when these MMAs have finished reading smem
                }
                __syncwarp();
            }
            ++n_issued;
        };
        float4 ra[TC_ITEMS], rb[TC_ITEMS];
        uint32_t ra_b[B_CELLS], rb_b[B_CELLS];
        if (n_vkb > 0) load_tile(0, ra, ra_b);
        for (int vkb = 0; vkb < n_vkb; vkb += 2) {
            step(vkb, ra, ra_b, rb, rb_b);
            if (vkb + 1 < n_vkb) step(vkb + 1, rb, rb_b, ra, ra_b);
        }
        // every thread is past the last block's __syncthreads, i.e. done reading s_rows: the next
        // chunk may overwrite the lists
    }

    // ---- drain + epilogue ---------------------------------------------------------------------
    if (n_issued > 0) {
        // the tensor pipe runs a CTA's MMAs in issue order and every step's commit covers the MMAs its issuing
        // thread started: the last commit on each stage's barrier therefore covers everything
#pragma unroll
        for (int st = 0; st < TC_STAGES; ++st) {
            const int uses = (n_issued - st + TC_STAGES - 1) / TC_STAGES;
            if (uses > 0) mbar_wait(&empty_bar[st], (uses - 1) & 1);
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // warp w reads TMEM lanes 32*(w%4)..+31 (its rows); 32-column chunks alternate between
        // the two warps that share a lane quadrant
        const int row_q = warp & 3;
        const int64_t row = m0 + row_q * 32 + lane;
        for (int ch = warp >> 2; ch < TC_BN / 32; ch += 2) {
            uint32_t r[32];
            const uint32_t taddr = tmem_d + ((uint32_t)(row_q * 32) << 16) + (uint32_t)(ch * 32);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < m) {
                float* o = out + row * n + n0 + ch * 32;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    if (n0 + ch * 32 + c >= n) break;
                    const float v = __uint_as_float(r[c]);
                    if (split_k > 1) { if (v != 0.0f) atomicAdd(o + c, v); }
                    else o[c] = v;
                }
            }
        }
    } else if (split_k == 1) {
        // nothing fired in this tile: the result is zero
        for (int e = tid; e < TC_BM * TC_BN; e += TC_THREADS) {
            const int64_t row = m0 + e / TC_BN, col = n0 + e % TC_BN;
            if (row < m && col < n) out[row * n + col] = 0.0f;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TC_BN) : "memory");
}

// Host entry used by dense.cu: W [k, m] fp32, packed events [k, wpr] -> out [m, n] fp32.
int launch_dense_mm_t_tc(const float* W, int64_t k, int64_t m, const uint32_t* packed, int wpr, int64_t n,
                         float* out, cudaStream_t s) {
    if (m == 0 || n == 0) return BE_B200_OK;
    const char* lm = getenv("BE_B200_TC_LANEMAP");
    const int lane_map = lm ? atoi(lm) : 0;         // 0: row-quad lanes (fastest measured); 1 / 2: alternative maps kept for A/B
    const char* pfe = getenv("BE_B200_TC_PREFETCH");
    const int pf_dist = pfe ? atoi(pfe) : 0;        // k-blocks of L2 prefetch distance; measured 8 % SLOWER at every distance 2..16, off
    // widest batch tile that the batch fills: W is read once per batch tile
    const int bn = n > 128 ? 256 : (n > 64 ? 128 : 64);
    const int total_kb = (int)ceil_div(k, TC_BK);
    const int64_t tiles = ceil_div(m, TC_BM) * ceil_div(n, bn);
    // split K so that the grid is ONE wave of resident CTAs (3 per SM at batch tile 64, 2 above): CTAs of a call all take
    // the same time, so 768 CTAs on 444 slots ran as two waves with the second 73 % full -- 384 CTAs (split 3) are 3 %
    // faster at batch 64 and 256 instead of 768 CTAs 11 % faster at batch 256 (profiles/r2_experiments.txt); fewer
    // splits also mean fewer RED epilogues
    const int64_t slots = (int64_t)(bn == 64 ? 3 : 2) * kNumSM;
    int split_k = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)total_kb, slots / std::max<int64_t>(tiles, 1), (int64_t)32}));
    if (const char* e = getenv("BE_B200_TC_SPLITK")) split_k = std::max(1, std::min(atoi(e), std::max(total_kb, 1)));   // A/B switch
    const int kb_per_split = (int)ceil_div(std::max(total_kb, 1), split_k);
    split_k = (int)ceil_div(std::max(total_kb, 1), kb_per_split);
    if (split_k > 1) BE_CUDA(cudaMemsetAsync(out, 0, (size_t)(m * n) * sizeof(float), s));
#define BE_TC_LAUNCH(BN)                                                                                       \
    do {                                                                                                       \
        BE_CUDA(cudaFuncSetAttribute(dense_mm_t_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                     tc_smem_bytes(BN)));                                                      \
        dim3 grid((unsigned)ceil_div(m, TC_BM), (unsigned)split_k, (unsigned)ceil_div(n, BN));                 \
        ProfileScope prof(s);                                                                                  \
        dense_mm_t_tc_kernel<BN><<<grid, TC_THREADS, tc_smem_bytes(BN), s>>>(W, k, m, packed, wpr, n, out,     \
                                                                            kb_per_split, split_k, lane_map, pf_dist); \
    } while (0)
    if (bn == 64) BE_TC_LAUNCH(64); else if (bn == 128) BE_TC_LAUNCH(128); else BE_TC_LAUNCH(256);
#undef BE_TC_LAUNCH
    BE_LAUNCH_CHECK("dense_mm_t_tc_kernel");
    return BE_B200_OK;
}

}  // namespace be
