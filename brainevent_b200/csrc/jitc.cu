// jitc.cu -- just-in-time connectivity: the matrix is regenerated inside the kernel from
// (seed, row, chunk, lane) with a bit-exact copy of the reference's light-RNG walk.
//
// Replaces ref: _jit_scalar/binary_jitsmv.cu, binary_jitsmm.cu; _jit_normal/binary_jitn{mv,mm}.cu;
// _jit_uniform/binary_jitu{mv,mm}.cu.  RNG spec: _numba_random.py:385-502.
//
// Differences from the reference CUDA (same results, different machine mapping):
//  * scatter ("trans") is driven by the compacted active-row list: one warp per (active row,
//    chunk) task instead of a warp for every row that exits early when its spike bit is clear.
//  * gather ("notrans") keeps one warp per output row across ALL chunks and writes out[row] once:
//    no memset, no atomics, deterministic; the packed-spike word index / bit shift are hoisted
//    out of the walk (they are invariant per (chunk, lane)).
//  * mm draws each connection ONCE and applies it to every batch column (lanes = columns); the
//    reference re-runs the whole walk per column (SURVEY F8 ii).
//  * the per-edge Normal weight uses explicitly rounded (non-fused) f32 arithmetic so that it is
//    bit-identical to the reference's strict `light_rng_normal01`; Acklam's central rational is
//    ill-conditioned near |u-0.5| ~ 0.47 and FMA contraction alone moves it by ~1e-4.
#include "jitc.cuh"

#include <algorithm>

namespace be {

// ------------------------------------------------------------------------------------------
// mv gather (corder=True): out[row] = sum_{j in C(row)} w(row,j) * bit(j)
// one warp per output row, lane = residue class, loops over chunks; single plain store.
// ------------------------------------------------------------------------------------------
template <typename T, int KIND>
__global__ void __launch_bounds__(256) jit_mv_gather_kernel(JitParams jp, const uint32_t* __restrict__ packed,
                                                            int64_t vec_len, int64_t chunk_size,
                                                            int64_t row_offset, int64_t row_count,
                                                            T* __restrict__ out) {
    using A = typename Acc<T>::type;
    A p0, p1; uint32_t cl, seed;
    load_params<T, KIND>(jp, p0, p1, cl, seed);
    const uint32_t lane = threadIdx.x & 31;
    const int n_chunks = (int)ceil_div(vec_len, chunk_size);
    for (int64_t lr = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); lr < row_count; lr += (int64_t)gridDim.x * 8) {
        const uint32_t row = (uint32_t)(row_offset + lr);
        A acc = A(0);
        uint32_t cnt = 0;
        for (int c = 0; c < n_chunks; ++c) {
            const int64_t cs = (int64_t)c * chunk_size;
            const uint32_t width = (uint32_t)min(chunk_size, vec_len - cs);
            const uint32_t qmax = lane_qmax<32>(width, lane);
            uint32_t s = rng_init(seed, row, (uint32_t)c, lane);
            uint32_t q = rng_initial_q(s, cl);
            // column j = cs + lane + 32 q  ->  word (cs+lane)/32 + q, bit (cs+lane)%32: invariant shift
            const int64_t b0 = cs + lane;
            const uint32_t* wp = packed + (b0 >> 5);
            // opaque to the optimiser: it otherwise folds (b0 >> 5) + q into one 64-bit index and spends four
            // instructions per walk step on the address instead of one IMAD.WIDE on this pointer
            asm volatile("" : "+l"(wp));
            const uint32_t sh = (uint32_t)(b0 & 31);
            while (q < qmax) {
                const uint32_t bit = (__ldg(wp + q) >> sh) & 1u;
                if (KIND == BE_B200_JIT_SCALAR) cnt += bit;
                else if (bit) acc += jit_weight<KIND, A>(p0, p1, seed, row, (uint32_t)(b0 + 32 * (int64_t)q));
                s = rng_next(s);
                q += 1U + __umulhi(s, cl - 1U);
            }
        }
        if (KIND == BE_B200_JIT_SCALAR) {
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            acc = (A)cnt * p0;
        } else {
            acc = warp_sum(acc);
        }
        if (lane == 0) out[lr] = from_acc<T>(acc);
    }
}

// ------------------------------------------------------------------------------------------
// mv scatter (corder=False): for every active row, out[j] += w(row,j) for j in C(row)
// one warp per (active row, chunk) task.  [col_begin, col_end) is the owned output window.
// ------------------------------------------------------------------------------------------
// Normal weights: the central branch of the inverse CDF is evaluated inline; a draw that falls in a
// tail (4.85 %) is parked as (row, column) in a per-warp shared-memory queue and evaluated later by
// converged lanes.  Evaluating it in place made the warp run the tail path (logf, sqrt, two Horner
// chains, a division) in ~2 of 3 walk iterations for the sake of one lane.  Scatter order is free
// (reductions), so the result is the same sum.
constexpr int kTailCap = 128;       // entries per warp
constexpr int kTailFlush = 48;      // flush at a task boundary once this many are waiting

template <typename T, int KIND>
__global__ void __launch_bounds__(256) jit_mv_scatter_kernel(JitParams jp, const int32_t* __restrict__ ids,
                                                             const int* __restrict__ n_active_ptr,
                                                             int64_t out_len, int64_t chunk_size,
                                                             int64_t row_offset, int64_t row_count,
                                                             int64_t col_begin, int64_t col_end,
                                                             T* __restrict__ out) {
    using A = typename Acc<T>::type;
    constexpr bool DEFER = (KIND == BE_B200_JIT_NORMAL);
    A p0, p1; uint32_t cl, seed;
    load_params<T, KIND>(jp, p0, p1, cl, seed);
    const uint32_t lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    __shared__ uint2 s_tail[DEFER ? 8 : 1][DEFER ? kTailCap : 1];
    __shared__ int s_ntail[8];
    if (DEFER) {
        if (lane == 0) s_ntail[wib] = 0;
        __syncwarp();
    }
    auto flush_tails = [&]() {       // converged: every lane of the warp calls it
        __syncwarp();
        const int n = min(s_ntail[wib], kTailCap);
        for (int i = lane; i < n; i += 32) {
            const uint2 e = s_tail[DEFER ? wib : 0][DEFER ? i : 0];
            const float u = normal01_clamp_u(hash_uniform01(seed, e.x, e.y));
            red_add(out + ((int64_t)e.y - col_begin), p0 + (A)normal01_tail(u) * p1);
        }
        __syncwarp();
        if (lane == 0) s_ntail[wib] = 0;
        __syncwarp();
    };
    // only the chunks that overlap the owned column window are enumerated: with all chunks in the task index and a
    // grid stride that is a multiple of the chunk count, three quarters of the warps of a 4- or 8-way shard drew
    // nothing but foreign chunks while the rest did all the work (8-way: 31 us for an eighth of the events)
    const int c_lo = (int)(col_begin / chunk_size);
    const int c_hi = (int)((min(col_end, out_len) - 1) / chunk_size);
    const int n_chunks = c_hi - c_lo + 1;
    const int64_t n_tasks = (int64_t)(*n_active_ptr) * n_chunks;
    for (int64_t t = (int64_t)blockIdx.x * 8 + wib; t < n_tasks; t += (int64_t)gridDim.x * 8) {
        const int64_t a = t / n_chunks;
        const int c = c_lo + (int)(t - a * n_chunks);
        const int64_t r = __ldg(ids + a);
        if (r < row_offset || r >= row_offset + row_count) continue;
        const int64_t cs = (int64_t)c * chunk_size;
        const int64_t ce = min(cs + chunk_size, out_len);
        if (ce <= col_begin || cs >= col_end) continue;   // (cannot happen any more; kept as a guard)
        const uint32_t width = (uint32_t)(min(ce, col_end) - cs);   // stop at the window's end
        const uint32_t qmax = lane_qmax<32>(width, lane);
        const uint32_t row = (uint32_t)r;
        uint32_t s = rng_init(seed, row, (uint32_t)c, lane);
        uint32_t q = rng_initial_q(s, cl);
        const int64_t b0 = cs + lane;
        while (q < qmax) {
            const int64_t j = b0 + 32 * (int64_t)q;
            if (j >= col_begin) {
                if (DEFER) {
                    const float u = normal01_clamp_u(hash_uniform01(seed, row, (uint32_t)j));
                    if (!normal01_is_tail(u)) {
                        red_add(out + (j - col_begin), p0 + (A)normal01_central(u) * p1);
                    } else {
                        const int slot = atomicAdd(&s_ntail[wib], 1);
                        if (slot < kTailCap) s_tail[DEFER ? wib : 0][DEFER ? slot : 0] = make_uint2(row, (uint32_t)j);
                        else red_add(out + (j - col_begin), p0 + (A)normal01_tail(u) * p1);   // queue full: in place
                    }
                } else {
                    red_add(out + (j - col_begin), jit_weight<KIND, A>(p0, p1, seed, row, (uint32_t)j));
                }
            }
            s = rng_next(s);
            q += 1U + __umulhi(s, cl - 1U);
        }
        if (DEFER) {
            __syncwarp();
            if (s_ntail[wib] >= kTailFlush) flush_tails();
        }
    }
    if (DEFER) flush_tails();
}

// ------------------------------------------------------------------------------------------
// mm (stride 4).  Lanes 0..3 of a warp walk the four residue streams of one (row, chunk); every
// connection they find is broadcast and applied by all 32 lanes (= batch columns, CPL per lane).
// ------------------------------------------------------------------------------------------
template <typename T, int KIND, int CPL>
__global__ void __launch_bounds__(256) jit_mm_gather_kernel(JitParams jp, const uint32_t* __restrict__ packed,
                                                            int wpr, int64_t vec_len, int64_t out_len,
                                                            int64_t n, int64_t chunk_size,
                                                            T* __restrict__ out) {
    using A = typename Acc<T>::type;
    A p0, p1; uint32_t cl, seed;
    load_params<T, KIND>(jp, p0, p1, cl, seed);
    const uint32_t lane = threadIdx.x & 31;
    const int n_chunks = (int)ceil_div(vec_len, chunk_size);
    const int n_pass = (wpr + CPL - 1) / CPL;
    for (int64_t rr = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); rr < out_len; rr += (int64_t)gridDim.x * 8) {
        const uint32_t row = (uint32_t)rr;
        for (int pass = 0; pass < n_pass; ++pass) {
            const int wb = pass * CPL;
            A acc[CPL];
#pragma unroll
            for (int k = 0; k < CPL; ++k) acc[k] = A(0);
            for (int c = 0; c < n_chunks; ++c) {
                const int64_t cs = (int64_t)c * chunk_size;
                const uint32_t width = (uint32_t)min(chunk_size, vec_len - cs);
                uint32_t s = 0, q = 0, qmax = 0;
                if (lane < 4) {
                    qmax = lane_qmax<4>(width, lane);
                    s = rng_init(seed, row, (uint32_t)c, lane);
                    q = rng_initial_q(s, cl);
                }
                uint32_t live = __ballot_sync(0xffffffffu, q < qmax);
                while (live) {
                    // each live walker lane holds one connection this round
                    int64_t my_j = 0;
                    A my_w = A(0);
                    uint32_t mw[CPL];
#pragma unroll
                    for (int k = 0; k < CPL; ++k) mw[k] = 0u;
                    if (q < qmax) {
                        my_j = cs + lane + 4 * (int64_t)q;
                        const uint32_t* brow = packed + my_j * wpr + wb;
#pragma unroll
                        for (int k = 0; k < CPL; ++k) mw[k] = (wb + k < wpr) ? __ldg(brow + k) : 0u;
                        my_w = jit_weight<KIND, A>(p0, p1, seed, row, (uint32_t)my_j);
                        s = rng_next(s);
                        q += 1U + __umulhi(s, cl - 1U);
                    }
                    uint32_t m = live;
                    while (m) {
                        const int t = __ffs(m) - 1;
                        m &= m - 1;
                        const A wv = __shfl_sync(0xffffffffu, my_w, t);
#pragma unroll
                        for (int k = 0; k < CPL; ++k) {
                            const uint32_t bits = __shfl_sync(0xffffffffu, mw[k], t);
                            if ((bits >> lane) & 1u) acc[k] += wv;
                        }
                    }
                    live = __ballot_sync(0xffffffffu, q < qmax);
                }
            }
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                const int64_t col = (int64_t)(wb + k) * 32 + lane;
                if (col < n) out[rr * n + col] = from_acc<T>(acc[k]);
            }
        }
    }
}

template <typename T, int KIND>
__global__ void __launch_bounds__(256) jit_mm_scatter_kernel(JitParams jp, const int32_t* __restrict__ ids,
                                                             const int* __restrict__ n_active_ptr,
                                                             const uint32_t* __restrict__ packed, int wpr,
                                                             int64_t out_len, int64_t n,
                                                             int64_t chunk_size, T* __restrict__ out) {
    using A = typename Acc<T>::type;
    A p0, p1; uint32_t cl, seed;
    load_params<T, KIND>(jp, p0, p1, cl, seed);
    const uint32_t lane = threadIdx.x & 31;
    const int n_chunks = (int)ceil_div(out_len, chunk_size);
    const int64_t n_tasks = (int64_t)(*n_active_ptr) * n_chunks;
    for (int64_t t = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); t < n_tasks; t += (int64_t)gridDim.x * 8) {
        const int64_t a = t / n_chunks;
        const int c = (int)(t - a * n_chunks);
        const int64_t r = __ldg(ids + a);
        const uint32_t row = (uint32_t)r;
        const int64_t cs = (int64_t)c * chunk_size;
        const uint32_t width = (uint32_t)min(chunk_size, out_len - cs);
        uint32_t s = 0, q = 0, qmax = 0;
        if (lane < 4) {
            qmax = lane_qmax<4>(width, lane);
            s = rng_init(seed, row, (uint32_t)c, lane);
            q = rng_initial_q(s, cl);
        }
        const uint32_t* mrow = packed + r * wpr;
        uint32_t live = __ballot_sync(0xffffffffu, q < qmax);
        while (live) {
            int64_t my_j = 0;
            A my_w = A(0);
            if (q < qmax) {
                my_j = cs + lane + 4 * (int64_t)q;
                my_w = jit_weight<KIND, A>(p0, p1, seed, row, (uint32_t)my_j);
                s = rng_next(s);
                q += 1U + __umulhi(s, cl - 1U);
            }
            uint32_t m = live;
            while (m) {
                const int tt = __ffs(m) - 1;
                m &= m - 1;
                const A wv = __shfl_sync(0xffffffffu, my_w, tt);
                const int64_t j = __shfl_sync(0xffffffffu, my_j, tt);
                for (int k = 0; k < wpr; ++k) {
                    const uint32_t bits = __ldg(mrow + k);
                    const int64_t col = (int64_t)k * 32 + lane;
                    if ((bits >> lane) & 1u) red_add(out + j * n + col, wv);
                }
            }
            live = __ballot_sync(0xffffffffu, q < qmax);
        }
    }
}

// ------------------------------------------------------------------------------------------
// float (non-event) twins (ref _jit_scalar/float.py:805-912, :1298-1405 and the
// normal/uniform analogues): same walks, a value vector / matrix instead of spike bits.
// ------------------------------------------------------------------------------------------
template <typename T, int KIND>
__global__ void __launch_bounds__(256) jit_mv_val_gather_kernel(JitParams jp, const T* __restrict__ x,
                                                                int64_t vec_len, int64_t chunk_size,
                                                                int64_t out_len, T* __restrict__ out) {
    using A = typename Acc<T>::type;
    A p0, p1; uint32_t cl, seed;
    load_params<T, KIND>(jp, p0, p1, cl, seed);
    const uint32_t lane = threadIdx.x & 31;
    const int n_chunks = (int)ceil_div(vec_len, chunk_size);
    for (int64_t rr = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); rr < out_len; rr += (int64_t)gridDim.x * 8) {
        const uint32_t row = (uint32_t)rr;
        A acc = A(0);
        for (int c = 0; c < n_chunks; ++c) {
            const int64_t cs = (int64_t)c * chunk_size;
            const uint32_t width = (uint32_t)min(chunk_size, vec_len - cs);
            const uint32_t qmax = lane_qmax<32>(width, lane);
            uint32_t s = rng_init(seed, row, (uint32_t)c, lane);
            uint32_t q = rng_initial_q(s, cl);
            const T* xp = x + cs + lane;
            while (q < qmax) {
                const A xv = to_acc<T>(__ldg(xp + 32 * (int64_t)q));
                if (KIND == BE_B200_JIT_SCALAR) acc += xv;
                else acc += jit_weight<KIND, A>(p0, p1, seed, row, (uint32_t)(cs + lane + 32 * (int64_t)q)) * xv;
                s = rng_next(s);
                q += 1U + __umulhi(s, cl - 1U);
            }
        }
        acc = warp_sum(acc);
        if (KIND == BE_B200_JIT_SCALAR) acc *= p0;
        if (lane == 0) out[rr] = from_acc<T>(acc);
    }
}

template <typename T, int KIND>
__global__ void __launch_bounds__(256) jit_mv_val_scatter_kernel(JitParams jp, const T* __restrict__ x,
                                                                 int64_t vec_len, int64_t out_len,
                                                                 int64_t chunk_size, T* __restrict__ out) {
    using A = typename Acc<T>::type;
    constexpr bool DEFER = (KIND == BE_B200_JIT_NORMAL);     // tails of the inverse CDF: see jit_mv_scatter_kernel
    A p0, p1; uint32_t cl, seed;
    load_params<T, KIND>(jp, p0, p1, cl, seed);
    const uint32_t lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    __shared__ uint2 s_tail[DEFER ? 8 : 1][DEFER ? kTailCap : 1];
    __shared__ int s_ntail[8];
    if (DEFER) {
        if (lane == 0) s_ntail[wib] = 0;
        __syncwarp();
    }
    auto flush_tails = [&]() {       // converged: every lane of the warp calls it
        __syncwarp();
        const int n = min(s_ntail[wib], kTailCap);
        for (int i = lane; i < n; i += 32) {
            const uint2 e = s_tail[DEFER ? wib : 0][DEFER ? i : 0];
            const float u = normal01_clamp_u(hash_uniform01(seed, e.x, e.y));
            red_add(out + e.y, (p0 + (A)normal01_tail(u) * p1) * to_acc<T>(__ldg(x + e.x)));
        }
        __syncwarp();
        if (lane == 0) s_ntail[wib] = 0;
        __syncwarp();
    };
    const int n_chunks = (int)ceil_div(out_len, chunk_size);
    const int64_t n_tasks = vec_len * n_chunks;
    for (int64_t t = (int64_t)blockIdx.x * 8 + wib; t < n_tasks; t += (int64_t)gridDim.x * 8) {
        const int64_t r = t / n_chunks;
        const int c = (int)(t - r * n_chunks);
        const A xv = to_acc<T>(__ldg(x + r));
        if (xv == A(0)) continue;     // ref: `if v == 0.: continue`
        const int64_t cs = (int64_t)c * chunk_size;
        const uint32_t width = (uint32_t)min(chunk_size, out_len - cs);
        const uint32_t qmax = lane_qmax<32>(width, lane);
        const uint32_t row = (uint32_t)r;
        uint32_t s = rng_init(seed, row, (uint32_t)c, lane);
        uint32_t q = rng_initial_q(s, cl);
        const int64_t b0 = cs + lane;
        while (q < qmax) {
            const int64_t j = b0 + 32 * (int64_t)q;
            if (DEFER) {
                const float u = normal01_clamp_u(hash_uniform01(seed, row, (uint32_t)j));
                if (!normal01_is_tail(u)) {
                    red_add(out + j, (p0 + (A)normal01_central(u) * p1) * xv);
                } else {
                    const int slot = atomicAdd(&s_ntail[wib], 1);
                    if (slot < kTailCap) s_tail[DEFER ? wib : 0][DEFER ? slot : 0] = make_uint2(row, (uint32_t)j);
                    else red_add(out + j, (p0 + (A)normal01_tail(u) * p1) * xv);   // queue full: in place
                }
            } else {
                red_add(out + j, jit_weight<KIND, A>(p0, p1, seed, row, (uint32_t)j) * xv);
            }
            s = rng_next(s);
            q += 1U + __umulhi(s, cl - 1U);
        }
        if (DEFER) {
            __syncwarp();
            if (s_ntail[wib] >= kTailFlush) flush_tails();
        }
    }
    if (DEFER) flush_tails();
}

// mm: lanes 0..3 walk the four residue streams of one (row, chunk), every connection is broadcast
// and applied by all lanes (= 32 columns of X / out, contiguous segments).
template <typename T, int KIND, bool GATHER>
__global__ void __launch_bounds__(256) jit_mm_val_kernel(JitParams jp, const T* __restrict__ X,
                                                         int64_t n_rows, int64_t walk_len, int64_t n,
                                                         int64_t chunk_size, T* __restrict__ out) {
    using A = typename Acc<T>::type;
    A p0, p1; uint32_t cl, seed;
    load_params<T, KIND>(jp, p0, p1, cl, seed);
    const uint32_t lane = threadIdx.x & 31;
    const int n_chunks = (int)ceil_div(walk_len, chunk_size);
    const int64_t n_ct = ceil_div(n, (int64_t)32);
    // gather: one warp owns (row, column tile) and loops over the chunks; scatter: (row, chunk, tile)
    const int64_t n_tasks = GATHER ? n_rows * n_ct : n_rows * n_ct * n_chunks;
    for (int64_t t = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); t < n_tasks; t += (int64_t)gridDim.x * 8) {
        int64_t rr, ct;
        int c_lo, c_hi;
        if (GATHER) { rr = t / n_ct; ct = t - rr * n_ct; c_lo = 0; c_hi = n_chunks; }
        else {
            const int64_t rc = t / n_ct;
            ct = t - rc * n_ct;
            rr = rc / n_chunks;
            c_lo = (int)(rc - rr * n_chunks);
            c_hi = c_lo + 1;
        }
        const uint32_t row = (uint32_t)rr;
        const int64_t col = ct * 32 + lane;
        const bool ok = col < n;
        A acc = A(0);
        A xrow = A(0);
        if (!GATHER && ok) xrow = to_acc<T>(__ldg(X + rr * n + col));
        for (int c = c_lo; c < c_hi; ++c) {
            const int64_t cs = (int64_t)c * chunk_size;
            const uint32_t width = (uint32_t)min(chunk_size, walk_len - cs);
            uint32_t s = 0, q = 0, qmax = 0;
            if (lane < 4) {
                qmax = lane_qmax<4>(width, lane);
                s = rng_init(seed, row, (uint32_t)c, lane);
                q = rng_initial_q(s, cl);
            }
            uint32_t live = __ballot_sync(0xffffffffu, q < qmax);
            while (live) {
                int64_t my_j = 0;
                A my_w = A(1);
                if (q < qmax) {
                    my_j = cs + lane + 4 * (int64_t)q;
                    if (KIND != BE_B200_JIT_SCALAR || !GATHER) my_w = jit_weight<KIND, A>(p0, p1, seed, row, (uint32_t)my_j);
                    s = rng_next(s);
                    q += 1U + __umulhi(s, cl - 1U);
                }
                uint32_t m = live;
                while (m) {
                    const int tt = __ffs(m) - 1;
                    m &= m - 1;
                    const A wv = __shfl_sync(0xffffffffu, my_w, tt);
                    const int64_t j = __shfl_sync(0xffffffffu, my_j, tt);
                    if (ok) {
                        if (GATHER) acc += wv * to_acc<T>(__ldg(X + j * n + col));
                        else if (xrow != A(0)) red_add(out + j * n + col, xrow * wv);
                    }
                }
                live = __ballot_sync(0xffffffffu, q < qmax);
            }
        }
        if (GATHER && ok) out[rr * n + col] = from_acc<T>(KIND == BE_B200_JIT_SCALAR ? acc * p0 : acc);
    }
}

// ------------------------------------------------------------------------------------------
// connectivity dump (parity instrument): per generated row, count and the first `cap` columns
// in the reference's walk order (chunk-major, then lane, then ascending within the lane).
// ------------------------------------------------------------------------------------------
template <int S>
__global__ void jit_conn_dump_kernel(const int32_t* __restrict__ clen_p, const int32_t* __restrict__ seed_p,
                                     int64_t row_begin, int64_t n_rows, int64_t k, int64_t chunk_size,
                                     int32_t* __restrict__ counts, int32_t* __restrict__ cols, int64_t cap) {
    const int64_t lr = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (lr >= n_rows) return;
    uint32_t cl = (uint32_t)__ldg(clen_p);
    if (cl < 2U) cl = 2U;
    const uint32_t seed = (uint32_t)__ldg(seed_p);
    const uint32_t row = (uint32_t)(row_begin + lr);
    const int n_chunks = (int)ceil_div(k, chunk_size);
    int64_t cnt = 0;
    for (int c = 0; c < n_chunks; ++c) {
        const int64_t cs = (int64_t)c * chunk_size;
        const uint32_t width = (uint32_t)min(chunk_size, k - cs);
        for (uint32_t lane = 0; lane < (uint32_t)S; ++lane) {
            const uint32_t qmax = lane_qmax<S>(width, lane);
            uint32_t s = rng_init(seed, row, (uint32_t)c, lane);
            uint32_t q = rng_initial_q(s, cl);
            while (q < qmax) {
                if (cols != nullptr && cnt < cap) cols[lr * cap + cnt] = (int32_t)(cs + lane + (int64_t)S * q);
                ++cnt;
                s = rng_next(s);
                q += 1U + __umulhi(s, cl - 1U);
            }
        }
    }
    counts[lr] = (int32_t)cnt;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <typename T, int KIND>
static int run_jit_mv(const JitParams& jp, const Events& ev, int64_t vec_len, int64_t out_len,
                      int64_t chunk_size, bool corder, int64_t row_offset, int64_t row_count,
                      int64_t col_begin, int64_t col_end, T* out, cudaStream_t s) {
    if (corder) {
        if (row_count == 0) return BE_B200_OK;
        int blocks = (int)std::min<int64_t>(ceil_div(row_count, 8), (int64_t)kNumSM * 32);
        {
            ProfileScope prof(s);
            jit_mv_gather_kernel<T, KIND><<<blocks, 256, 0, s>>>(jp, ev.packed, vec_len, chunk_size,
                                                                row_offset, row_count, out);
        }
        BE_LAUNCH_CHECK("jit_mv_gather_kernel");
    } else {
        const int64_t own = col_end - col_begin;
        if (own <= 0 || vec_len == 0) return BE_B200_OK;   // output already cleared with the compaction
        const int64_t n_chunks = ceil_div(out_len, chunk_size);
        int blocks = (int)std::min<int64_t>(ceil_div(std::max<int64_t>(vec_len * n_chunks, 1), 8), (int64_t)kNumSM * 8);
        {
            ProfileScope prof(s);
            jit_mv_scatter_kernel<T, KIND><<<blocks, 256, 0, s>>>(jp, ev.active_ids, ev.n_active, out_len,
                                                                 chunk_size, row_offset, row_count,
                                                                 col_begin, col_end, out);
        }
        BE_LAUNCH_CHECK("jit_mv_scatter_kernel");
    }
    return BE_B200_OK;
}

template <typename T, int KIND>
static int run_jit_mm(const JitParams& jp, const Events& ev, int64_t vec_len, int64_t out_len,
                      int64_t n, int64_t chunk_size, bool corder, T* out, cudaStream_t s) {
    const int wpr = ev.words_per_row;
    if (corder) {
        int blocks = (int)std::min<int64_t>(ceil_div(out_len, 8), (int64_t)kNumSM * 32);
#define BE_JG(CPL) jit_mm_gather_kernel<T, KIND, CPL><<<blocks, 256, 0, s>>>(jp, ev.packed, wpr, vec_len, out_len, n, chunk_size, out)
        {
            ProfileScope prof(s);
            if (wpr == 1) BE_JG(1); else if (wpr == 2) BE_JG(2); else BE_JG(4);
        }
#undef BE_JG
        BE_LAUNCH_CHECK("jit_mm_gather_kernel");
    } else {
        BE_CUDA(cudaMemsetAsync(out, 0, (size_t)(out_len * n) * sizeof(T), s));
        if (vec_len == 0) return BE_B200_OK;
        const int64_t n_chunks = ceil_div(out_len, chunk_size);
        int blocks = (int)std::min<int64_t>(ceil_div(std::max<int64_t>(vec_len * n_chunks, 1), 8), (int64_t)kNumSM * 8);
        {
            ProfileScope prof(s);
            jit_mm_scatter_kernel<T, KIND><<<blocks, 256, 0, s>>>(jp, ev.active_ids, ev.n_active, ev.packed,
                                                                 wpr, out_len, n, chunk_size, out);
        }
        BE_LAUNCH_CHECK("jit_mm_scatter_kernel");
    }
    return BE_B200_OK;
}


template <typename T, int KIND>
static int run_jit_val(const JitParams& jp, const T* x, int64_t vec_len, int64_t out_len, int64_t n,
                       int64_t chunk_size, bool corder, T* out, cudaStream_t s) {
    if (out_len == 0) return BE_B200_OK;
    const int64_t width = std::max<int64_t>(n, 1);
    if (!corder || vec_len == 0) BE_CUDA(cudaMemsetAsync(out, 0, (size_t)(out_len * width) * sizeof(T), s));
    if (vec_len == 0) return BE_B200_OK;
    const int64_t walk = corder ? vec_len : out_len;
    const int64_t n_chunks = ceil_div(walk, chunk_size);
    const int64_t n_ct = ceil_div(width, (int64_t)32);
    int64_t tasks;
    if (n == 0) tasks = corder ? out_len : vec_len * n_chunks;
    else tasks = corder ? out_len * n_ct : vec_len * n_chunks * n_ct;
    const int blocks = (int)std::min<int64_t>(ceil_div(std::max<int64_t>(tasks, 1), 8), (int64_t)kNumSM * 32);
    {
        ProfileScope prof(s);
        if (n == 0) {
            if (corder) jit_mv_val_gather_kernel<T, KIND><<<blocks, 256, 0, s>>>(jp, x, vec_len, chunk_size, out_len, out);
            else jit_mv_val_scatter_kernel<T, KIND><<<blocks, 256, 0, s>>>(jp, x, vec_len, out_len, chunk_size, out);
        } else {
            if (corder) jit_mm_val_kernel<T, KIND, true><<<blocks, 256, 0, s>>>(jp, x, out_len, vec_len, n, chunk_size, out);
            else jit_mm_val_kernel<T, KIND, false><<<blocks, 256, 0, s>>>(jp, x, vec_len, out_len, n, chunk_size, out);
        }
    }
    BE_LAUNCH_CHECK("jit_val_kernel");
    return BE_B200_OK;
}

static int jit_val_entry(int kind, const void* p0, const void* p1, int w_dtype, const int32_t* clen,
                         const int32_t* seed, const void* x, int64_t vec_len, int64_t out_len, int64_t n,
                         int64_t chunk_size, int corder, void* out, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    BE_REQUIRE(p0 && clen && seed && out, "jit: NULL pointer");
    BE_REQUIRE(kind == BE_B200_JIT_SCALAR || p1, "jit: p1 is NULL");
    BE_REQUIRE(x || vec_len == 0, "jit: operand is NULL");
    BE_REQUIRE(vec_len >= 0 && out_len >= 0 && vec_len < (1ll << 31) && out_len < (1ll << 31), "jit: shape out of range");
    BE_REQUIRE(chunk_size >= 1, "jit: chunk_size must be positive");
    JitParams jp{p0, p1, clen, seed};
    BE_DISPATCH_WT(w_dtype, T, {
        BE_DISPATCH_KIND(kind, K, {
            return run_jit_val<T, K>(jp, (const T*)x, vec_len, out_len, n, chunk_size, corder != 0, (T*)out, s);
        });
    });
    return BE_B200_OK;
}

}  // namespace be

using namespace be;

extern "C" int be_b200_binary_jitmv(int kind, const void* p0, const void* p1, int w_dtype,
                                    const int32_t* clen, const int32_t* seed, const void* spikes,
                                    const void* aux, int spk_kind, int64_t vec_len, int64_t out_len,
                                    int64_t chunk_size, int corder, int64_t row_offset,
                                    int64_t row_count, int64_t col_begin, int64_t col_end, void* out,
                                    void* workspace, size_t workspace_bytes, void* stream) { BE_NVTX_RANGE("be_b200_binary_jitmv");
    cudaStream_t s = (cudaStream_t)stream;
    BE_REQUIRE(p0 && clen && seed && out, "jitmv: NULL pointer");
    BE_REQUIRE(kind == BE_B200_JIT_SCALAR || p1, "jitmv: p1 is NULL");
    BE_REQUIRE(vec_len >= 0 && out_len >= 0 && vec_len < (1ll << 31) && out_len < (1ll << 31), "jitmv: shape out of range");
    BE_REQUIRE(chunk_size >= 1, "jitmv: chunk_size must be positive");
    const int64_t n_rows_total = corder ? out_len : vec_len;
    if (row_count < 0) row_count = n_rows_total - row_offset;
    if (col_end < 0) col_end = out_len;
    BE_REQUIRE(row_offset >= 0 && row_offset + row_count <= n_rows_total, "jitmv: row window out of range");
    BE_REQUIRE(col_begin >= 0 && col_begin <= col_end && col_end <= out_len, "jitmv: column window out of range");
    BE_REQUIRE(!corder || (col_begin == 0 && col_end == out_len), "jitmv: column window only applies to corder=0");
    Events ev;
    static const size_t kWt[4] = {4, 8, 2, 2};
    BE_REQUIRE(w_dtype >= 0 && w_dtype < 4, "jitmv: unsupported weight dtype %d", w_dtype);
    int rc = prepare_events_1d(spikes, aux, spk_kind, vec_len, !corder, corder != 0, workspace, workspace_bytes, s, &ev,
                               corder ? nullptr : out, corder ? 0 : (size_t)(col_end - col_begin) * kWt[w_dtype],
                               /*ordered_ids=*/false);
    if (rc) return rc;
    JitParams jp{p0, p1, clen, seed};
    BE_DISPATCH_WT(w_dtype, T, {
        BE_DISPATCH_KIND(kind, K, {
            return run_jit_mv<T, K>(jp, ev, vec_len, out_len, chunk_size, corder != 0, row_offset, row_count, col_begin, col_end, (T*)out, s);
        });
    });
    return BE_B200_OK;
}

extern "C" int be_b200_binary_jitmm(int kind, const void* p0, const void* p1, int w_dtype,
                                    const int32_t* clen, const int32_t* seed, const void* spikes,
                                    const void* aux, int spk_kind, int64_t vec_len, int64_t out_len,
                                    int64_t n, int64_t chunk_size, int corder, void* out,
                                    void* workspace, size_t workspace_bytes, void* stream) { BE_NVTX_RANGE("be_b200_binary_jitmm");
    cudaStream_t s = (cudaStream_t)stream;
    BE_REQUIRE(p0 && clen && seed && out, "jitmm: NULL pointer");
    BE_REQUIRE(kind == BE_B200_JIT_SCALAR || p1, "jitmm: p1 is NULL");
    BE_REQUIRE(vec_len >= 0 && out_len >= 0 && vec_len < (1ll << 31) && out_len < (1ll << 31) && n >= 1, "jitmm: shape out of range");
    BE_REQUIRE(chunk_size >= 1, "jitmm: chunk_size must be positive");
    if (out_len == 0) return BE_B200_OK;
    Events ev;
    int rc = prepare_events_2d(spikes, aux, spk_kind, vec_len, n, !corder, workspace, workspace_bytes, s, &ev);
    if (rc) return rc;
    JitParams jp{p0, p1, clen, seed};
    BE_DISPATCH_WT(w_dtype, T, {
        BE_DISPATCH_KIND(kind, K, {
            return run_jit_mm<T, K>(jp, ev, vec_len, out_len, n, chunk_size, corder != 0, (T*)out, s);
        });
    });
    return BE_B200_OK;
}

extern "C" int be_b200_jitc_row_connections(const int32_t* clen, const int32_t* seed,
                                            int64_t row_begin, int64_t n_rows, int64_t k,
                                            int64_t chunk_size, int stride, int32_t* counts,
                                            int32_t* cols, int64_t cap, void* stream) { BE_NVTX_RANGE("be_b200_jitc_row_connections");
    cudaStream_t s = (cudaStream_t)stream;
    BE_REQUIRE(clen && seed && counts, "jitc_row_connections: NULL pointer");
    BE_REQUIRE(stride == 32 || stride == 4, "jitc_row_connections: stride must be 32 (mv) or 4 (mm)");
    BE_REQUIRE(chunk_size >= 1 && k >= 0 && n_rows >= 0, "jitc_row_connections: bad sizes");
    if (n_rows == 0) return BE_B200_OK;
    int blocks = (int)ceil_div(n_rows, 128);
    if (stride == 32) jit_conn_dump_kernel<32><<<blocks, 128, 0, s>>>(clen, seed, row_begin, n_rows, k, chunk_size, counts, cols, cap);
    else jit_conn_dump_kernel<4><<<blocks, 128, 0, s>>>(clen, seed, row_begin, n_rows, k, chunk_size, counts, cols, cap);
    BE_LAUNCH_CHECK("jit_conn_dump_kernel");
    return BE_B200_OK;
}

extern "C" int be_b200_jitmv(int kind, const void* p0, const void* p1, int w_dtype, const int32_t* clen,
                             const int32_t* seed, const void* x, int64_t vec_len, int64_t out_len,
                             int64_t chunk_size, int corder, void* out, void* stream) { BE_NVTX_RANGE("be_b200_jitmv");
    return jit_val_entry(kind, p0, p1, w_dtype, clen, seed, x, vec_len, out_len, 0, chunk_size, corder, out, stream);
}

extern "C" int be_b200_jitmm(int kind, const void* p0, const void* p1, int w_dtype, const int32_t* clen,
                             const int32_t* seed, const void* X, int64_t vec_len, int64_t out_len, int64_t n,
                             int64_t chunk_size, int corder, void* out, void* stream) { BE_NVTX_RANGE("be_b200_jitmm");
    BE_REQUIRE(n >= 1, "jitmm: n must be >= 1");
    return jit_val_entry(kind, p0, p1, w_dtype, clen, seed, X, vec_len, out_len, n, chunk_size, corder, out, stream);
}
