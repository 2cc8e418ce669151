/*
 * be_oracle.c -- CPU restatement of brainevent's event-driven synaptic propagation path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under brainevent_b200/ may link, import or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * leg use it, as the checker / CPU baseline, never as the product path.
 *
 * Every function cites the reference file:line (relative to /root/reference/brainevent) whose
 * loop structure and accumulation order it restates.  The reference kernels are numba-jitted
 * Python; this file is the same arithmetic in scalar C (single thread, like the reference's
 * default `parallel=False`, config.py:41).  It is pinned against the reference itself by
 * tests/test_oracle_vs_reference.py (run in the build container through oracle/ref_shim.py)
 * and against the reference's own known-answer vectors (tests/test_oracle_golden.py).
 *
 * Spike activity convention: callers pass `act` as uint8 (0 / non-zero).  The Python wrapper
 * (oracle/__init__.py) applies the reference's predicate first: bool -> as is, float -> `> 0`
 * for the operators (_csr/binary.py:413), `!= 0` for compaction (_event/compact.py:327).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * light-RNG  (_numba_random.py:385-502  ==  _jit_scalar/binary_jitsmv.cu:38-105)
 * ---------------------------------------------------------------------------------------- */
uint32_t ora_mix32(uint32_t x) { /* _numba_random.py:385-393 */
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
uint32_t ora_rng_bounded(uint32_t r, uint32_t bound) { /* :396-398 */
    return (uint32_t)(((uint64_t)r * (uint64_t)bound) >> 32);
}
uint32_t ora_rng_next(uint32_t x) { /* :401-409 xorshift32 */
    x ^= x << 13; x ^= x >> 17; x ^= x << 5;
    return x == 0U ? 0x6d2b79f5U : x;
}
uint32_t ora_rng_init(uint32_t seed, uint32_t row, uint32_t chunk, uint32_t lane) { /* :412-421 */
    uint32_t x = seed ^ 0xd1b54a35U;
    x ^= row * 0x85ebca6bU; x ^= chunk * 0xc2b2ae35U; x ^= lane * 0x27d4eb2dU;
    x = ora_mix32(x);
    return x == 0U ? 0x6d2b79f5U : x;
}
/* :489-502  stationary residual start; returns q, writes the advanced state */
uint32_t ora_rng_initial_q(uint32_t* state, uint32_t cl) {
    uint32_t n = cl - 1U, s = *state;
    for (;;) {
        s = ora_rng_next(s); uint32_t q = ora_rng_bounded(s, n);
        s = ora_rng_next(s); uint32_t g = ora_rng_bounded(s, n);
        if (g < n - q) { *state = s; return q; }
    }
}
float ora_uniform01(uint32_t seed, uint32_t row, uint32_t col) { /* :424-430 */
    uint32_t h = seed ^ 0xa0761d65U;
    h ^= row * 0xe7037ed1U; h ^= col * 0x8ebc6af1U;
    h = ora_mix32(h);
    return (float)(h & 0x00ffffffU) * (1.0f / 16777216.0f);
}
float ora_normal01(uint32_t seed, uint32_t row, uint32_t col) { /* :433-486 Acklam, f32 */
    float u = ora_uniform01(seed, row, col);
    const float lo = 1e-10f, hi = 1.0f - 1e-10f;
    if (u < lo) u = lo; else if (u > hi) u = hi;
    const float a1 = -39.696830f, a2 = 220.94609f, a3 = -275.92851f, a4 = 138.35775f,
                a5 = -30.664799f, a6 = 2.5066283f;
    const float b1 = -54.476099f, b2 = 161.58584f, b3 = -155.69898f, b4 = 66.801312f,
                b5 = -13.280681f;
    const float c1 = -0.007784894f, c2 = -0.32239646f, c3 = -2.4007583f, c4 = -2.5497325f,
                c5 = 4.3746641f, c6 = 2.9381640f;
    const float d1 = 0.007784696f, d2 = 0.32246713f, d3 = 2.4451342f, d4 = 3.7544087f;
    float z;
    if (u < 0.02425f) {
        float v = sqrtf(-2.0f * logf(u));
        z = -((((((c1 * v + c2) * v + c3) * v + c4) * v + c5) * v + c6) /
              ((((d1 * v + d2) * v + d3) * v + d4) * v + 1.0f));
    } else if (u > 0.97575f) {
        float v = sqrtf(-2.0f * logf(1.0f - u));
        z = (((((c1 * v + c2) * v + c3) * v + c4) * v + c5) * v + c6) /
            ((((d1 * v + d2) * v + d3) * v + d4) * v + 1.0f);
    } else {
        float v = u - 0.5f, r = v * v;
        z = (((((a1 * r + a2) * r + a3) * r + a4) * r + a5) * r + a6) * v /
            (((((b1 * r + b2) * r + b3) * r + b4) * r + b5) * r + 1.0f);
    }
    return z;
}

/* ------------------------------------------------------------------------------------------
 * The JITC walk (_jit_scalar/binary.py:352-377,392-416).
 * Visits every connection (row, j) of one generated row in the reference's order:
 * chunk-major, then lane, then ascending within the lane.
 * ---------------------------------------------------------------------------------------- */
#define ORA_WALK_ROW(seed0, cl, row, k, chunk_size, stride, BODY)                              \
    do {                                                                                        \
        int64_t n_chunks_ = ((k) + (chunk_size) - 1) / (chunk_size);                            \
        for (int64_t chunk_id_ = 0; chunk_id_ < n_chunks_; ++chunk_id_) {                       \
            int64_t chunk_start_ = chunk_id_ * (chunk_size);                                    \
            if (chunk_start_ >= (k)) break;                                                     \
            int64_t chunk_end_ = chunk_start_ + (chunk_size);                                   \
            if (chunk_end_ > (k)) chunk_end_ = (k);                                             \
            int64_t chunk_width_ = chunk_end_ - chunk_start_;                                   \
            for (int lane_ = 0; lane_ < (stride); ++lane_) {                                    \
                uint32_t st_ = ora_rng_init((seed0), (uint32_t)(row), (uint32_t)chunk_id_,      \
                                            (uint32_t)lane_);                                   \
                uint32_t q_ = ora_rng_initial_q(&st_, (cl));                                    \
                int64_t local_j_ = lane_ + (int64_t)(stride) * (int64_t)q_;                     \
                while (local_j_ < chunk_width_) {                                               \
                    int64_t j = chunk_start_ + local_j_;                                        \
                    BODY;                                                                       \
                    st_ = ora_rng_next(st_);                                                    \
                    q_ = q_ + 1U + ora_rng_bounded(st_, (cl) - 1U);                             \
                    local_j_ = lane_ + (int64_t)(stride) * (int64_t)q_;                         \
                }                                                                               \
            }                                                                                   \
        }                                                                                       \
    } while (0)

/* Connectivity dump of one generated row: columns in walk order.  Returns the count (may
 * exceed cap; only the first cap entries are written).  Used for index-for-index parity. */
int64_t ora_jitc_row_conn(uint32_t seed0, uint32_t clen, int64_t row, int64_t k,
                          int64_t chunk_size, int stride, int64_t* cols, int64_t cap) {
    uint32_t cl = clen < 2U ? 2U : clen;
    int64_t cnt = 0;
    ORA_WALK_ROW(seed0, cl, row, k, chunk_size, stride, {
        if (cnt < cap) cols[cnt] = j;
        ++cnt;
    });
    return cnt;
}

/* ------------------------------------------------------------------------------------------
 * bit-pack / compaction   (_event/bitpack_binary.py:32-76, _event/compact.py:158-199,
 *                          305-335, 451-503)
 * ---------------------------------------------------------------------------------------- */
void ora_bitpack_1d(const uint8_t* act, int64_t n, uint32_t* words) {
    int64_t nw = (n + 31) / 32;
    for (int64_t w = 0; w < nw; ++w) {
        uint32_t bits = 0;
        for (int b = 0; b < 32; ++b) {
            int64_t i = w * 32 + b;
            if (i < n && act[i]) bits |= (1U << b);
        }
        words[w] = bits;
    }
}
/* ascending ids, zero tail (the jax_raw convention, compact.py:338-358) */
void ora_compact_1d(const uint8_t* act, int64_t n, int32_t* ids, int32_t* n_active) {
    int64_t na = 0;
    for (int64_t i = 0; i < n; ++i) if (act[i]) ids[na++] = (int32_t)i;
    for (int64_t i = na; i < n; ++i) ids[i] = 0;
    n_active[0] = (int32_t)na;
}
/* 2-D: pack along the batch axis, a row is active if any column is (compact.py:451-503);
 * packed may be NULL for the compact-only primitive (compact.py:158-199). */
void ora_compact_2d(const uint8_t* act, int64_t n_pre, int64_t n_batch, uint32_t* packed,
                    int32_t* ids, int32_t* n_active) {
    int64_t nbw = (n_batch + 31) / 32, na = 0;
    for (int64_t i = 0; i < n_pre; ++i) {
        int row_active = 0;
        for (int64_t w = 0; w < nbw; ++w) {
            uint32_t word = 0;
            for (int b = 0; b < 32; ++b) {
                int64_t c = w * 32 + b;
                if (c < n_batch && act[i * n_batch + c]) word |= (1U << b);
            }
            if (packed) packed[i * nbw + w] = word;
            if (word) row_active = 1;
        }
        if (row_active) ids[na++] = (int32_t)i;
    }
    for (int64_t i = na; i < n_pre; ++i) ids[i] = 0;
    n_active[0] = (int32_t)na;
}

/* CSC mirror: stable argsort by column (_misc.py:1341-1384). */
void ora_csr_to_csc(const int64_t* indptr, const int32_t* indices, int64_t m, int64_t n_post,
                    int64_t* csc_indptr, int32_t* csc_indices, int64_t* perm) {
    int64_t nse = indptr[m];
    memset(csc_indptr, 0, sizeof(int64_t) * (size_t)(n_post + 1));
    for (int64_t j = 0; j < nse; ++j) csc_indptr[indices[j] + 1]++;
    for (int64_t c = 0; c < n_post; ++c) csc_indptr[c + 1] += csc_indptr[c];
    int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n_post > 0 ? n_post : 1));
    memcpy(cur, csc_indptr, sizeof(int64_t) * (size_t)n_post);
    for (int64_t i = 0; i < m; ++i)
        for (int64_t j = indptr[i]; j < indptr[i + 1]; ++j) {
            int64_t p = cur[indices[j]]++;
            csc_indices[p] = (int32_t)i;
            perm[p] = j;
        }
    free(cur);
}

/* ------------------------------------------------------------------------------------------
 * typed operators, instantiated for f32 and f64
 * ---------------------------------------------------------------------------------------- */
/* ---- JITC -> CSR materialisation  (_jit_scalar/csr.py:112-205 count, :310-411 fill; the normal
 * and uniform modules are the same walk with per-entry weights) --------------------------------
 * corder=1: generator row == CSR row, walk over the n_cols columns.
 * corder=0: generator rows are the n_cols COLUMNS, the walk runs over the n_rows CSR rows.
 * chunk_size always comes from n_cols (the reference's quirk, csr.py:135).  counts is
 * int64[n_rows]. */
void ora_jit_csr_count(uint32_t clen, uint32_t seed0, int64_t n_rows, int64_t n_cols,
                       int64_t chunk_size, int stride, int corder, int64_t* counts) {
    uint32_t cl = clen < 2U ? 2U : clen;
    for (int64_t r = 0; r < n_rows; ++r) counts[r] = 0;
    if (corder) {
        for (int64_t row = 0; row < n_rows; ++row)
            ORA_WALK_ROW(seed0, cl, row, n_cols, chunk_size, stride, { (void)j; counts[row] += 1; });
    } else {
        for (int64_t row = 0; row < n_cols; ++row)
            ORA_WALK_ROW(seed0, cl, row, n_rows, chunk_size, stride, { counts[j] += 1; });
    }
}

#define T float
#define SFX(name) name##_f32
#include "be_oracle_ops.inc"
#undef T
#undef SFX
#define T double
#define SFX(name) name##_f64
#include "be_oracle_ops.inc"
#undef T
#undef SFX
