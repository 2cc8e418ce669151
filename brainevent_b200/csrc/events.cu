// events.cu -- spike encoders: bit-pack and ORDERED stream compaction in one pass.
//
// Replaces ref: _event/compact.cu:70-430 (compact_1d, compact_2d_only, fused_bitpack_compact_2d)
// and _jit_scalar/binary_jitsmv.cu:107-125 (pack_bool).  The reference's CUDA compaction reserves
// slots with a per-warp atomicAdd and is therefore not order-stable; here a
// single-pass decoupled look-back scan yields ascending `active_ids` (the numba / jax_raw order,
// ref: _event/compact.py:305-358) while reading the spikes exactly once with 128-bit loads.
#include "common.cuh"

#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cstring>

namespace be {

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return BE_B200_EINTERNAL;
}

NvtxRange::NvtxRange(const char* name) { nvtxRangePushA(name); }
NvtxRange::~NvtxRange() { nvtxRangePop(); }

int sm_count() {
    static int cached[64] = {0};      // benign race: every writer stores the same value
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int n = cached[dev];
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return n;
}

// ------------------------------------------------------------------------------------------
// kernel-duration profiling (diagnostic; process-wide)
// ------------------------------------------------------------------------------------------
static struct {
    volatile int on;
    int cap;
    int count;
    cudaEvent_t* start;
    cudaEvent_t* stop;
} g_prof = {0, 0, 0, nullptr, nullptr};

ProfileScope::ProfileScope(cudaStream_t stream) : s(stream), slot(-1) {
    if (!g_prof.on) return;
    int i = __sync_fetch_and_add(&g_prof.count, 1);
    if (i >= g_prof.cap) return;
    slot = i;
    cudaEventRecord(g_prof.start[i], s);
}
ProfileScope::~ProfileScope() {
    if (slot >= 0) cudaEventRecord(g_prof.stop[slot], s);
}

// ------------------------------------------------------------------------------------------
// tile status word: [63:62] flag, [31:0] value
// ------------------------------------------------------------------------------------------
constexpr unsigned long long kFlagAgg = 1ull << 62;
constexpr unsigned long long kFlagIncl = 2ull << 62;
constexpr unsigned long long kFlagMask = 3ull << 62;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Block-wide exclusive scan of one count per thread (256 threads) + decoupled look-back across
// tiles.  Returns this thread's global exclusive offset; `tile` is the dynamic tile id.
// After the call, *tile_total_incl holds the inclusive prefix up to and including this tile.
__device__ __forceinline__ unsigned int scan_and_lookback(unsigned int count, unsigned int tile,
                                                          unsigned long long* status,
                                                          unsigned int* tile_total_incl,
                                                          unsigned int* cursor = nullptr) {
    __shared__ unsigned int s_warp[kCompactThreads / 32];
    __shared__ unsigned int s_excl;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int incl = count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned int warp_base = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < kCompactThreads / 32; ++w) {
        unsigned int t = s_warp[w];
        if (w < warp) warp_base += t;
        block_total += t;
    }
    if (cursor != nullptr) {
        // unordered variant: reserve this tile's slots with one atomic, no look-back chain
        if (threadIdx.x == 0) s_excl = atomicAdd(cursor, block_total);
        __syncthreads();
        *tile_total_incl = s_excl + block_total;
        return s_excl + warp_base + (incl - count);
    }
    if (warp == 0) {
        unsigned int excl = 0;
        if (tile == 0) {
            if (lane == 0) st_status(status, kFlagIncl | block_total);
        } else {
            if (lane == 0) st_status(status + tile, kFlagAgg | block_total);
            int look = (int)tile - 1;
            while (true) {
                int idx = look - lane;
                unsigned long long v = kFlagIncl;  // virtual predecessor of tile 0: inclusive 0
                if (idx >= 0) {
                    do { v = ld_status(status + idx); } while ((v & kFlagMask) == 0);
                }
                unsigned int incl_mask = __ballot_sync(0xffffffffu, (v & kFlagMask) == kFlagIncl);
                unsigned int val = (unsigned int)(v & 0xffffffffull);
                if (incl_mask) {
                    int first = __ffs(incl_mask) - 1;  // nearest predecessor with an inclusive prefix
                    if (lane > first) val = 0;
                    excl += warp_sum(val);
                    break;
                }
                excl += warp_sum(val);
                look -= 32;
            }
            if (lane == 0) st_status(status + tile, kFlagIncl | (unsigned long long)(excl + block_total));
        }
        if (lane == 0) s_excl = excl;
    }
    __syncthreads();
    *tile_total_incl = s_excl + block_total;
    return s_excl + warp_base + (incl - count);
}

// Called by every block after its look-back: the last block to arrive restores the workspace
// (ticket, done, tile status) to zero and publishes the event total, so the next call needs no
// memset.
__device__ __forceinline__ void finish_tile(WsHeader* hdr, unsigned long long* status,
                                            unsigned int n_tiles, unsigned int block_events,
                                            bool unordered = false, int* n_active_out = nullptr) {
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        if (block_events) atomicAdd(&hdr->ev_accum, block_events);
        __threadfence();
        unsigned int prev = atomicAdd(&hdr->done, 1u);
        s_last = (prev == n_tiles - 1);
    }
    __syncthreads();
    if (s_last) {
        if (!unordered)
            for (unsigned int i = threadIdx.x; i < n_tiles; i += blockDim.x) status[i] = 0ull;
        if (threadIdx.x == 0) {
            __threadfence();
            if (unordered) {
                const int total = (int)atomicExch(&hdr->cursor, 0u);   // every tile has reserved by now
                hdr->n_active = total;
                if (n_active_out != nullptr) *n_active_out = total;
            }
            hdr->n_events = atomicExch(&hdr->ev_accum, 0u);
            hdr->ticket = 0u;
            hdr->done = 0u;
        }
    }
}

// ------------------------------------------------------------------------------------------
// 16 consecutive elements -> 16-bit activity mask
// ------------------------------------------------------------------------------------------
template <int KIND, bool NONZERO>
__device__ __forceinline__ bool elem_active(const void* spikes, int64_t i) {
    if (KIND == BE_B200_SPK_BOOL) return ((const uint8_t*)spikes)[i] != 0;
    float f = ((const float*)spikes)[i];
    return NONZERO ? (f != 0.0f) : (f > 0.0f);
}

template <int KIND, bool NONZERO, bool ALIGNED>
__device__ __forceinline__ unsigned int load_mask16(const void* __restrict__ spikes, int64_t base,
                                                    int64_t n) {
    unsigned int m = 0;
    if (base >= n) return 0;
    if (KIND == BE_B200_SPK_PACKED) {
        const uint32_t* w = (const uint32_t*)spikes;
        m = (__ldg(w + (base >> 5)) >> (base & 31)) & 0xffffu;
        int64_t rem = n - base;
        if (rem < 16) m &= (1u << rem) - 1u;
        return m;
    }
    if (ALIGNED && base + 16 <= n) {
        if (KIND == BE_B200_SPK_BOOL) {
            int4 v = ldg_stream((const int4*)((const uint8_t*)spikes + base));
            unsigned int w[4] = {(unsigned)v.x, (unsigned)v.y, (unsigned)v.z, (unsigned)v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if ((w[q] >> (8 * b)) & 0xffu) m |= 1u << (q * 4 + b);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int4 v = ldg_stream((const int4*)((const float*)spikes + base) + q);
                float f[4] = {__int_as_float(v.x), __int_as_float(v.y), __int_as_float(v.z),
                              __int_as_float(v.w)};
#pragma unroll
                for (int b = 0; b < 4; ++b)
                    if (NONZERO ? (f[b] != 0.0f) : (f[b] > 0.0f)) m |= 1u << (q * 4 + b);
            }
        }
        return m;
    }
    for (int b = 0; b < 16; ++b)
        if (base + b < n && elem_active<KIND, NONZERO>(spikes, base + b)) m |= 1u << b;
    return m;
}

// ------------------------------------------------------------------------------------------
// 1-D: pack only (no scan) -- used when a consumer needs the bit vector but not the id list
// ------------------------------------------------------------------------------------------
template <int KIND, bool NONZERO, bool ALIGNED>
__global__ void __launch_bounds__(256) bitpack_1d_kernel(const void* __restrict__ spikes, int64_t n,
                                                         uint32_t* __restrict__ packed) {
    const int64_t n_words = ceil_div(n, 32);
    const int lane = threadIdx.x & 31;
    // h indexes 16-element halves; the loop bound is warp-uniform so the shuffle is convergent
    for (int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; (h - lane) < 2 * n_words;
         h += (int64_t)gridDim.x * blockDim.x) {
        unsigned int m = load_mask16<KIND, NONZERO, ALIGNED>(spikes, h * 16, n);
        unsigned int hi = __shfl_down_sync(0xffffffffu, m, 1);
        if ((h & 1) == 0 && (h >> 1) < n_words) packed[h >> 1] = m | (hi << 16);
    }
}

// ------------------------------------------------------------------------------------------
// 1-D: ordered compaction (+ optional packed output)
// ------------------------------------------------------------------------------------------
template <int KIND, bool NONZERO, bool ALIGNED, bool ORDERED>
__global__ void __launch_bounds__(kCompactThreads) compact_1d_kernel(
    const void* __restrict__ spikes, int64_t n, unsigned int n_tiles, WsHeader* hdr,
    unsigned long long* status, int32_t* __restrict__ active_ids, int* __restrict__ n_active_out,
    uint32_t* __restrict__ packed, int4* __restrict__ zero_ptr, size_t zero_vec16) {
    __shared__ unsigned int s_tile;
    if (ORDERED) {
        if (threadIdx.x == 0) s_tile = atomicAdd(&hdr->ticket, 1u);
        __syncthreads();
    }
    const unsigned int tile = ORDERED ? s_tile : blockIdx.x;
    if (zero_ptr != nullptr) {
        // fused clear of the consumer's output: this block's slice, 16-byte stores
        const size_t per = (zero_vec16 + n_tiles - 1) / n_tiles;
        const size_t z0 = (size_t)blockIdx.x * per, z1 = min(z0 + per, zero_vec16);
        for (size_t z = z0 + threadIdx.x; z < z1; z += kCompactThreads) zero_ptr[z] = make_int4(0, 0, 0, 0);
    }
    const int64_t base = (int64_t)tile * kCompactTile + (int64_t)threadIdx.x * kCompactElemsPerThread;
    unsigned int m = load_mask16<KIND, NONZERO, ALIGNED>(spikes, base, n);
    if (packed != nullptr) {
        unsigned int hi = __shfl_down_sync(0xffffffffu, m, 1);
        if ((threadIdx.x & 1) == 0 && base < n) packed[base >> 5] = m | (hi << 16);
    }
    unsigned int total_incl;
    unsigned int pos = scan_and_lookback(__popc(m), tile, status, &total_incl, ORDERED ? nullptr : &hdr->cursor);
    while (m) {
        int b = __ffs(m) - 1;
        m &= m - 1;
        active_ids[pos++] = (int32_t)(base + b);
    }
    if (ORDERED && tile == n_tiles - 1 && threadIdx.x == 0) {
        hdr->n_active = (int)total_incl;
        if (n_active_out != nullptr) *n_active_out = (int)total_incl;
    }
    finish_tile(hdr, status, n_tiles, 0u, !ORDERED, n_active_out);
}

// ------------------------------------------------------------------------------------------
// 2-D [n, batch] row-major: pack along batch, rows with any active column, ascending
// ------------------------------------------------------------------------------------------
template <int KIND, bool NONZERO>
__device__ __forceinline__ uint32_t pack_row_word(const void* __restrict__ spikes, int64_t row,
                                                  int64_t batch, int w, bool vec_ok) {
    int64_t c0 = (int64_t)w * 32;
    int64_t off = row * batch + c0;
    uint32_t bits = 0;
    if (KIND == BE_B200_SPK_BOOL && vec_ok && c0 + 32 <= batch) {
        const int4* p = (const int4*)((const uint8_t*)spikes + off);
        int4 a = ldg_stream(p), b = ldg_stream(p + 1);
        unsigned int v[8] = {(unsigned)a.x, (unsigned)a.y, (unsigned)a.z, (unsigned)a.w,
                             (unsigned)b.x, (unsigned)b.y, (unsigned)b.z, (unsigned)b.w};
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if ((v[q] >> (8 * k)) & 0xffu) bits |= 1u << (q * 4 + k);
        return bits;
    }
    int lim = (int)((batch - c0) < 32 ? (batch - c0) : 32);
    for (int b = 0; b < lim; ++b)
        if (elem_active<KIND, NONZERO>(spikes, off + b)) bits |= 1u << b;
    return bits;
}

template <int KIND, bool NONZERO>
__global__ void __launch_bounds__(kCompactThreads) compact_2d_kernel(
    const void* __restrict__ spikes, int64_t n, int64_t batch, int wpr, bool vec_ok,
    unsigned int n_tiles, WsHeader* hdr, unsigned long long* status,
    int32_t* __restrict__ active_ids, int* __restrict__ n_active_out,
    uint32_t* __restrict__ packed, bool input_is_packed) {
    __shared__ unsigned int s_tile;
    __shared__ unsigned int s_row_any[kCompact2dRowsPerTile / 32];  // bit per row of the tile
    __shared__ unsigned int s_events;
    if (threadIdx.x == 0) { s_tile = atomicAdd(&hdr->ticket, 1u); s_events = 0; }
    if (threadIdx.x < kCompact2dRowsPerTile / 32) s_row_any[threadIdx.x] = 0u;
    __syncthreads();
    const unsigned int tile = s_tile;
    const int64_t row0 = (int64_t)tile * kCompact2dRowsPerTile;
    const int rows_here = (int)((n - row0) < kCompact2dRowsPerTile ? (n - row0) : kCompact2dRowsPerTile);
    // stage 1: one thread per packed word of the tile (coalesced word writes)
    unsigned int ev = 0;
    const int n_words = rows_here * wpr;
    for (int t = threadIdx.x; t < n_words; t += blockDim.x) {
        int r = t / wpr, w = t - r * wpr;
        uint32_t bits;
        if (input_is_packed) bits = __ldg((const uint32_t*)spikes + (row0 + r) * wpr + w);
        else {
            bits = pack_row_word<KIND, NONZERO>(spikes, row0 + r, batch, w, vec_ok);
            if (packed != nullptr) packed[(row0 + r) * wpr + w] = bits;
        }
        if (bits) { atomicOr(&s_row_any[r >> 5], 1u << (r & 31)); ev += __popc(bits); }
    }
    ev = warp_sum(ev);
    if ((threadIdx.x & 31) == 0 && ev) atomicAdd(&s_events, ev);
    __syncthreads();
    // stage 2: 4 rows per thread -> scan -> ordered ids
    unsigned int m4 = (s_row_any[threadIdx.x >> 3] >> ((threadIdx.x & 7) * 4)) & 0xfu;
    unsigned int total_incl;
    unsigned int pos = scan_and_lookback(__popc(m4), tile, status, &total_incl);
    while (m4) {
        int b = __ffs(m4) - 1;
        m4 &= m4 - 1;
        active_ids[pos++] = (int32_t)(row0 + threadIdx.x * 4 + b);
    }
    if (tile == n_tiles - 1 && threadIdx.x == 0) {
        hdr->n_active = (int)total_incl;
        if (n_active_out != nullptr) *n_active_out = (int)total_incl;
    }
    finish_tile(hdr, status, n_tiles, s_events);
}

// ------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------
static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

template <int KIND>
static int launch_bitpack(const void* spikes, int64_t n, bool nonzero, uint32_t* packed,
                          cudaStream_t s) {
    int64_t n_half = 2 * ceil_div(n, 32);
    int blocks = (int)std::min<int64_t>(ceil_div(n_half, 256), kNumSM * 8);
    if (blocks < 1) blocks = 1;
    bool al = aligned16(spikes);
#define BE_BP(NZ, AL) bitpack_1d_kernel<KIND, NZ, AL><<<blocks, 256, 0, s>>>(spikes, n, packed)
    if (nonzero) { if (al) BE_BP(true, true); else BE_BP(true, false); }
    else { if (al) BE_BP(false, true); else BE_BP(false, false); }
#undef BE_BP
    BE_LAUNCH_CHECK("bitpack_1d_kernel");
    return BE_B200_OK;
}

template <int KIND>
static int launch_compact_1d(const void* spikes, int64_t n, bool nonzero, const Workspace& ws,
                             int32_t* ids, int* n_active_out, uint32_t* packed, cudaStream_t s,
                             void* zero_ptr = nullptr, size_t zero_bytes = 0, bool ordered = true) {
    unsigned int n_tiles = (unsigned int)ceil_div(n, kCompactTile);
    bool al = aligned16(spikes);
    int4* zp = (int4*)zero_ptr;
    const size_t zv = zero_bytes / 16;
#define BE_C1(NZ, AL)                                                                        \
    do {                                                                                     \
        if (ordered) compact_1d_kernel<KIND, NZ, AL, true><<<n_tiles, kCompactThreads, 0, s>>>(  \
            spikes, n, n_tiles, ws.hdr, ws.status, ids, n_active_out, packed, zp, zv);       \
        else compact_1d_kernel<KIND, NZ, AL, false><<<n_tiles, kCompactThreads, 0, s>>>(     \
            spikes, n, n_tiles, ws.hdr, ws.status, ids, n_active_out, packed, zp, zv);       \
    } while (0)
    if (nonzero) { if (al) BE_C1(true, true); else BE_C1(true, false); }
    else { if (al) BE_C1(false, true); else BE_C1(false, false); }
#undef BE_C1
    BE_LAUNCH_CHECK("compact_1d_kernel");
    return BE_B200_OK;
}

static int check_ws(void* workspace, size_t workspace_bytes, int64_t n, int64_t batch) {
    BE_REQUIRE(workspace != nullptr, "workspace is NULL");
    BE_REQUIRE(workspace_bytes >= be_b200_workspace_bytes(n, batch),
               "workspace too small: %zu < %zu", workspace_bytes, be_b200_workspace_bytes(n, batch));
    BE_REQUIRE(((uintptr_t)workspace & 255u) == 0, "workspace must be 256-byte aligned");
    return BE_B200_OK;
}

static int zero_n_active(int* dst, cudaStream_t s) {
    BE_CUDA(cudaMemsetAsync(dst, 0, sizeof(int), s));
    return BE_B200_OK;
}

int compact_1d_any(const void* spikes, int spk_kind, int64_t n, bool nonzero, const Workspace& ws,
                   int32_t* ids, int* n_active_out, uint32_t* packed, cudaStream_t s,
                   void* zero_ptr = nullptr, size_t zero_bytes = 0, bool ordered = true) {
    // the fused clear needs a 16-byte aligned region; the (<16 B) tail is cleared separately
    if (zero_ptr != nullptr && ((uintptr_t)zero_ptr & 15u) != 0) {
        BE_CUDA(cudaMemsetAsync(zero_ptr, 0, zero_bytes, s));
        zero_ptr = nullptr;
    } else if (zero_ptr != nullptr && (zero_bytes & 15u) != 0) {
        BE_CUDA(cudaMemsetAsync((char*)zero_ptr + (zero_bytes & ~(size_t)15), 0, zero_bytes & 15u, s));
    }
    if (n == 0) {
        if (zero_ptr != nullptr && zero_bytes >= 16) BE_CUDA(cudaMemsetAsync(zero_ptr, 0, zero_bytes, s));
        int rc = zero_n_active(&ws.hdr->n_active, s);
        if (rc) return rc;
        if (n_active_out) return zero_n_active(n_active_out, s);
        return BE_B200_OK;
    }
    switch (spk_kind) {
        case BE_B200_SPK_BOOL: return launch_compact_1d<BE_B200_SPK_BOOL>(spikes, n, nonzero, ws, ids, n_active_out, packed, s, zero_ptr, zero_bytes, ordered);
        case BE_B200_SPK_F32: return launch_compact_1d<BE_B200_SPK_F32>(spikes, n, nonzero, ws, ids, n_active_out, packed, s, zero_ptr, zero_bytes, ordered);
        case BE_B200_SPK_PACKED: return launch_compact_1d<BE_B200_SPK_PACKED>(spikes, n, nonzero, ws, ids, n_active_out, nullptr, s, zero_ptr, zero_bytes, ordered);
        default: set_error("compact_1d: unsupported event kind %d", spk_kind); return BE_B200_EINVAL;
    }
}

int compact_2d_any(const void* spikes, int spk_kind, int64_t n, int64_t batch, bool nonzero,
                   const Workspace& ws, uint32_t* packed, int32_t* ids, int* n_active_out,
                   cudaStream_t s) {
    if (n == 0) {
        int rc = zero_n_active(&ws.hdr->n_active, s);
        if (rc) return rc;
        if (n_active_out) return zero_n_active(n_active_out, s);
        return BE_B200_OK;
    }
    unsigned int n_tiles = (unsigned int)ceil_div(n, kCompact2dRowsPerTile);
    int wpr = (int)ceil_div(batch, 32);
    bool vec_ok = aligned16(spikes) && (batch % 16 == 0);
#define BE_C2(KIND, NZ, PK)                                                                   \
    compact_2d_kernel<KIND, NZ><<<n_tiles, kCompactThreads, 0, s>>>(                          \
        spikes, n, batch, wpr, vec_ok, n_tiles, ws.hdr, ws.status, ids, n_active_out, packed, PK)
    switch (spk_kind) {
        case BE_B200_SPK_BOOL: if (nonzero) BE_C2(BE_B200_SPK_BOOL, true, false); else BE_C2(BE_B200_SPK_BOOL, false, false); break;
        case BE_B200_SPK_F32: if (nonzero) BE_C2(BE_B200_SPK_F32, true, false); else BE_C2(BE_B200_SPK_F32, false, false); break;
        case BE_B200_SPK_PACKED: BE_C2(BE_B200_SPK_BOOL, false, true); break;
        default: set_error("compact_2d: unsupported event kind %d", spk_kind); return BE_B200_EINVAL;
    }
#undef BE_C2
    BE_LAUNCH_CHECK("compact_2d_kernel");
    return BE_B200_OK;
}

int prepare_events_1d(const void* spikes, const void* aux, int spk_kind, int64_t n, bool need_ids,
                      bool need_packed, void* workspace, size_t workspace_bytes, cudaStream_t s,
                      Events* ev, void* zero_ptr, size_t zero_bytes, bool ordered_ids, bool nonzero) {
    int rc = check_ws(workspace, workspace_bytes, n, 1);
    if (rc) return rc;
    Workspace ws = ws_carve(workspace, n, 1);
    ev->words_per_row = 1;
    const bool is_compact = spk_kind == BE_B200_SPK_COMPACT || spk_kind == BE_B200_SPK_COMPACT_CLEARED;
    const bool compacts = need_ids && !is_compact;
    if (!compacts && zero_ptr != nullptr && zero_bytes > 0 && spk_kind != BE_B200_SPK_COMPACT_CLEARED)
        BE_CUDA(cudaMemsetAsync(zero_ptr, 0, zero_bytes, s));
    if (is_compact) {
        BE_REQUIRE(aux != nullptr, "COMPACT events need aux = active_ids[n] ++ n_active");
        BE_REQUIRE(!need_packed || spikes != nullptr, "COMPACT events: packed words required here");
        ev->packed = (const uint32_t*)spikes;
        ev->active_ids = (const int32_t*)aux;
        ev->n_active = (const int*)aux + n;
        return BE_B200_OK;
    }
    BE_REQUIRE(spikes != nullptr || n == 0, "spikes is NULL");
    ev->active_ids = ws.active_ids;
    ev->n_active = &ws.hdr->n_active;
    if (spk_kind == BE_B200_SPK_PACKED) {
        ev->packed = (const uint32_t*)spikes;
        if (need_ids) return compact_1d_any(spikes, spk_kind, n, false, ws, ws.active_ids, nullptr, nullptr, s, zero_ptr, zero_bytes, ordered_ids);
        return BE_B200_OK;
    }
    ev->packed = ws.packed;
    if (need_ids)
        return compact_1d_any(spikes, spk_kind, n, nonzero, ws, ws.active_ids, nullptr,
                              need_packed ? ws.packed : nullptr, s, zero_ptr, zero_bytes, ordered_ids);
    if (need_packed && n > 0) {
        if (spk_kind == BE_B200_SPK_BOOL && ev->defer_pack_ok && n <= kDeferPackMax) {
            ev->raw_bool = (const uint8_t*)spikes;      // nonzero byte == active either way for bool
            ev->raw_n = n;
            return BE_B200_OK;
        }
        if (spk_kind == BE_B200_SPK_BOOL) return launch_bitpack<BE_B200_SPK_BOOL>(spikes, n, nonzero, ws.packed, s);
        if (spk_kind == BE_B200_SPK_F32) return launch_bitpack<BE_B200_SPK_F32>(spikes, n, nonzero, ws.packed, s);
        set_error("unsupported event kind %d", spk_kind);
        return BE_B200_EINVAL;
    }
    return BE_B200_OK;
}

int pack_deferred(Events* ev, cudaStream_t s) {
    if (ev->raw_bool == nullptr) return BE_B200_OK;
    const void* raw = ev->raw_bool;
    ev->raw_bool = nullptr;
    return launch_bitpack<BE_B200_SPK_BOOL>(raw, ev->raw_n, false, const_cast<uint32_t*>(ev->packed), s);
}

int prepare_events_2d(const void* spikes, const void* aux, int spk_kind, int64_t n, int64_t batch,
                      bool need_ids, void* workspace, size_t workspace_bytes, cudaStream_t s,
                      Events* ev) {
    int rc = check_ws(workspace, workspace_bytes, n, batch);
    if (rc) return rc;
    Workspace ws = ws_carve(workspace, n, batch);
    ev->words_per_row = (int)ceil_div(batch, 32);
    if (spk_kind == BE_B200_SPK_COMPACT) {
        BE_REQUIRE(aux != nullptr && spikes != nullptr, "COMPACT events need packed words and aux");
        ev->packed = (const uint32_t*)spikes;
        ev->active_ids = (const int32_t*)aux;
        ev->n_active = (const int*)aux + n;
        return BE_B200_OK;
    }
    BE_REQUIRE(spikes != nullptr || n == 0, "spikes is NULL");
    ev->active_ids = ws.active_ids;
    ev->n_active = &ws.hdr->n_active;
    if (spk_kind == BE_B200_SPK_PACKED) {
        ev->packed = (const uint32_t*)spikes;
        if (!need_ids) return BE_B200_OK;
        return compact_2d_any(spikes, spk_kind, n, batch, false, ws, nullptr, ws.active_ids, nullptr, s);
    }
    ev->packed = ws.packed;
    // packing a 2-D spike matrix always goes through the fused pack+compact kernel
    return compact_2d_any(spikes, spk_kind, n, batch, false, ws, ws.packed, ws.active_ids, nullptr, s);
}

}  // namespace be

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
using namespace be;

extern "C" const char* be_b200_last_error(void) { return be::g_err; }
extern "C" int be_b200_abi_version(void) { return BE_B200_ABI_VERSION; }

extern "C" int be_b200_check_device(void) {
    int dev = 0;
    BE_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    BE_CUDA(cudaGetDeviceProperties(&p, dev));
    BE_REQUIRE(p.major == 10, "brainevent_b200 is built for sm_100a only; device %d is sm_%d%d",
               dev, p.major, p.minor);
    return BE_B200_OK;
}

extern "C" int be_b200_profile_begin(int max_records) {
    BE_REQUIRE(max_records > 0 && max_records <= (1 << 20), "profile_begin: bad capacity");
    BE_REQUIRE(!be::g_prof.on, "profile_begin: already profiling");
    be::g_prof.start = new cudaEvent_t[max_records];
    be::g_prof.stop = new cudaEvent_t[max_records];
    for (int i = 0; i < max_records; ++i) {
        BE_CUDA(cudaEventCreate(&be::g_prof.start[i]));
        BE_CUDA(cudaEventCreate(&be::g_prof.stop[i]));
    }
    be::g_prof.cap = max_records;
    be::g_prof.count = 0;
    be::g_prof.on = 1;
    return BE_B200_OK;
}

extern "C" int be_b200_profile_end(float* ms_out, int capacity, int* n_out) {
    BE_REQUIRE(be::g_prof.on, "profile_end: not profiling");
    be::g_prof.on = 0;
    int n = be::g_prof.count < be::g_prof.cap ? be::g_prof.count : be::g_prof.cap;
    int rc = BE_B200_OK;
    for (int i = 0; i < n; ++i) {
        float ms = 0.f;
        cudaError_t e = cudaEventSynchronize(be::g_prof.stop[i]);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, be::g_prof.start[i], be::g_prof.stop[i]);
        if (e != cudaSuccess) rc = be::cuda_fail(e, "profile_end");
        if (ms_out && i < capacity) ms_out[i] = ms;
    }
    for (int i = 0; i < be::g_prof.cap; ++i) { cudaEventDestroy(be::g_prof.start[i]); cudaEventDestroy(be::g_prof.stop[i]); }
    delete[] be::g_prof.start; delete[] be::g_prof.stop;
    be::g_prof.start = be::g_prof.stop = nullptr;
    if (n_out) *n_out = n;
    return rc;
}

extern "C" size_t be_b200_workspace_bytes(int64_t n_src, int64_t n_batch) {
    if (n_src < 0) n_src = 0;
    if (n_batch < 1) n_batch = 1;
    return ws_carve(nullptr, n_src, n_batch).bytes;
}

extern "C" int be_b200_bitpack_1d(const void* spikes, int spk_kind, int64_t n, int nonzero_pred,
                                  uint32_t* packed, void* stream) { BE_NVTX_RANGE("be_b200_bitpack_1d");
    cudaStream_t s = (cudaStream_t)stream;
    BE_REQUIRE(n >= 0, "n must be >= 0");
    if (n == 0) return BE_B200_OK;
    BE_REQUIRE(spikes && packed, "NULL pointer");
    if (spk_kind == BE_B200_SPK_BOOL) return launch_bitpack<BE_B200_SPK_BOOL>(spikes, n, nonzero_pred != 0, packed, s);
    if (spk_kind == BE_B200_SPK_F32) return launch_bitpack<BE_B200_SPK_F32>(spikes, n, nonzero_pred != 0, packed, s);
    set_error("bitpack_1d: unsupported event kind %d", spk_kind);
    return BE_B200_EINVAL;
}

extern "C" int be_b200_compact_1d(const void* spikes, int spk_kind, int64_t n, int nonzero_pred,
                                  int32_t* active_ids, int32_t* n_active, uint32_t* packed,
                                  void* workspace, size_t workspace_bytes, void* stream) { BE_NVTX_RANGE("be_b200_compact_1d");
    cudaStream_t s = (cudaStream_t)stream;
    BE_REQUIRE(n >= 0 && n < (1ll << 31), "n out of range");
    BE_REQUIRE((active_ids || n == 0) && n_active, "NULL output");
    int rc = check_ws(workspace, workspace_bytes, n, 1);
    if (rc) return rc;
    Workspace ws = ws_carve(workspace, n, 1);
    if (n > 0) BE_CUDA(cudaMemsetAsync(active_ids, 0, (size_t)n * 4, s));  // zero tail
    return compact_1d_any(spikes, spk_kind, n, nonzero_pred != 0, ws, active_ids, n_active, packed, s);
}

extern "C" int be_b200_compact_2d(const void* spikes, int spk_kind, int64_t n, int64_t batch,
                                  int nonzero_pred, uint32_t* packed, int32_t* active_ids,
                                  int32_t* n_active, void* workspace, size_t workspace_bytes,
                                  void* stream) { BE_NVTX_RANGE("be_b200_compact_2d");
    cudaStream_t s = (cudaStream_t)stream;
    BE_REQUIRE(n >= 0 && n < (1ll << 31) && batch >= 1, "shape out of range");
    BE_REQUIRE((active_ids || n == 0) && n_active, "NULL output");
    int rc = check_ws(workspace, workspace_bytes, n, batch);
    if (rc) return rc;
    Workspace ws = ws_carve(workspace, n, batch);
    if (n > 0) BE_CUDA(cudaMemsetAsync(active_ids, 0, (size_t)n * 4, s));
    return compact_2d_any(spikes, spk_kind, n, batch, nonzero_pred != 0, ws, packed, active_ids, n_active, s);
}
