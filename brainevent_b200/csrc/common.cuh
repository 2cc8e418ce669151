// common.cuh -- shared device/host helpers for the sm_100a event-propagation kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/brainevent_b200.h"

namespace be {

// ------------------------------------------------------------------------------------------
// errors: thread-local message, integer return codes (no exceptions across the C ABI)
// ------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define BE_REQUIRE(cond, ...)                         \
    do {                                              \
        if (!(cond)) {                                \
            ::be::set_error(__VA_ARGS__);             \
            return BE_B200_EINVAL;                    \
        }                                             \
    } while (0)

#define BE_CUDA(call)                                                   \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return ::be::cuda_fail(e__, #call);     \
    } while (0)

#define BE_LAUNCH_CHECK(name)                                           \
    do {                                                                \
        cudaError_t e__ = cudaGetLastError();                           \
        if (e__ != cudaSuccess) return ::be::cuda_fail(e__, name);      \
    } while (0)

// SM count of the CURRENT device, queried once per device (148 on B200: 2 dies x 74 SMs);
// every persistent grid is sized from it.
int sm_count();
#define kNumSM (::be::sm_count())
constexpr int kWarp = 32;

// NVTX range around every C-ABI / XLA-FFI entry: visible in Nsight Systems / Compute timelines,
// a null-pointer check when no tool is attached (header-only NVTX3, resolved lazily through dlopen).
struct NvtxRange {
    explicit NvtxRange(const char* name);
    ~NvtxRange();
};
#define BE_NVTX_CAT2(a, b) a##b
#define BE_NVTX_CAT(a, b) BE_NVTX_CAT2(a, b)
#define BE_NVTX_RANGE(name) ::be::NvtxRange BE_NVTX_CAT(nvtx_range_, __LINE__)(name)

__host__ __device__ static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------
// workspace layout (see be_b200_workspace_bytes)
//   header (256 B) | active_ids int32[n_src + 1] | packed uint32[...] | tile status u64[...]
// The header and the tile-status region must be zero before first use; the compaction kernel
// leaves them zero again when it finishes (self-resetting), so steady-state calls need no memset.
// ------------------------------------------------------------------------------------------
struct WsHeader {
    int n_active;          // number of active source rows of the last compaction
    unsigned int ticket;   // dynamic tile counter (decoupled look-back)
    unsigned int done;     // tiles finished (self-reset trigger)
    unsigned int n_events; // total set bits (2-D: sum of popcounts), for rate-based selection
    unsigned int ev_accum; // running event total of the compaction in flight
    unsigned int cursor;   // slot cursor of the unordered (block-atomic) compaction
    int pad[58];
};
static_assert(sizeof(WsHeader) == 256, "header is 256 B");

constexpr int kCompactThreads = 256;
constexpr int kCompactElemsPerThread = 16;
constexpr int kCompactTile = kCompactThreads * kCompactElemsPerThread;  // 4096 elements / tile
constexpr int kCompact2dRowsPerTile = 1024;

struct Workspace {
    WsHeader* hdr;
    int32_t* active_ids;
    uint32_t* packed;
    unsigned long long* status;
    size_t bytes;
};

static inline size_t ws_packed_words(int64_t n_src, int64_t n_batch) {
    int64_t w1 = ceil_div(n_src, 32) + 4;
    int64_t w2 = n_src * ceil_div(n_batch < 1 ? 1 : n_batch, 32);  // 2-D events with batch 1 still use one word per row
    return (size_t)(w1 > w2 ? w1 : w2);
}
static inline size_t ws_status_words(int64_t n_src) {
    int64_t t1 = ceil_div(n_src, kCompactTile), t2 = ceil_div(n_src, kCompact2dRowsPerTile);
    return (size_t)((t1 > t2 ? t1 : t2) + 1);
}
static inline Workspace ws_carve(void* base, int64_t n_src, int64_t n_batch) {
    Workspace w;
    char* p = (char*)base;
    w.hdr = (WsHeader*)p;
    p += sizeof(WsHeader);
    w.active_ids = (int32_t*)p;
    p += align_up((size_t)(n_src + 1) * 4, 256);
    w.packed = (uint32_t*)p;
    p += align_up(ws_packed_words(n_src, n_batch) * 4, 256);
    w.status = (unsigned long long*)p;
    p += align_up(ws_status_words(n_src) * 8, 256);
    w.bytes = (size_t)(p - (char*)base);
    return w;
}

// Events after normalisation: what every propagation kernel consumes.
struct Events {
    const uint32_t* packed;     // 1-D: ceil(n/32) words.  2-D: [n, words_per_row]
    const int32_t* active_ids;  // ascending active source rows (may be null if not needed)
    const int* n_active;        // device scalar
    int words_per_row;          // 2-D only
    // rate-based selection between the atomic scatter and the deterministic gather of the same
    // product (decided on the device, no host sync): with sel_thr >= 0 a scatter kernel runs only
    // when n_active <= sel_thr and a gather kernel only when n_active > sel_thr.
    int sel_thr = -1;
    // small bool vectors consumed by the shared-memory gather: the caller sets defer_pack_ok, and
    // prepare_events_1d then skips the pack kernel and hands the raw bytes over (raw_bool); the
    // gather CTAs pack them straight into shared memory (one launch instead of two).  A consumer
    // that cannot use raw_bool calls pack_deferred() first.
    bool defer_pack_ok = false;
    const uint8_t* raw_bool = nullptr;
    int64_t raw_n = 0;
};
constexpr int64_t kDeferPackMax = 32768;    // sources

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 128-bit loads that do not pollute L1 (connectivity is read once per step)
#ifdef BE_STREAM_EVICT_FIRST
// experiment: connectivity streams marked evict-first in L2 so that they do not displace the L2-resident output
// tile the reductions land in (make EXTRA=-DBE_STREAM_EVICT_FIRST)
__device__ __forceinline__ uint64_t l2_evict_first() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ int4 ldg_stream(const int4* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(l2_evict_first()));
    return r;
}
#else
__device__ __forceinline__ int4 ldg_stream(const int4* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
#endif
__device__ __forceinline__ int2 ldg_stream(const int2* p) {
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];"
                 : "=r"(r.x), "=r"(r.y)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ bool bit_test(const uint32_t* __restrict__ packed, int64_t idx) {
    return (__ldg(packed + (idx >> 5)) >> (idx & 31)) & 1u;
}

// accumulation type: f64 for f64 weights, f32 otherwise (ref: cuda_common.h READ/WRITE macros)
template <typename T> struct Acc { using type = float; };
template <> struct Acc<double> { using type = double; };

template <typename T> __device__ __forceinline__ typename Acc<T>::type to_acc(T v) { return (typename Acc<T>::type)v; }
template <> __device__ __forceinline__ float to_acc<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_acc<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_acc(typename Acc<T>::type v) { return (T)v; }
template <> __device__ __forceinline__ __half from_acc<__half>(float v) { return __float2half(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_acc<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

// reduction without return value: compiles to RED.E.ADD (fire-and-forget L2 atomic)
__device__ __forceinline__ void red_add(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add(double* p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add(__half* p, float v) { atomicAdd(p, __float2half(v)); }
__device__ __forceinline__ void red_add(__nv_bfloat16* p, float v) { atomicAdd(p, __float2bfloat16(v)); }

// load 4 consecutive weights (address must be 4-element aligned) into accumulator-typed regs
template <typename T> struct Vec4 { typename Acc<T>::type v[4]; };
__device__ __forceinline__ Vec4<float> load4(const float* p) {
    int4 r = ldg_stream((const int4*)p);
    Vec4<float> o;
    o.v[0] = __int_as_float(r.x); o.v[1] = __int_as_float(r.y);
    o.v[2] = __int_as_float(r.z); o.v[3] = __int_as_float(r.w);
    return o;
}
__device__ __forceinline__ Vec4<double> load4(const double* p) {
    int4 a = ldg_stream((const int4*)p), b = ldg_stream((const int4*)p + 1);
    Vec4<double> o;
    o.v[0] = __hiloint2double(a.y, a.x); o.v[1] = __hiloint2double(a.w, a.z);
    o.v[2] = __hiloint2double(b.y, b.x); o.v[3] = __hiloint2double(b.w, b.z);
    return o;
}
__device__ __forceinline__ Vec4<__half> load4(const __half* p) {
    int2 r = ldg_stream((const int2*)p);
    __half2 a = *reinterpret_cast<__half2*>(&r.x), b = *reinterpret_cast<__half2*>(&r.y);
    Vec4<__half> o;
    o.v[0] = __low2float(a); o.v[1] = __high2float(a); o.v[2] = __low2float(b); o.v[3] = __high2float(b);
    return o;
}
__device__ __forceinline__ Vec4<__nv_bfloat16> load4(const __nv_bfloat16* p) {
    int2 r = ldg_stream((const int2*)p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x), b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
    Vec4<__nv_bfloat16> o;
    o.v[0] = __low2float(a); o.v[1] = __high2float(a); o.v[2] = __low2float(b); o.v[3] = __high2float(b);
    return o;
}

#endif  // __CUDACC__

// host-side dtype dispatch
#define BE_DISPATCH_WT(dtype, T, ...)                                             \
    switch (dtype) {                                                              \
        case BE_B200_F32: { using T = float; __VA_ARGS__; break; }                \
        case BE_B200_F64: { using T = double; __VA_ARGS__; break; }               \
        case BE_B200_F16: { using T = __half; __VA_ARGS__; break; }               \
        case BE_B200_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }       \
        default: ::be::set_error("unsupported weight dtype %d", (int)(dtype));    \
                 return BE_B200_EUNIMPL;                                          \
    }

// Optional timing of the dominant kernel of each operator (bench.py's roofline): when enabled via
// be_b200_profile_begin, a ProfileScope records a CUDA event pair around the kernel launch on
// the launching stream.  Disabled (the default) it costs one predictable branch.
struct ProfileScope {
    cudaStream_t s;
    int slot;
    explicit ProfileScope(cudaStream_t stream);
    ~ProfileScope();
};

// events.cu: normalise any supported event encoding into `Events` (packing / compaction as
// needed by the consumer), enqueued on `stream`.
// `zero_ptr/zero_bytes` (optional): an output region the caller needs cleared before it scatters;
// it is zeroed by the compaction kernel itself when one is launched (no separate memset node),
// otherwise by cudaMemsetAsync.  `ordered_ids=false` lets the id list be block-ordered only
// (one atomic per tile instead of the look-back chain) -- fine for scatter consumers.
// `nonzero`: float spikes are active when != 0 (plasticity, encoders) instead of > 0 (operators).
int prepare_events_1d(const void* spikes, const void* aux, int spk_kind, int64_t n,
                      bool need_ids, bool need_packed, void* workspace, size_t workspace_bytes,
                      cudaStream_t stream, Events* ev, void* zero_ptr = nullptr,
                      size_t zero_bytes = 0, bool ordered_ids = true, bool nonzero = false);
int prepare_events_2d(const void* spikes, const void* aux, int spk_kind, int64_t n, int64_t batch,
                      bool need_ids, void* workspace, size_t workspace_bytes, cudaStream_t stream,
                      Events* ev);
// runs the pack kernel that prepare_events_1d deferred (ev->raw_bool set) and clears the deferral
int pack_deferred(Events* ev, cudaStream_t stream);

}  // namespace be
