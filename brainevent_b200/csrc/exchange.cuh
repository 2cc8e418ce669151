// exchange.cuh -- device-side protocol pieces of the peer-memory spike exchange, shared by exchange.cu
// (generic pack / wait / pull kernels) and sharded.cu (exchange fused into the scatter consumer).
#pragma once
#include "common.cuh"

namespace be {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct ExchangeCtl {      // local (non-symmetric) control block, zero-initialised by the caller
    uint32_t step;        // exchanges completed by this rank
    uint32_t done;        // blocks of the pack kernel in flight that have finished their stores
    uint32_t timeout;     // STICKY: set when a wait gave up (a peer never published).  From then on every wait
                          // kernel of this exchange skips its pull (n_active stays 0, the gathered words keep
                          // their last value) and the host entry points raise: results after a timeout are not
                          // trusted because the rank may have run ahead of a peer's double buffer
    uint32_t spin_ms;     // bound of one flag wait in milliseconds (0 = kDefaultSpinMs)
    // %globaltimer stamps of the last exchange (diagnostics, tools/mgpu_exchange_bench.py):
    // [0] pack start, [1] last pack CTA after its fence, [2] flags published,
    // [3] wait start, [4] all flags seen, [5] copy done
    unsigned long long stamp[8];
    // work cursor of the fused exchange + scatter kernel (sharded.cu): next word chunk to hand out; reset by
    // the pack kernel of the same step (earlier in stream order)
    uint32_t next_chunk;
    uint32_t pad2[3];
};
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr uint32_t kDefaultSpinMs = 20000;   // 20 s: a host stall (GC, checkpoint, late warm-up) must not trip it

// Thread `threadIdx.x < world` waits until peer threadIdx.x has published step t; bounded by wall time
// (%globaltimer), never by poll count.  Returns false (and latches ctl->timeout) when a wait gave up
// or an earlier exchange already had: the caller must then skip its pull.
__device__ __forceinline__ bool wait_flags(const uint32_t* my_flags, int world, uint32_t t, ExchangeCtl* ctl) {
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = (int)*(volatile uint32_t*)&ctl->timeout;
    __syncthreads();
    if (!s_bad && (int)threadIdx.x < world) {
        const unsigned long long limit = (unsigned long long)(ctl->spin_ms ? ctl->spin_ms : kDefaultSpinMs) * 1000000ull;
        unsigned long long t0 = 0;
        unsigned int polls = 0;
        while ((int32_t)(ld_acquire_sys(my_flags + threadIdx.x) - t) < 0) {
            if ((++polls & 1023u) == 0) {
                const unsigned long long now = gtimer();
                if (t0 == 0) t0 = now;
                else if (now - t0 > limit) { ctl->timeout = 1u; s_bad = 1; break; }
            }
            __nanosleep(20);
        }
    }
    __syncthreads();
    return s_bad == 0;
}


}  // namespace be
