// exchange.cu -- fused bit-pack + all-gather of spike words over NVLink peer memory.
//
// Net-new relative to the reference (no distributed code there).  The compute
// step "pack my block of spikes" and the collective "all-gather the packed words" share two
// kernels: every rank packs its own block into its OWN symmetric buffer (local stores: the
// system-scope fence that follows has no NVLink writes to wait for), then publishes a
// monotonically increasing step number into a flag slot on every peer with relaxed stores behind
// that one fence.  The consumer kernel acquires the `world` flags and PULLS each peer's words over
// NVLink straight into the local gathered vector (peer pointers from a symmetric-memory
// rendezvous; NVSwitch gives each pair full bandwidth).  The whole exchange moves N/8 bytes and is
// latency-bound: one fence, one flag hop, one remote read.  (A first push-based version stored
// every word into all peers and released each flag separately: 8.7 us of pack + remote-write
// fences and 5.4 us of serial release stores at N=2, 32 us per exchange at N=8.)
//
// Buffers are double-buffered by the parity of the device-side step counter: a rank can run at
// most one exchange ahead of a peer (its step t+1 needs every peer's step t+1 pack, which follows
// that peer's step t consumption in stream order), so parity t never overwrites words of step t
// that a slower peer is still reading.  The wait kernel copies the current half into a fixed
// local buffer, so consumers (and CUDA graphs) see one static pointer.
#include "exchange.cuh"

#include <algorithm>

namespace be {

template <int KIND>
__global__ void __launch_bounds__(256) pack_publish_kernel(const void* __restrict__ spikes, int64_t n_local,
                                                           uint32_t* const* __restrict__ peer_bufs,
                                                           uint32_t* const* __restrict__ peer_flags, int world,
                                                           int rank, int64_t words_per_rank, ExchangeCtl* ctl,
                                                           int32_t* __restrict__ reset_counter,
                                                           int4* __restrict__ zero_ptr, size_t zero_vec16) {
    const uint32_t t = ctl->step + 1;      // read by every block before the last one bumps it
    const int lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        ctl->stamp[0] = gtimer();
        // n_active of the fused wait + compaction below / chunk cursor of the fused exchange + scatter:
        // the previous step's consumer is behind us in stream order, this step's ahead of us
        if (reset_counter != nullptr) *reset_counter = 0;
        ctl->next_chunk = 0u;
    }
    const int64_t slot0 = (int64_t)(t & 1u) * words_per_rank * world + (int64_t)rank * words_per_rank;
    for (int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; (h - lane) < 2 * words_per_rank;
         h += (int64_t)gridDim.x * blockDim.x) {
        unsigned int m = 0;
        const int64_t base = h * 16;
        if (base < n_local) {
            if (KIND == BE_B200_SPK_BOOL) {
                if (base + 16 <= n_local) {
                    const int4 v = ldg_stream((const int4*)((const uint8_t*)spikes + base));
                    const unsigned int w[4] = {(unsigned)v.x, (unsigned)v.y, (unsigned)v.z, (unsigned)v.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q)
#pragma unroll
                        for (int b = 0; b < 4; ++b)
                            if ((w[q] >> (8 * b)) & 0xffu) m |= 1u << (q * 4 + b);
                } else {
                    for (int b = 0; b < 16 && base + b < n_local; ++b)
                        if (((const uint8_t*)spikes)[base + b]) m |= 1u << b;
                }
            } else {
                for (int b = 0; b < 16 && base + b < n_local; ++b)
                    if (((const float*)spikes)[base + b] > 0.0f) m |= 1u << b;
            }
        }
        const unsigned int hi = __shfl_down_sync(0xffffffffu, m, 1);
        if ((h & 1) == 0 && (h >> 1) < words_per_rank)
            peer_bufs[rank][slot0 + (h >> 1)] = m | (hi << 16);      // my own symmetric buffer
    }
    __syncthreads();                       // this CTA's peer stores are ordered before thread 0's fence
    if (threadIdx.x == 0) {
        // gpu-scope fence only: the words live in THIS GPU's memory, the `done` counter orders every
        // CTA's words before the last CTA's single system-scope fence below (release cumulativity),
        // and peers read them through this GPU's L2.  A system fence per CTA cost ~5 us of the pack.
        __threadfence();
        const uint32_t prev = atomicAdd(&ctl->done, 1u);
        if (prev == gridDim.x - 1) {
            ctl->stamp[1] = gtimer();
            __threadfence_system();          // cumulative over every CTA's words (observed through `done`)
            for (int p = 0; p < world; ++p) st_relaxed_sys(peer_flags[p] + rank, t);
            ctl->stamp[2] = gtimer();
            ctl->done = 0u;
            ctl->step = t;
        }
    }
    // the consumer's output is cleared here, after this CTA's words are out and while the flags travel:
    // the scatter that follows in stream order then needs neither a memset node nor a clearing prologue
    if (zero_ptr != nullptr) {
        const size_t per = (zero_vec16 + gridDim.x - 1) / gridDim.x;
        const size_t z0 = (size_t)blockIdx.x * per, z1 = min(z0 + per, zero_vec16);
        for (size_t z = z0 + threadIdx.x; z < z1; z += blockDim.x) zero_ptr[z] = make_int4(0, 0, 0, 0);
    }
}

// every block waits until all peers have published this rank's current step, then pulls its share
// of every peer's words (current parity half of the peer's own buffer) into the local vector
__global__ void __launch_bounds__(256) wait_copy_kernel(const uint32_t* __restrict__ my_flags,
                                                        uint32_t* const* __restrict__ peer_bufs, int world,
                                                        int64_t words_per_rank, uint32_t* __restrict__ dst,
                                                        ExchangeCtl* ctl) {
    const uint32_t t = ctl->step;
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->stamp[3] = gtimer();
    if (!wait_flags(my_flags, world, t, ctl)) return;      // timed out (now or earlier): no stale words consumed
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->stamp[4] = gtimer();
    const int64_t n = words_per_rank * world;
    const int64_t half = (int64_t)(t & 1u) * n;
    // remote (NVLink) reads, ordered after the flag acquire.  A remote load is a ~2 us round trip, so
    // all of a thread's words are requested before the first one is stored (with one load per loop
    // trip the copy of 10 M bits took 14.5 us = 8 serial round trips; stamps [4] -> [5])
    constexpr int U = 8;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += stride * U) {
        uint32_t v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < n) v[u] = ld_relaxed_sys(peer_bufs[(int)(i / words_per_rank)] + half + i);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < n) dst[i] = v[u];
        }
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) ctl->stamp[5] = gtimer();
}


// wait + pull + COMPACTION + output clear in one kernel, for scatter consumers at N > 1.  While the
// peers are still publishing, every block clears its slice of the consumer's output vector; after the
// flags it pulls the words exactly like wait_copy_kernel and, with the words still in registers,
// turns them into the active-id list (block scan of the popcounts, one atomic per block and trip on
// the global cursor; the order of the list is irrelevant to a scatter).  This removes the separate
// compaction kernel (5-11 us on 10 M bits, mostly launch + tail latency) from the per-timestep chain.
__global__ void __launch_bounds__(256) wait_copy_compact_kernel(const uint32_t* __restrict__ my_flags,
                                                                uint32_t* const* __restrict__ peer_bufs, int world,
                                                                int64_t words_per_rank, int64_t n_bits,
                                                                uint32_t* __restrict__ dst, int32_t* __restrict__ ids,
                                                                int32_t* __restrict__ n_active,
                                                                int4* __restrict__ zero_ptr, size_t zero_vec16,
                                                                ExchangeCtl* ctl) {
    const uint32_t t = ctl->step;
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->stamp[3] = gtimer();
    if (zero_ptr != nullptr) {
        const size_t per = (zero_vec16 + gridDim.x - 1) / gridDim.x;
        const size_t z0 = (size_t)blockIdx.x * per, z1 = min(z0 + per, zero_vec16);
        for (size_t z = z0 + threadIdx.x; z < z1; z += blockDim.x) zero_ptr[z] = make_int4(0, 0, 0, 0);
    }
    if (!wait_flags(my_flags, world, t, ctl)) return;      // n_active stays 0: the consumer scatters nothing
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->stamp[4] = gtimer();
    const int64_t n = words_per_rank * world;
    const int64_t half = (int64_t)(t & 1u) * n;
    constexpr int U = 8;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ int s_warp[8];
    __shared__ int s_base;
    for (int64_t b0 = (int64_t)blockIdx.x * blockDim.x; b0 < n; b0 += stride * U) {     // block-uniform trip count
        const int64_t i0 = b0 + threadIdx.x;
        uint32_t v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            v[u] = (i < n) ? ld_relaxed_sys(peer_bufs[(int)(i / words_per_rank)] + half + i) : 0u;
        }
        int cnt = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < n) dst[i] = v[u];
            cnt += __popc(v[u]);
        }
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        int woff = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const int c = s_warp[w];
            if (w < warp) woff += c;
            total += c;
        }
        if (threadIdx.x == 0) s_base = total ? atomicAdd(n_active, total) : 0;
        __syncthreads();
        int pos = s_base + woff + incl - cnt;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint32_t w = v[u];
            const int64_t bit0 = (i0 + u * stride) * 32;
            while (w) {
                const int b = __ffs(w) - 1;
                w &= w - 1;
                if (bit0 + b < n_bits) ids[pos] = (int32_t)(bit0 + b);
                ++pos;
            }
        }
        __syncthreads();      // s_warp / s_base are reused by the next trip
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) ctl->stamp[5] = gtimer();
}

}  // namespace be

using namespace be;

extern "C" size_t be_b200_exchange_ctl_bytes(void) { return sizeof(ExchangeCtl); }

extern "C" int be_b200_pack_publish_clear(const void* spikes_local, int spk_kind, int64_t n_local,
                                          void* const* peer_bufs, void* const* peer_flags, int world, int rank,
                                          int64_t words_per_rank, void* ctl, int32_t* reset_counter,
                                          void* clear_ptr, size_t clear_bytes, void* stream) { BE_NVTX_RANGE("be_b200_pack_publish_clear");
    cudaStream_t s = (cudaStream_t)stream;
    BE_REQUIRE(peer_bufs && peer_flags && ctl, "pack_publish: NULL pointer");
    BE_REQUIRE(spikes_local || n_local == 0, "pack_publish: spikes is NULL");   // a rank may own an empty block
    BE_REQUIRE(world >= 1 && world <= 32 && rank >= 0 && rank < world, "pack_publish: bad rank/world");
    BE_REQUIRE(n_local >= 0 && n_local <= words_per_rank * 32, "pack_publish: n_local exceeds the slot");
    int4* zp = nullptr;
    size_t zv = 0;
    if (clear_ptr != nullptr && clear_bytes > 0) {
        if ((((uintptr_t)clear_ptr) & 15u) == 0 && clear_bytes % 16 == 0) { zp = (int4*)clear_ptr; zv = clear_bytes / 16; }
        else BE_CUDA(cudaMemsetAsync(clear_ptr, 0, clear_bytes, s));
    }
    int64_t want = ceil_div(2 * words_per_rank, 256);
    if (zp) want = std::max<int64_t>(want, ceil_div((int64_t)zv, 1024));
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(want, kNumSM * 4));
#define BE_PACK(K)                                                                                              \
    pack_publish_kernel<K><<<blocks, 256, 0, s>>>(spikes_local, n_local, (uint32_t* const*)peer_bufs,            \
                                                  (uint32_t* const*)peer_flags, world, rank, words_per_rank,     \
                                                  (ExchangeCtl*)ctl, reset_counter, zp, zv)
    if (spk_kind == BE_B200_SPK_BOOL) BE_PACK(BE_B200_SPK_BOOL);
    else if (spk_kind == BE_B200_SPK_F32) BE_PACK(BE_B200_SPK_F32);
    else { set_error("pack_publish: unsupported event kind %d", spk_kind); return BE_B200_EINVAL; }
#undef BE_PACK
    BE_LAUNCH_CHECK("pack_publish_kernel");
    return BE_B200_OK;
}

extern "C" int be_b200_pack_publish_reset(const void* spikes_local, int spk_kind, int64_t n_local,
                                          void* const* peer_bufs, void* const* peer_flags, int world, int rank,
                                          int64_t words_per_rank, void* ctl, int32_t* reset_counter, void* stream) { BE_NVTX_RANGE("be_b200_pack_publish_reset");
    return be_b200_pack_publish_clear(spikes_local, spk_kind, n_local, peer_bufs, peer_flags, world, rank,
                                      words_per_rank, ctl, reset_counter, nullptr, 0, stream);
}

extern "C" int be_b200_pack_publish(const void* spikes_local, int spk_kind, int64_t n_local,
                                    void* const* peer_bufs, void* const* peer_flags, int world, int rank,
                                    int64_t words_per_rank, void* ctl, void* stream) { BE_NVTX_RANGE("be_b200_pack_publish");
    return be_b200_pack_publish_clear(spikes_local, spk_kind, n_local, peer_bufs, peer_flags, world, rank,
                                      words_per_rank, ctl, nullptr, nullptr, 0, stream);
}

extern "C" int be_b200_wait_exchange(const void* my_flags, void* const* peer_bufs, int world,
                                     int64_t words_per_rank, uint32_t* gathered_out, void* ctl, void* stream) { BE_NVTX_RANGE("be_b200_wait_exchange");
    BE_REQUIRE(my_flags && peer_bufs && gathered_out && ctl && world >= 1 && world <= 32, "wait_exchange: bad arguments");
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(words_per_rank * world, 1024), kNumSM));
    wait_copy_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const uint32_t*)my_flags, (uint32_t* const*)peer_bufs, world,
                                                              words_per_rank, gathered_out, (ExchangeCtl*)ctl);
    BE_LAUNCH_CHECK("wait_copy_kernel");
    return BE_B200_OK;
}

extern "C" int be_b200_wait_exchange_compact(const void* my_flags, void* const* peer_bufs, int world,
                                             int64_t words_per_rank, int64_t n_bits, uint32_t* gathered_out,
                                             int32_t* ids_and_count, void* clear_ptr, size_t clear_bytes,
                                             void* ctl, void* stream) { BE_NVTX_RANGE("be_b200_wait_exchange_compact");
    BE_REQUIRE(my_flags && peer_bufs && gathered_out && ids_and_count && ctl && world >= 1 && world <= 32,
               "wait_exchange_compact: bad arguments");
    BE_REQUIRE(n_bits >= 0 && n_bits <= words_per_rank * world * 32 && n_bits < ((int64_t)1 << 31),
               "wait_exchange_compact: n_bits outside the gathered vector");
    cudaStream_t s = (cudaStream_t)stream;
    int4* zp = nullptr;
    size_t zv = 0;
    if (clear_ptr != nullptr && clear_bytes > 0) {
        if ((((uintptr_t)clear_ptr) & 15u) == 0 && clear_bytes % 16 == 0) { zp = (int4*)clear_ptr; zv = clear_bytes / 16; }
        else BE_CUDA(cudaMemsetAsync(clear_ptr, 0, clear_bytes, s));
    }
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(words_per_rank * world, 1024), kNumSM));
    wait_copy_compact_kernel<<<blocks, 256, 0, s>>>((const uint32_t*)my_flags, (uint32_t* const*)peer_bufs, world,
                                                    words_per_rank, n_bits, gathered_out, ids_and_count,
                                                    ids_and_count + n_bits, zp, zv, (ExchangeCtl*)ctl);
    BE_LAUNCH_CHECK("wait_copy_compact_kernel");
    return BE_B200_OK;
}
