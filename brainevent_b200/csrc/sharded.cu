// sharded.cu -- one rank's share of a post-sharded projection as a C object: the per-timestep step
// (spike exchange fused with the consumer) and the host-buffer pipeline around it.
//
// Net-new relative to the reference (single-device).  Layout: rank g owns post block
// [g*per, (g+1)*per) (per = 32 * words_per_rank) and stores only synapses ending there (pre-major rows
// over ALL pre neurons, local post ids).  Per timestep:
//
//   stored connectivity (CSR shard / ELL shard), TWO kernels:
//     1. pack_publish_kernel (exchange.cu): bit-pack my spike block into my own symmetric buffer,
//        publish the step number to every peer, clear my output block, reset the chunk cursor;
//     2. exchange_scatter_kernel (here): the all-gather is fused INTO the scatter.  The global spike
//        vector is cut into chunks of kChunkWords words; CTAs draw chunks from a device-side cursor,
//        my own block's chunks first (no wait: they are local), then peer rank+1, rank+2, ... as
//        their flags arrive.  A CTA pulls the chunk's words (over NVLink for a peer), turns them
//        into a shared-memory list of active rows and scatters those rows at once.  No gathered
//        vector, no global id list, no separate compaction; the flag / pull latency of the peers is
//        hidden behind the scatter of the local rows, and the dynamic cursor balances the Poisson
//        spread of active rows over the CTAs.
//   JITC / dense consumers: pack + publish, wait + pull (+ compaction for scatter orders), consumer.
//
// The host pipeline (be_b200_sharded_submit_host / _wait / _run_host) is the multi-GPU twin of
// be_b200_projection_submit_host: per slot one CUDA graph of the device step, copy-in / step /
// copy-out on three streams, `slots` timesteps in flight, no caller code between timesteps in
// _run_host.
#include "conn.cuh"
#include "exchange.cuh"

#include <algorithm>
#include <cstdlib>
#include <new>
#include <string_view>
#include <type_traits>

namespace be {

constexpr int kChunkWords = 64;              // LARGEST work item: 2048 source neurons (two words per producer lane); the launch
                                             // halves it (32, 16 words) until every CTA can expect >= 4 draws from the cursor --
                                             // with one chunk per CTA the Poisson spread of active rows per chunk is the runtime

// Chunk pipeline inside a CTA: warp 0 PRODUCES the row list of chunk k+1 (cursor draw, flag wait for a peer's
// first chunk, pull of 64 words, ballot-free compaction: two words per lane, one warp scan) into one half of a
// double-buffered shared-memory list while ALL warps -- warp 0 joins when it is done -- CONSUME chunk k: groups of
// LPR lanes take rows from the list through a shared ticket counter (no static row -> warp assignment, so the
// only barrier is the one per chunk swap) and keep the next row's extent in flight.
template <typename T, bool HOMO, int VEC, int LPR, typename Rows>
__global__ void __launch_bounds__(256, 8) exchange_scatter_kernel(Rows rows, const int32_t* __restrict__ indices,
                                                               const T* __restrict__ weights,
                                                               const uint32_t* __restrict__ my_flags,
                                                               uint32_t* const* __restrict__ peer_bufs, int world,
                                                               int rank, int64_t wpr, int64_t n_bits, int cw,
                                                               T* __restrict__ out, ExchangeCtl* ctl) {
    using A = typename Acc<T>::type;
    __shared__ int s_ids[2][kChunkWords * 32];
    __shared__ int s_n[2];                 // rows in the list; -1 = no more chunks / exchange gave up
    __shared__ int s_ticket[2];
    __shared__ uint32_t s_seen;
    const uint32_t t = ctl->step;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const A w0 = HOMO ? to_acc<T>(__ldg(weights)) : A(0);
    const int chunks_per_rank = (int)ceil_div(wpr, cw);       // cw = words per chunk: 16, 32 or 64
    const int total_chunks = chunks_per_rank * world;
    const int64_t half = (int64_t)(t & 1u) * wpr * world;
    constexpr int GPW = 32 / LPR;          // row groups per warp
    const int gl = lane % LPR;
    const unsigned gmask = (LPR >= 32) ? 0xffffffffu : (((1u << LPR) - 1u) << ((lane / LPR) * LPR));

    // warp 0: build list `b` from the next chunk of the cursor
    auto produce = [&](int b) {
        int c = 0;
        if (lane == 0) c = (int)atomicAdd(&ctl->next_chunk, 1u);
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c >= total_chunks) { if (lane == 0) s_n[b] = -1; return; }
        const int o = c / chunks_per_rank, lc = c - o * chunks_per_rank;
        int p = rank + o;
        if (p >= world) p -= world;
        if (!((s_seen >> p) & 1u)) {       // this CTA's first chunk of peer p: wait for its flag (bounded by wall time)
            int bad = 0;
            if (lane == 0) {
                const unsigned long long limit = (unsigned long long)(ctl->spin_ms ? ctl->spin_ms : kDefaultSpinMs) * 1000000ull;
                unsigned long long t0 = 0;
                unsigned int polls = 0;
                while ((int32_t)(ld_acquire_sys(my_flags + p) - t) < 0) {
                    if ((++polls & 1023u) == 0) {
                        const unsigned long long now = gtimer();
                        if (t0 == 0) t0 = now;
                        else if (now - t0 > limit) { ctl->timeout = 1u; bad = 1; break; }
                    }
                    __nanosleep(20);
                }
                if (!bad) s_seen |= 1u << p;
            }
            bad = __shfl_sync(0xffffffffu, bad, 0);
            if (bad) { if (lane == 0) s_n[b] = -1; return; }
        }
        const int64_t w_first = (int64_t)lc * cw + 2 * lane;
        uint32_t wd[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int64_t wi = w_first + q;
            uint32_t v = 0;
            if (2 * lane + q < cw && wi < wpr) {
                v = ld_relaxed_sys(peer_bufs[p] + half + (int64_t)p * wpr + wi);
                const int64_t bit0 = ((int64_t)p * wpr + wi) * 32;
                if (bit0 + 32 > n_bits) v = (bit0 >= n_bits) ? 0u : (v & ((1u << (int)(n_bits - bit0)) - 1u));
            }
            wd[q] = v;
        }
        const int cnt = __popc(wd[0]) + __popc(wd[1]);
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        int pos = incl - cnt;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            uint32_t v = wd[q];
            const int base = (int)(((int64_t)p * wpr + w_first + q) * 32);
            while (v) {
                const int bb = __ffs(v) - 1;
                v &= v - 1;
                s_ids[b][pos++] = base + bb;
            }
        }
        if (lane == 31) { s_n[b] = incl; s_ticket[b] = 0; }
    };

    if (threadIdx.x == 0) s_seen = 1u << rank;               // my own words were packed earlier in stream order
    __syncthreads();
    if (warp == 0) {
        // sticky: an earlier exchange gave up -> no list at all
        int bad = 0;
        if (lane == 0) bad = (int)*(volatile uint32_t*)&ctl->timeout;
        bad = __shfl_sync(0xffffffffu, bad, 0);
        if (bad) { if (lane == 0) s_n[0] = -1; }
        else produce(0);
    }
    int cur = 0;
    for (;;) {
        __syncthreads();                                      // list `cur` is complete; list `cur ^ 1` is free again
        const int n_rows = s_n[cur];
        if (n_rows < 0) break;
        if (warp == 0) produce(cur ^ 1);
        // consume: tickets of GPW rows per warp draw, the next draw's row extents requested before this one is scattered
        const int sub = lane / LPR;
        int a = 0;
        if (lane == 0) a = atomicAdd(&s_ticket[cur], GPW);
        a = __shfl_sync(0xffffffffu, a, 0) + sub;
        int64_t s_nx = 0, e_nx = 0;
        if (a < n_rows) rows.get(s_ids[cur][a], s_nx, e_nx);
        while (a - sub < n_rows) {                            // warp-uniform
            const int64_t s = s_nx, e = e_nx;
            const bool mine = a < n_rows;
            int an = 0;
            if (lane == 0) an = atomicAdd(&s_ticket[cur], GPW);
            an = __shfl_sync(0xffffffffu, an, 0) + sub;
            if (an < n_rows) rows.get(s_ids[cur][an], s_nx, e_nx);
            if (mine)
                for_each_synapse<T, HOMO, 0, VEC, LPR>(s, e, gl, indices, weights, nullptr, w0,
                                                       [&](int col, A w) { red_add(out + col, w); });
            a = an;
        }
        (void)gmask;
        cur ^= 1;
    }
}

}  // namespace be

using namespace be;

struct be_b200_sharded {
    int world, rank, device;
    int64_t n_neurons, n_local, n_out, wpr;
    // exchange state
    void** d_peer_bufs;       // device array [world] of peer-mapped buffers
    void** d_peer_flags;      // device array [world]: flag words inside each peer's buffer
    uint32_t* my_flags;       // flags inside MY buffer (local address)
    ExchangeCtl* ctl;
    int own_ctl;
    uint32_t* gathered;       // lazily allocated: JITC / dense consumers
    int32_t* ids;             // lazily allocated: scatter-order JITC / dense
    // consumer
    int conn;                 // -1 unset, BE_B200_CONN_*
    const void* weights; int w_dtype, homo;
    const int32_t* indices; const void* indptr; int indptr_is64; int64_t n_conn, nse;
    int jit_kind, corder; const void *p0, *p1; const int32_t *clen, *seed; int64_t n_post_global, chunk_size;
    void* ws; size_t ws_bytes;
    int fused;                // stored connectivity: 1 = exchange fused into the scatter (default)
    int lpr;                  // lanes per row of the fused scatter (0 = by mean row length)
    int chunk_words;          // words per work item of the fused scatter (0 = by network size; 16, 32, 64)
    // host pipeline
    int slots, slot;
    uint8_t** d_in; float** d_out;
    cudaGraphExec_t* graphs;
    cudaStream_t s_in, s_main, s_out;
    cudaEvent_t *ev_in, *ev_done, *ev_out;
    int* inflight;
};

static size_t wt_bytes(int w_dtype) { static const size_t k[4] = {4, 8, 2, 2}; return k[w_dtype & 3]; }

extern "C" be_b200_sharded* be_b200_sharded_create(int world, int rank, int64_t n_neurons, int64_t n_out,
                                                   int64_t words_per_rank, void* const* peer_bufs_host,
                                                   void* ctl) {
    if (world < 1 || world > 32 || rank < 0 || rank >= world || n_neurons < 0 || n_out < 0 || words_per_rank < 1 ||
        !peer_bufs_host || n_neurons > words_per_rank * 32 * world || n_neurons >= (1ll << 31)) {
        set_error("sharded_create: bad arguments");
        return nullptr;
    }
    be_b200_sharded* h = new (std::nothrow) be_b200_sharded();
    if (!h) { set_error("out of host memory"); return nullptr; }
    h->world = world; h->rank = rank; h->n_neurons = n_neurons; h->n_out = n_out; h->wpr = words_per_rank;
    h->conn = -1; h->fused = 1;
    cudaGetDevice(&h->device);
    const int64_t per = words_per_rank * 32;
    h->n_local = std::max<int64_t>(0, std::min<int64_t>(per, n_neurons - per * rank));
    void* flags[32];
    const size_t n_buf = (size_t)2 * world * words_per_rank;
    for (int p = 0; p < world; ++p) {
        if (!peer_bufs_host[p]) { set_error("sharded_create: peer buffer %d is NULL", p); delete h; return nullptr; }
        flags[p] = (char*)peer_bufs_host[p] + 4 * n_buf;
    }
    h->my_flags = (uint32_t*)flags[rank];
    bool ok = cudaMalloc((void**)&h->d_peer_bufs, sizeof(void*) * world) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&h->d_peer_flags, sizeof(void*) * world) == cudaSuccess;
    // `ctl` (nullable): an existing zero-initialised control block of be_b200_exchange_ctl_bytes() to share the
    // step counter with direct be_b200_pack_publish / be_b200_wait_exchange calls on the same buffers
    if (ctl) h->ctl = (ExchangeCtl*)ctl;
    else {
        h->own_ctl = 1;
        ok = ok && cudaMalloc((void**)&h->ctl, sizeof(ExchangeCtl)) == cudaSuccess;
        ok = ok && cudaMemset(h->ctl, 0, sizeof(ExchangeCtl)) == cudaSuccess;
    }
    ok = ok && cudaMemcpy(h->d_peer_bufs, peer_bufs_host, sizeof(void*) * world, cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(h->d_peer_flags, flags, sizeof(void*) * world, cudaMemcpyHostToDevice) == cudaSuccess;
    h->ws_bytes = be_b200_workspace_bytes(n_neurons, 1);
    ok = ok && cudaMalloc(&h->ws, h->ws_bytes) == cudaSuccess;
    ok = ok && cudaMemset(h->ws, 0, h->ws_bytes) == cudaSuccess;
    ok = ok && cudaDeviceSynchronize() == cudaSuccess;      // callers enqueue on non-blocking streams
    if (!ok) {
        set_error("sharded_create: device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        be_b200_sharded_destroy(h);
        return nullptr;
    }
    return h;
}

/* bytes of the symmetric buffer every rank must allocate (zeroed) before the rendezvous */
extern "C" size_t be_b200_sharded_symm_bytes(int world, int64_t words_per_rank) {
    return (size_t)4 * ((size_t)2 * world * words_per_rank + 64);
}

extern "C" int be_b200_sharded_set_fcn(be_b200_sharded* h, const void* weights, int w_dtype, int homo,
                                       const int32_t* indices, int64_t n_rows, int64_t n_conn) { BE_NVTX_RANGE("be_b200_sharded_set_fcn");
    BE_REQUIRE(h && weights && (indices || n_rows * n_conn == 0), "sharded_set_fcn: NULL pointer");
    BE_REQUIRE(n_rows == h->n_neurons, "sharded_set_fcn: %lld rows but the exchange carries %lld neurons",
               (long long)n_rows, (long long)h->n_neurons);
    BE_REQUIRE(w_dtype >= 0 && w_dtype < 4 && n_conn >= 0, "sharded_set_fcn: bad dtype / n_conn");
    h->conn = BE_B200_CONN_FCN; h->weights = weights; h->w_dtype = w_dtype; h->homo = homo;
    h->indices = indices; h->n_conn = n_conn; h->nse = n_rows * n_conn;
    return BE_B200_OK;
}

extern "C" int be_b200_sharded_set_csr(be_b200_sharded* h, const void* weights, int w_dtype, int homo,
                                       const int32_t* indices, const void* indptr, int indptr_is64,
                                       int64_t n_rows, int64_t nse) { BE_NVTX_RANGE("be_b200_sharded_set_csr");
    BE_REQUIRE(h && weights && indptr && (indices || nse == 0), "sharded_set_csr: NULL pointer");
    BE_REQUIRE(n_rows == h->n_neurons, "sharded_set_csr: %lld rows but the exchange carries %lld neurons",
               (long long)n_rows, (long long)h->n_neurons);
    BE_REQUIRE(w_dtype >= 0 && w_dtype < 4 && nse >= 0, "sharded_set_csr: bad dtype / nse");
    h->conn = BE_B200_CONN_CSR; h->weights = weights; h->w_dtype = w_dtype; h->homo = homo;
    h->indices = indices; h->indptr = indptr; h->indptr_is64 = indptr_is64; h->nse = nse;
    return BE_B200_OK;
}

extern "C" int be_b200_sharded_set_jitc(be_b200_sharded* h, int jit_kind, const void* p0, const void* p1,
                                        int w_dtype, const int32_t* clen, const int32_t* seed,
                                        int64_t n_post_global, int64_t chunk_size, int corder) { BE_NVTX_RANGE("be_b200_sharded_set_jitc");
    BE_REQUIRE(h && p0 && clen && seed && (jit_kind == BE_B200_JIT_SCALAR || p1), "sharded_set_jitc: NULL pointer");
    BE_REQUIRE(chunk_size >= 1 && n_post_global >= h->n_out, "sharded_set_jitc: bad shape");
    h->conn = BE_B200_CONN_JITC; h->jit_kind = jit_kind; h->p0 = p0; h->p1 = p1; h->w_dtype = w_dtype;
    h->clen = clen; h->seed = seed; h->n_post_global = n_post_global; h->chunk_size = chunk_size; h->corder = corder;
    return BE_B200_OK;
}

extern "C" int be_b200_sharded_set_dense(be_b200_sharded* h, const void* W, int w_dtype) { BE_NVTX_RANGE("be_b200_sharded_set_dense");
    BE_REQUIRE(h && (W || h->n_neurons * h->n_out == 0), "sharded_set_dense: NULL pointer");
    h->conn = BE_B200_CONN_DENSE; h->weights = W; h->w_dtype = w_dtype;
    return BE_B200_OK;
}

extern "C" int be_b200_sharded_set_option(be_b200_sharded* h, const char* name, int64_t value) {
    BE_REQUIRE(h && name, "sharded_set_option: NULL pointer");
    const std::string_view n(name);
    if (n == "fused") h->fused = value != 0;
    else if (n == "lanes_per_row") {
        BE_REQUIRE(value == 0 || value == 8 || value == 16 || value == 32, "sharded_set_option: lanes_per_row must be 0, 8, 16 or 32");
        h->lpr = (int)value;
    } else if (n == "chunk_words") {
        BE_REQUIRE(value == 0 || value == 16 || value == 32 || value == 64, "sharded_set_option: chunk_words must be 0, 16, 32 or 64");
        h->chunk_words = (int)value;
    } else if (n == "spin_ms") {
        const uint32_t ms = (uint32_t)std::max<int64_t>(0, value);
        BE_CUDA(cudaMemcpy(&h->ctl->spin_ms, &ms, 4, cudaMemcpyHostToDevice));
    } else { set_error("sharded_set_option: unknown option %s", name); return BE_B200_EINVAL; }
    return BE_B200_OK;
}

namespace {

template <typename T, bool HOMO, typename Rows>
int launch_exchange_scatter(be_b200_sharded* h, const Rows& rows, int vec, T* out, cudaStream_t s) {
    const double avg = h->n_neurons > 0 ? (double)h->nse / (double)h->n_neurons : 0.0;
    int lpr = h->lpr ? h->lpr : (avg < 40.0 ? 8 : (avg < 96.0 ? 16 : 32));
    if (const char* e = getenv("BE_B200_SHARDED_LPR")) { const int v = atoi(e); if (v == 8 || v == 16 || v == 32) lpr = v; }   // A/B switch
    if (const char* e = getenv("BE_B200_SHARDED_CW")) { const int v = atoi(e); if (v == 16 || v == 32 || v == 64) h->chunk_words = v; }
    int cw = kChunkWords;
    if (h->chunk_words) cw = h->chunk_words;
    else while (cw > 16 && ceil_div(h->wpr, cw) * h->world < (int64_t)4 * kNumSM * 8) cw >>= 1;
    const int chunks = (int)ceil_div(h->wpr, cw) * h->world;
    const int blocks = std::max(1, std::min(chunks, kNumSM * 8));
#define BE_XS(VEC, LPR)                                                                                       \
    exchange_scatter_kernel<T, HOMO, VEC, LPR, Rows><<<blocks, 256, 0, s>>>(                                   \
        rows, h->indices, (const T*)h->weights, h->my_flags, (uint32_t* const*)h->d_peer_bufs, h->world,       \
        h->rank, h->wpr, h->n_neurons, cw, out, h->ctl)
#define BE_XS_L(VEC)                                                                                           \
    do { if (lpr == 8) BE_XS(VEC, 8); else if (lpr == 16) BE_XS(VEC, 16); else BE_XS(VEC, 32); } while (0)
    {
        ProfileScope prof(s);
        if (vec == 2) BE_XS_L(2); else if (vec == 1) BE_XS_L(1); else BE_XS_L(0);
    }
#undef BE_XS_L
#undef BE_XS
    BE_LAUNCH_CHECK("exchange_scatter_kernel");
    return BE_B200_OK;
}

template <typename T>
int step_stored_fused(be_b200_sharded* h, T* out, cudaStream_t s) {
    ConnArgs c{h->weights, h->w_dtype, h->homo, h->indices, nullptr, 0, 0.0};
    if (h->conn == BE_B200_CONN_FCN) {
        FcnRows rows{h->n_conn};
        const int vec = vec_mode(rows, c);
        return h->homo ? launch_exchange_scatter<T, true>(h, rows, vec, out, s) : launch_exchange_scatter<T, false>(h, rows, vec, out, s);
    }
    if (h->indptr_is64) {
        CsrRows<int64_t> rows{(const int64_t*)h->indptr};
        const int vec = vec_mode(rows, c);
        return h->homo ? launch_exchange_scatter<T, true>(h, rows, vec, out, s) : launch_exchange_scatter<T, false>(h, rows, vec, out, s);
    }
    CsrRows<int32_t> rows{(const int32_t*)h->indptr};
    const int vec = vec_mode(rows, c);
    return h->homo ? launch_exchange_scatter<T, true>(h, rows, vec, out, s) : launch_exchange_scatter<T, false>(h, rows, vec, out, s);
}

int ensure_gathered(be_b200_sharded* h, bool need_ids) {
    if (!h->gathered) {
        BE_CUDA(cudaMalloc((void**)&h->gathered, (size_t)4 * h->wpr * h->world + 16));
        BE_CUDA(cudaMemset(h->gathered, 0, (size_t)4 * h->wpr * h->world + 16));
    }
    if (need_ids && !h->ids) {
        BE_CUDA(cudaMalloc((void**)&h->ids, (size_t)4 * (h->n_neurons + 1)));
        BE_CUDA(cudaMemset(h->ids, 0, (size_t)4 * (h->n_neurons + 1)));
    }
    BE_CUDA(cudaDeviceSynchronize());
    return BE_B200_OK;
}

}  // namespace

// One timestep on the device: spikes_local = the spikes of the n_local neurons this rank owns (kind BOOL or
// F32), out = this rank's post block [n_out] in the weight dtype.  Enqueued on `stream`, never synchronises;
// every rank must call it the same number of times.  The step has two phases that a caller may also issue
// separately (publish early, consume later; tests emulate several ranks on one GPU by running every rank's
// publish before any rank's consume, so that no wait ever blocks):
//   publish: bit-pack + flag the peers (+ clear `out`, reset the cursors)      be_b200_sharded_publish
//   consume: pull the peers' words and propagate them into `out`              be_b200_sharded_consume
static bool sharded_is_scatter(const be_b200_sharded* h) {
    return h->conn == BE_B200_CONN_FCN || h->conn == BE_B200_CONN_CSR || h->conn == BE_B200_CONN_DENSE ||
           (h->conn == BE_B200_CONN_JITC && !h->corder);
}
static bool sharded_is_fused(const be_b200_sharded* h) {
    return (h->conn == BE_B200_CONN_FCN || h->conn == BE_B200_CONN_CSR) && h->fused;
}

extern "C" int be_b200_sharded_publish(be_b200_sharded* h, const void* spikes_local, int spk_kind, void* out,
                                       void* stream) { BE_NVTX_RANGE("be_b200_sharded_publish");
    BE_REQUIRE(h && out && h->conn >= 0, "sharded_publish: projection not configured");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t out_bytes = (size_t)h->n_out * wt_bytes(h->w_dtype);
    if (sharded_is_fused(h))
        return be_b200_pack_publish_clear(spikes_local, spk_kind, h->n_local, h->d_peer_bufs, h->d_peer_flags, h->world,
                                          h->rank, h->wpr, h->ctl, nullptr, out, out_bytes, s);
    const bool scatter = sharded_is_scatter(h);
    if (!h->gathered || (scatter && !h->ids)) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(s, &cs);
        BE_REQUIRE(cs == cudaStreamCaptureStatusNone, "sharded_step: run one step outside graph capture first (lazy buffers)");
        const int rc = ensure_gathered(h, scatter);
        if (rc) return rc;
    }
    if (scatter)
        return be_b200_pack_publish_reset(spikes_local, spk_kind, h->n_local, h->d_peer_bufs, h->d_peer_flags, h->world,
                                          h->rank, h->wpr, h->ctl, h->ids + h->n_neurons, s);
    return be_b200_pack_publish(spikes_local, spk_kind, h->n_local, h->d_peer_bufs, h->d_peer_flags, h->world, h->rank,
                                h->wpr, h->ctl, s);
}

extern "C" int be_b200_sharded_consume(be_b200_sharded* h, void* out, void* stream) { BE_NVTX_RANGE("be_b200_sharded_consume");
    BE_REQUIRE(h && out && h->conn >= 0, "sharded_consume: projection not configured");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t out_bytes = (size_t)h->n_out * wt_bytes(h->w_dtype);
    if (sharded_is_fused(h)) {
        if (h->n_out == 0 || h->nse == 0) return BE_B200_OK;
        BE_DISPATCH_WT(h->w_dtype, T, { return step_stored_fused<T>(h, (T*)out, s); });
        return BE_B200_OK;
    }
    // scatter consumers take the id list from the wait kernel (which also clears `out`); gather consumers the words
    const bool scatter = sharded_is_scatter(h);
    BE_REQUIRE(h->gathered && (!scatter || h->ids), "sharded_consume: publish first");
    int rc;
    if (scatter)
        rc = be_b200_wait_exchange_compact(h->my_flags, h->d_peer_bufs, h->world, h->wpr, h->n_neurons, h->gathered,
                                           h->ids, out, out_bytes, h->ctl, s);
    else
        rc = be_b200_wait_exchange(h->my_flags, h->d_peer_bufs, h->world, h->wpr, h->gathered, h->ctl, s);
    if (rc) return rc;
    const int kind = scatter ? BE_B200_SPK_COMPACT_CLEARED : BE_B200_SPK_PACKED;
    const void* aux = scatter ? (const void*)h->ids : nullptr;
    const int64_t lo = (int64_t)h->rank * h->wpr * 32;
    switch (h->conn) {
        case BE_B200_CONN_FCN:
            return be_b200_binary_fcnmv(h->weights, h->w_dtype, h->homo, h->indices, h->n_neurons, h->n_conn, h->gathered,
                                        aux, kind, h->n_neurons, h->n_out, 1, out, h->ws, h->ws_bytes, s);
        case BE_B200_CONN_CSR:
            return be_b200_binary_csrmv(h->weights, h->w_dtype, h->homo, h->indices, h->indptr, h->indptr_is64, nullptr, 0,
                                        h->gathered, aux, kind, h->n_neurons, h->n_out, h->nse, 1, out, h->ws, h->ws_bytes, s);
        case BE_B200_CONN_JITC:
            // the GLOBAL shape keys the RNG walk; the rank's post block is a row window of the gather
            // order (rows = post) or a column window of the scatter order (columns = post)
            if (h->corder)
                return be_b200_binary_jitmv(h->jit_kind, h->p0, h->p1, h->w_dtype, h->clen, h->seed, h->gathered, aux, kind,
                                            h->n_neurons, h->n_post_global, h->chunk_size, 1, lo, h->n_out, 0, -1, out,
                                            h->ws, h->ws_bytes, s);
            return be_b200_binary_jitmv(h->jit_kind, h->p0, h->p1, h->w_dtype, h->clen, h->seed, h->gathered, aux, kind,
                                        h->n_neurons, h->n_post_global, h->chunk_size, 0, 0, -1, lo, lo + h->n_out, out,
                                        h->ws, h->ws_bytes, s);
        case BE_B200_CONN_DENSE:
            return be_b200_binary_densemv(h->weights, h->w_dtype, h->n_neurons, h->n_out, h->gathered, aux, kind, 1, out,
                                          h->ws, h->ws_bytes, s);
    }
    set_error("sharded_step: unknown consumer %d", h->conn);
    return BE_B200_EINVAL;
}

extern "C" int be_b200_sharded_step(be_b200_sharded* h, const void* spikes_local, int spk_kind, void* out,
                                    void* stream) { BE_NVTX_RANGE("be_b200_sharded_step");
    const int rc = be_b200_sharded_publish(h, spikes_local, spk_kind, out, stream);
    return rc ? rc : be_b200_sharded_consume(h, out, stream);
}

// 1 when a flag wait of this rank ever gave up (sticky).  Synchronises the device: call it at
// wait() / close() boundaries, not per timestep.
extern "C" int be_b200_sharded_timed_out(be_b200_sharded* h) {
    if (!h) return 0;
    uint32_t v = 0;
    if (cudaMemcpy(&v, &h->ctl->timeout, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
    return v != 0;
}

// ---- host-buffer pipeline ---------------------------------------------------------------------------------
extern "C" int be_b200_sharded_host_init(be_b200_sharded* h, int slots, int use_graph) { BE_NVTX_RANGE("be_b200_sharded_host_init");
    BE_REQUIRE(h && h->conn >= 0 && slots >= 1 && slots <= 16, "sharded_host_init: bad arguments");
    BE_REQUIRE(h->w_dtype == BE_B200_F32, "sharded_host_init: the host pipeline carries float32 outputs");
    BE_REQUIRE(h->slots == 0, "sharded_host_init: already initialised");
    h->d_in = new uint8_t*[slots](); h->d_out = new float*[slots]();
    h->graphs = new cudaGraphExec_t[slots]();
    h->ev_in = new cudaEvent_t[slots](); h->ev_done = new cudaEvent_t[slots](); h->ev_out = new cudaEvent_t[slots]();
    h->inflight = new int[slots]();
    h->slots = slots;
    BE_CUDA(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
    BE_CUDA(cudaStreamCreateWithFlags(&h->s_main, cudaStreamNonBlocking));
    BE_CUDA(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < slots; ++i) {
        BE_CUDA(cudaMalloc((void**)&h->d_in[i], (size_t)std::max<int64_t>(h->n_local, 16)));
        BE_CUDA(cudaMemsetAsync(h->d_in[i], 0, (size_t)std::max<int64_t>(h->n_local, 16), h->s_main));
        BE_CUDA(cudaMalloc((void**)&h->d_out[i], (size_t)std::max<int64_t>(h->n_out, 4) * sizeof(float)));
        BE_CUDA(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
        BE_CUDA(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
        // the host blocks on ev_out when it re-uses a slot: a blocking event lets that thread sleep, which
        // matters when 8 ranks (and their driver threads) share the host's cores
        BE_CUDA(cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming | cudaEventBlockingSync));
    }
    BE_CUDA(cudaStreamSynchronize(h->s_main));
    if (use_graph) {
        // Every rank captures the same number of steps, and capture enqueues nothing, so the step counters
        // stay aligned.  Lazy buffers of the unfused paths must exist before capture: the caller runs
        // be_b200_sharded_step once eagerly (collectively) first when the consumer is not fused.
        for (int i = 0; i < slots; ++i) {
            cudaGraph_t g = nullptr;
            BE_CUDA(cudaStreamBeginCapture(h->s_main, cudaStreamCaptureModeThreadLocal));
            const int rc = be_b200_sharded_step(h, h->d_in[i], BE_B200_SPK_BOOL, h->d_out[i], h->s_main);
            const cudaError_t e = cudaStreamEndCapture(h->s_main, &g);
            if (rc) { if (g) cudaGraphDestroy(g); return rc; }
            if (e != cudaSuccess) return cuda_fail(e, "cudaStreamEndCapture");
            const cudaError_t e2 = cudaGraphInstantiate(&h->graphs[i], g, 0);
            cudaGraphDestroy(g);
            if (e2 != cudaSuccess) return cuda_fail(e2, "cudaGraphInstantiate");
        }
    }
    return BE_B200_OK;
}

// spikes_host: pinned uint8[n_local]; out_host: pinned float[n_out].  Both must stay untouched until
// be_b200_sharded_wait() (or the `slots`-th next submit) returns.  Collective: every rank submits the
// same number of timesteps.
extern "C" int be_b200_sharded_submit_host(be_b200_sharded* h, const uint8_t* spikes_host, float* out_host) { BE_NVTX_RANGE("be_b200_sharded_submit_host");
    BE_REQUIRE(h && h->slots > 0 && out_host && (spikes_host || h->n_local == 0), "sharded_submit_host: bad arguments");
    const int i = h->slot;
    if (h->inflight[i]) { BE_CUDA(cudaEventSynchronize(h->ev_out[i])); h->inflight[i] = 0; }
    if (h->n_local > 0) BE_CUDA(cudaMemcpyAsync(h->d_in[i], spikes_host, (size_t)h->n_local, cudaMemcpyHostToDevice, h->s_in));
    BE_CUDA(cudaEventRecord(h->ev_in[i], h->s_in));
    BE_CUDA(cudaStreamWaitEvent(h->s_main, h->ev_in[i], 0));
    if (h->graphs[i]) BE_CUDA(cudaGraphLaunch(h->graphs[i], h->s_main));
    else {
        const int rc = be_b200_sharded_step(h, h->d_in[i], BE_B200_SPK_BOOL, h->d_out[i], h->s_main);
        if (rc) return rc;
    }
    BE_CUDA(cudaEventRecord(h->ev_done[i], h->s_main));
    BE_CUDA(cudaStreamWaitEvent(h->s_out, h->ev_done[i], 0));
    if (h->n_out > 0) BE_CUDA(cudaMemcpyAsync(out_host, h->d_out[i], (size_t)h->n_out * sizeof(float), cudaMemcpyDeviceToHost, h->s_out));
    BE_CUDA(cudaEventRecord(h->ev_out[i], h->s_out));
    h->inflight[i] = 1;
    h->slot = (i + 1) % h->slots;
    return BE_B200_OK;
}

extern "C" int be_b200_sharded_wait(be_b200_sharded* h) { BE_NVTX_RANGE("be_b200_sharded_wait");
    BE_REQUIRE(h, "sharded_wait: NULL pointer");
    for (int i = 0; i < h->slots; ++i)
        if (h->inflight[i]) { BE_CUDA(cudaEventSynchronize(h->ev_out[i])); h->inflight[i] = 0; }
    if (be_b200_sharded_timed_out(h)) {
        set_error("sharded: a spike-exchange wait timed out (a peer never published); results since then are invalid");
        return BE_B200_EINTERNAL;
    }
    return BE_B200_OK;
}

// n timesteps back to back without returning to the caller: timestep t reads spikes_host[t % n_in] and
// writes out_host[t % n_out_bufs]; returns after the last result has landed.
extern "C" int be_b200_sharded_run_host(be_b200_sharded* h, int64_t n_steps, const uint8_t* const* spikes_host,
                                        int n_in, float* const* out_host, int n_out_bufs) { BE_NVTX_RANGE("be_b200_sharded_run_host");
    BE_REQUIRE(h && spikes_host && out_host && n_in >= 1 && n_out_bufs >= 1, "sharded_run_host: bad arguments");
    BE_REQUIRE(n_out_bufs >= h->slots, "sharded_run_host: needs at least `slots` output buffers (results in flight)");
    for (int64_t t = 0; t < n_steps; ++t) {
        const int rc = be_b200_sharded_submit_host(h, spikes_host[t % n_in], out_host[t % n_out_bufs]);
        if (rc) return rc;
    }
    return be_b200_sharded_wait(h);
}

extern "C" void be_b200_sharded_destroy(be_b200_sharded* h) {
    if (!h) return;
    cudaDeviceSynchronize();
    for (int i = 0; i < h->slots; ++i) {
        if (h->graphs && h->graphs[i]) cudaGraphExecDestroy(h->graphs[i]);
        if (h->d_in) cudaFree(h->d_in[i]);
        if (h->d_out) cudaFree(h->d_out[i]);
        if (h->ev_in && h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
        if (h->ev_done && h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
        if (h->ev_out && h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
    }
    delete[] h->d_in; delete[] h->d_out; delete[] h->graphs; delete[] h->ev_in; delete[] h->ev_done; delete[] h->ev_out;
    delete[] h->inflight;
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_main) cudaStreamDestroy(h->s_main);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    cudaFree(h->d_peer_bufs); cudaFree(h->d_peer_flags); cudaFree(h->gathered); cudaFree(h->ids);
    if (h->own_ctl) cudaFree(h->ctl);
    cudaFree(h->ws);
    delete h;
}
