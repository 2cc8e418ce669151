// session.cu -- host-buffer projection: connectivity resident in HBM, per call host spikes in,
// host result out.  This is the end-to-end entry a host-side caller (the reference's plugin
// boundary) would bind when its arrays live in host memory.
#include "common.cuh"

#include <cmath>
#include <new>

struct be_b200_projection {
    int kind;  // 0 csr, 1 fcn, 2 jitc
    int homo;
    int64_t n_pre, n_post, n_conn, nse, chunk_size;
    int jit_kind, corder;
    void* d_weights;
    int32_t* d_indices;
    int32_t* d_indptr;
    float* d_p;        // p0, p1
    int32_t* d_clen_seed;
    uint8_t* d_events;
    float* d_out;
    void* d_ws;
    size_t ws_bytes;
    cudaStream_t stream;
    int borrowed;  // connectivity pointers belong to the caller
    // pipelined submit/wait: kSlots sets of device-side event/output buffers, copy streams either
    // side of the compute stream so PCIe transfers of neighbouring calls overlap the kernels.
    // One slot's cycle is H2D + kernels + D2H; four slots keep the D2H engine (the bound for a
    // 4-byte-per-post result) busy back to back.
    uint8_t* d_ev2[4];
    float* d_out2[4];
    cudaStream_t s_in, s_out;
    cudaEvent_t ev_in[4], ev_done[4], ev_out[4];
    int slot, inflight[4];
};
constexpr int kSlots = 4;

using namespace be;

static be_b200_projection* proj_new(int kind, int64_t n_pre, int64_t n_post) {
    be_b200_projection* p = new (std::nothrow) be_b200_projection();
    if (!p) { set_error("out of host memory"); return nullptr; }
    p->kind = kind; p->n_pre = n_pre; p->n_post = n_post;
    bool ok = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) == cudaSuccess;
    p->ws_bytes = be_b200_workspace_bytes(n_pre, 1);
    ok = ok && cudaMalloc(&p->d_ws, p->ws_bytes) == cudaSuccess;
    // zeroed ON the compute stream and waited for: the non-blocking streams below never synchronise with
    // the legacy default stream, so a plain cudaMemset would not be ordered before the first compaction
    ok = ok && cudaMemsetAsync(p->d_ws, 0, p->ws_bytes, p->stream) == cudaSuccess;
    ok = ok && cudaStreamSynchronize(p->stream) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&p->d_events, (size_t)std::max<int64_t>(n_pre, 16)) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&p->d_out, (size_t)std::max<int64_t>(n_post, 4) * sizeof(float)) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < kSlots && ok; ++i) {
        ok = ok && cudaMalloc((void**)&p->d_ev2[i], (size_t)std::max<int64_t>(n_pre, 16)) == cudaSuccess;
        ok = ok && cudaMalloc((void**)&p->d_out2[i], (size_t)std::max<int64_t>(n_post, 4) * sizeof(float)) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&p->ev_in[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&p->ev_done[i], cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&p->ev_out[i], cudaEventDisableTiming) == cudaSuccess;
    }
    if (!ok) { set_error("projection: device allocation failed: %s", cudaGetErrorString(cudaGetLastError())); be_b200_projection_destroy(p); return nullptr; }
    return p;
}

// allocates at least 16 bytes but copies only `bytes`; the copy is complete (not merely staged) on return
static bool upload(void** dst, const void* src, size_t bytes) {
    if (cudaMalloc(dst, bytes ? bytes : 16) != cudaSuccess) return false;
    if (src && bytes && cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return false;
    return cudaDeviceSynchronize() == cudaSuccess;     // pageable H2D may return before its DMA lands
}

extern "C" be_b200_projection* be_b200_projection_create_csr(const float* weights, int homo,
                                                             const int32_t* indices,
                                                             const int32_t* indptr, int64_t m,
                                                             int64_t k) {
    if (!weights || !indptr || m < 0 || k < 0) { set_error("projection_create_csr: bad arguments"); return nullptr; }
    be_b200_projection* p = proj_new(0, m, k);
    if (!p) return nullptr;
    p->homo = homo;
    p->nse = indptr[m];
    bool ok = upload(&p->d_weights, weights, (size_t)(homo ? 1 : p->nse) * sizeof(float));
    ok = ok && upload((void**)&p->d_indices, indices, (size_t)p->nse * 4);
    ok = ok && upload((void**)&p->d_indptr, indptr, (size_t)(m + 1) * 4);
    if (!ok) { set_error("projection_create_csr: upload failed"); be_b200_projection_destroy(p); return nullptr; }
    return p;
}

extern "C" be_b200_projection* be_b200_projection_create_fcn(const float* weights, int homo,
                                                             const int32_t* indices, int64_t n_rows,
                                                             int64_t n_conn, int64_t n_post) {
    if (!weights || n_rows < 0 || n_conn < 0 || n_post < 0) { set_error("projection_create_fcn: bad arguments"); return nullptr; }
    be_b200_projection* p = proj_new(1, n_rows, n_post);
    if (!p) return nullptr;
    p->homo = homo; p->n_conn = n_conn; p->nse = n_rows * n_conn;
    bool ok = upload(&p->d_weights, weights, (size_t)(homo ? 1 : p->nse) * sizeof(float));
    ok = ok && upload((void**)&p->d_indices, indices, (size_t)p->nse * 4);
    if (!ok) { set_error("projection_create_fcn: upload failed"); be_b200_projection_destroy(p); return nullptr; }
    return p;
}

extern "C" be_b200_projection* be_b200_projection_adopt_fcn(const float* d_weights, int homo,
                                                            const int32_t* d_indices,
                                                            int64_t n_rows, int64_t n_conn,
                                                            int64_t n_post) {
    if (!d_weights || n_rows < 0 || n_conn < 0 || n_post < 0) { set_error("projection_adopt_fcn: bad arguments"); return nullptr; }
    be_b200_projection* p = proj_new(1, n_rows, n_post);
    if (!p) return nullptr;
    p->homo = homo; p->n_conn = n_conn; p->nse = n_rows * n_conn; p->borrowed = 1;
    p->d_weights = (void*)d_weights;
    p->d_indices = (int32_t*)d_indices;
    return p;
}

extern "C" be_b200_projection* be_b200_projection_create_jitc(int jit_kind, float p0, float p1,
                                                              int32_t clen, int32_t seed,
                                                              int64_t n_pre, int64_t n_post,
                                                              int64_t chunk_size, int corder) {
    if (n_pre < 0 || n_post < 0 || chunk_size < 1) { set_error("projection_create_jitc: bad arguments"); return nullptr; }
    be_b200_projection* p = proj_new(2, n_pre, n_post);
    if (!p) return nullptr;
    p->jit_kind = jit_kind; p->corder = corder; p->chunk_size = chunk_size;
    float hp[2] = {p0, p1};
    int32_t hc[2] = {clen, seed};
    bool ok = upload((void**)&p->d_p, hp, sizeof(hp)) && upload((void**)&p->d_clen_seed, hc, sizeof(hc));
    if (!ok) { set_error("projection_create_jitc: upload failed"); be_b200_projection_destroy(p); return nullptr; }
    return p;
}

static int proj_compute(be_b200_projection* p, const uint8_t* d_events, float* d_out, cudaStream_t s) {
    if (p->kind == 0)
        return be_b200_binary_csrmv(p->d_weights, BE_B200_F32, p->homo, p->d_indices, p->d_indptr, 0, nullptr, 0,
                                    d_events, nullptr, BE_B200_SPK_BOOL, p->n_pre, p->n_post, p->nse, 1,
                                    d_out, p->d_ws, p->ws_bytes, s);
    if (p->kind == 1)
        return be_b200_binary_fcnmv(p->d_weights, BE_B200_F32, p->homo, p->d_indices, p->n_pre, p->n_conn,
                                    d_events, nullptr, BE_B200_SPK_BOOL, p->n_pre, p->n_post, 1, d_out,
                                    p->d_ws, p->ws_bytes, s);
    return be_b200_binary_jitmv(p->jit_kind, p->d_p, p->d_p + 1, BE_B200_F32, p->d_clen_seed,
                                p->d_clen_seed + 1, d_events, nullptr, BE_B200_SPK_BOOL, p->n_pre,
                                p->n_post, p->chunk_size, p->corder, 0, -1, 0, -1, d_out, p->d_ws,
                                p->ws_bytes, s);
}

// Asynchronous variant: returns once the call is enqueued.  Host buffers must be pinned and stay
// untouched until be_b200_projection_wait() (or the kSlots-th next submit) returns; at most
// kSlots = 4 calls are in flight, the next submit first waits for the oldest one.
extern "C" int be_b200_projection_submit_host(be_b200_projection* p, const uint8_t* events_host,
                                              float* out_host) { BE_NVTX_RANGE("be_b200_projection_submit_host");
    BE_REQUIRE(p && events_host && out_host, "projection_submit_host: NULL pointer");
    const int i = p->slot;
    if (p->inflight[i]) { BE_CUDA(cudaEventSynchronize(p->ev_out[i])); p->inflight[i] = 0; }
    if (p->n_pre > 0) BE_CUDA(cudaMemcpyAsync(p->d_ev2[i], events_host, (size_t)p->n_pre, cudaMemcpyHostToDevice, p->s_in));
    BE_CUDA(cudaEventRecord(p->ev_in[i], p->s_in));
    BE_CUDA(cudaStreamWaitEvent(p->stream, p->ev_in[i], 0));
    int rc = proj_compute(p, p->d_ev2[i], p->d_out2[i], p->stream);
    if (rc) return rc;
    BE_CUDA(cudaEventRecord(p->ev_done[i], p->stream));
    BE_CUDA(cudaStreamWaitEvent(p->s_out, p->ev_done[i], 0));
    if (p->n_post > 0) BE_CUDA(cudaMemcpyAsync(out_host, p->d_out2[i], (size_t)p->n_post * sizeof(float), cudaMemcpyDeviceToHost, p->s_out));
    BE_CUDA(cudaEventRecord(p->ev_out[i], p->s_out));
    p->inflight[i] = 1;
    p->slot = (p->slot + 1) % kSlots;
    return BE_B200_OK;
}

// Blocks until every submitted call has delivered its result to host memory.
extern "C" int be_b200_projection_wait(be_b200_projection* p) { BE_NVTX_RANGE("be_b200_projection_wait");
    BE_REQUIRE(p, "projection_wait: NULL pointer");
    for (int i = 0; i < kSlots; ++i)
        if (p->inflight[i]) { BE_CUDA(cudaEventSynchronize(p->ev_out[i])); p->inflight[i] = 0; }
    return BE_B200_OK;
}

extern "C" int be_b200_projection_apply_host(be_b200_projection* p, const uint8_t* events_host,
                                             float* out_host) { BE_NVTX_RANGE("be_b200_projection_apply_host");
    BE_REQUIRE(p && events_host && out_host, "projection_apply_host: NULL pointer");
    cudaStream_t s = p->stream;
    if (p->n_pre > 0) BE_CUDA(cudaMemcpyAsync(p->d_events, events_host, (size_t)p->n_pre, cudaMemcpyHostToDevice, s));
    int rc = proj_compute(p, p->d_events, p->d_out, s);
    if (rc) return rc;
    if (p->n_post > 0) BE_CUDA(cudaMemcpyAsync(out_host, p->d_out, (size_t)p->n_post * sizeof(float), cudaMemcpyDeviceToHost, s));
    BE_CUDA(cudaStreamSynchronize(s));
    return BE_B200_OK;
}

extern "C" void be_b200_projection_destroy(be_b200_projection* p) {
    if (!p) return;
    if (!p->borrowed) { cudaFree(p->d_weights); cudaFree(p->d_indices); cudaFree(p->d_indptr); }
    cudaFree(p->d_p);
    cudaFree(p->d_clen_seed); cudaFree(p->d_events); cudaFree(p->d_out); cudaFree(p->d_ws);
    for (int i = 0; i < kSlots; ++i) {
        cudaFree(p->d_ev2[i]); cudaFree(p->d_out2[i]);
        if (p->ev_in[i]) cudaEventDestroy(p->ev_in[i]);
        if (p->ev_done[i]) cudaEventDestroy(p->ev_done[i]);
        if (p->ev_out[i]) cudaEventDestroy(p->ev_out[i]);
    }
    if (p->s_in) cudaStreamDestroy(p->s_in);
    if (p->s_out) cudaStreamDestroy(p->s_out);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}
