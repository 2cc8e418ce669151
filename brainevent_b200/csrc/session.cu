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
};

using namespace be;

static be_b200_projection* proj_new(int kind, int64_t n_pre, int64_t n_post) {
    be_b200_projection* p = new (std::nothrow) be_b200_projection();
    if (!p) { set_error("out of host memory"); return nullptr; }
    p->kind = kind; p->n_pre = n_pre; p->n_post = n_post;
    bool ok = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) == cudaSuccess;
    p->ws_bytes = be_b200_workspace_bytes(n_pre, 1);
    ok = ok && cudaMalloc(&p->d_ws, p->ws_bytes) == cudaSuccess;
    ok = ok && cudaMemset(p->d_ws, 0, p->ws_bytes) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&p->d_events, (size_t)std::max<int64_t>(n_pre, 16)) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&p->d_out, (size_t)std::max<int64_t>(n_post, 4) * sizeof(float)) == cudaSuccess;
    if (!ok) { set_error("projection: device allocation failed: %s", cudaGetErrorString(cudaGetLastError())); be_b200_projection_destroy(p); return nullptr; }
    return p;
}

static bool upload(void** dst, const void* src, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (cudaMalloc(dst, bytes) != cudaSuccess) return false;
    if (src && cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return false;
    return true;
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

extern "C" int be_b200_projection_apply_host(be_b200_projection* p, const uint8_t* events_host,
                                             float* out_host) {
    BE_REQUIRE(p && events_host && out_host, "projection_apply_host: NULL pointer");
    cudaStream_t s = p->stream;
    if (p->n_pre > 0) BE_CUDA(cudaMemcpyAsync(p->d_events, events_host, (size_t)p->n_pre, cudaMemcpyHostToDevice, s));
    int rc = BE_B200_OK;
    if (p->kind == 0)
        rc = be_b200_binary_csrmv(p->d_weights, BE_B200_F32, p->homo, p->d_indices, p->d_indptr, 0, nullptr, 0,
                                  p->d_events, nullptr, BE_B200_SPK_BOOL, p->n_pre, p->n_post, p->nse, 1,
                                  p->d_out, p->d_ws, p->ws_bytes, s);
    else if (p->kind == 1)
        rc = be_b200_binary_fcnmv(p->d_weights, BE_B200_F32, p->homo, p->d_indices, p->n_pre, p->n_conn,
                                  p->d_events, nullptr, BE_B200_SPK_BOOL, p->n_pre, p->n_post, 1, p->d_out,
                                  p->d_ws, p->ws_bytes, s);
    else
        rc = be_b200_binary_jitmv(p->jit_kind, p->d_p, p->d_p + 1, BE_B200_F32, p->d_clen_seed,
                                  p->d_clen_seed + 1, p->d_events, nullptr, BE_B200_SPK_BOOL, p->n_pre,
                                  p->n_post, p->chunk_size, p->corder, 0, -1, 0, -1, p->d_out, p->d_ws,
                                  p->ws_bytes, s);
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
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}
