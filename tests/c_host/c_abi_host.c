/* c_abi_host.c -- the drop-in boundary driven from plain C: no Python, no torch, only the CUDA runtime.
 *
 * A FixedNumPerPre projection (ref: _fcn/binary.py:43 binary_fcnmv; CUDA twin _fcn/binary_fcnmv.cu) is built on the
 * host, uploaded with cudaMemcpy, propagated through include/brainevent_b200.h in both orientations and with all four
 * event encodings, and compared with the CPU oracle (oracle/be_oracle.c, linked here as the CHECKER only -- this
 * program is test infrastructure, tests/test_c_host_gpu.py runs it).  Exit code 0 = every case matched.
 *
 * Build (tests/c_host/Makefile):  gcc c_abi_host.c -I../../include -I$CUDA/include -L../../brainevent_b200
 *        -lbrainevent_b200 -L../../oracle -lbe_oracle -L$CUDA/lib64 -lcudart -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "brainevent_b200.h"

void ora_fcnmv_f32(const float* w, int homo, const int32_t* indices, int64_t n_rows, int64_t n_conn, const uint8_t* act,
                   int64_t n_out, int transpose, float* out);

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define BE(x) do { int r_ = (x); if (r_ != BE_B200_OK) { fprintf(stderr, "%s -> %d: %s\n", #x, r_, be_b200_last_error()); return 3; } } while (0)

static uint32_t lcg(uint32_t* s) { *s = *s * 1664525u + 1013904223u; return *s >> 8; }

static int close_enough(const float* a, const float* b, int64_t n, const char* what) {
    double worst = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        const double d = fabs((double)a[i] - (double)b[i]), lim = 1e-5 * fabs((double)b[i]) + 2e-5;
        if (d > lim) { fprintf(stderr, "%s: element %lld: got %.9g, oracle %.9g\n", what, (long long)i, a[i], b[i]); return 0; }
        if (d > worst) worst = d;
    }
    printf("  %-44s ok (max abs diff %.3g over %lld outputs)\n", what, worst, (long long)n);
    return 1;
}

int main(void) {
    BE(be_b200_check_device());
    const int64_t n_pre = 20000, n_post = 30000, n_conn = 100;
    uint32_t seed = 12345u;
    int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * n_pre * n_conn);
    float* w = (float*)malloc(sizeof(float) * n_pre * n_conn);
    for (int64_t i = 0; i < n_pre * n_conn; ++i) { idx[i] = (int32_t)(lcg(&seed) % n_post); w[i] = (float)(lcg(&seed) % 2001) / 1000.0f - 1.0f; }
    /* scatter: events over the pre rows; gather: events over the post neurons */
    uint8_t* spk_pre = (uint8_t*)malloc(n_pre);
    uint8_t* spk_post = (uint8_t*)malloc(n_post);
    float* spk_pre_f = (float*)malloc(sizeof(float) * n_pre);
    for (int64_t i = 0; i < n_pre; ++i) { spk_pre[i] = (lcg(&seed) % 100) < 3; spk_pre_f[i] = spk_pre[i] ? 0.5f : 0.0f; }
    for (int64_t i = 0; i < n_post; ++i) spk_post[i] = (lcg(&seed) % 100) < 3;
    float* ref_sc = (float*)malloc(sizeof(float) * n_post);
    float* ref_ga = (float*)malloc(sizeof(float) * n_pre);
    float* ref_homo = (float*)malloc(sizeof(float) * n_post);
    const float w_homo = 0.25f;
    ora_fcnmv_f32(w, 0, idx, n_pre, n_conn, spk_pre, n_post, 1, ref_sc);
    ora_fcnmv_f32(w, 0, idx, n_pre, n_conn, spk_post, n_post, 0, ref_ga);
    ora_fcnmv_f32(&w_homo, 1, idx, n_pre, n_conn, spk_pre, n_post, 1, ref_homo);

    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int32_t* d_idx; float *d_w, *d_wh, *d_out, *d_spk_f; uint8_t *d_spk_pre, *d_spk_post; uint32_t* d_packed; int32_t* d_ids; void* d_ws;
    const size_t ws_bytes = be_b200_workspace_bytes(n_post > n_pre ? n_post : n_pre, 1);
    CK(cudaMalloc((void**)&d_idx, sizeof(int32_t) * n_pre * n_conn));
    CK(cudaMalloc((void**)&d_w, sizeof(float) * n_pre * n_conn));
    CK(cudaMalloc((void**)&d_wh, sizeof(float)));
    CK(cudaMalloc((void**)&d_out, sizeof(float) * (n_post > n_pre ? n_post : n_pre)));
    CK(cudaMalloc((void**)&d_spk_pre, n_pre));
    CK(cudaMalloc((void**)&d_spk_post, n_post));
    CK(cudaMalloc((void**)&d_spk_f, sizeof(float) * n_pre));
    CK(cudaMalloc((void**)&d_packed, sizeof(uint32_t) * ((n_pre + 31) / 32)));
    CK(cudaMalloc((void**)&d_ids, sizeof(int32_t) * (n_pre + 1)));
    CK(cudaMalloc(&d_ws, ws_bytes));
    CK(cudaMemset(d_ws, 0, ws_bytes));
    CK(cudaMemcpy(d_idx, idx, sizeof(int32_t) * n_pre * n_conn, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_w, w, sizeof(float) * n_pre * n_conn, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_wh, &w_homo, sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_spk_pre, spk_pre, n_pre, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_spk_post, spk_post, n_post, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_spk_f, spk_pre_f, sizeof(float) * n_pre, cudaMemcpyHostToDevice));

    float* got = (float*)malloc(sizeof(float) * (n_post > n_pre ? n_post : n_pre));
    int ok = 1;
    printf("C host over include/brainevent_b200.h (ABI %d), FixedNumPerPre %lld x %lld, %lld synapses/row\n",
           be_b200_abi_version(), (long long)n_pre, (long long)n_post, (long long)n_conn);
#define FETCH(n) do { CK(cudaStreamSynchronize(st)); CK(cudaMemcpy(got, d_out, sizeof(float) * (n), cudaMemcpyDeviceToHost)); } while (0)
    /* 1. bool events, scatter (event_array @ FixedNumPerPre) */
    BE(be_b200_binary_fcnmv(d_w, BE_B200_F32, 0, d_idx, n_pre, n_conn, d_spk_pre, NULL, BE_B200_SPK_BOOL, n_pre, n_post, 1, d_out, d_ws, ws_bytes, st));
    FETCH(n_post); ok &= close_enough(got, ref_sc, n_post, "scatter, bool events, hetero fp32");
    /* 2. float events (active when > 0) */
    BE(be_b200_binary_fcnmv(d_w, BE_B200_F32, 0, d_idx, n_pre, n_conn, d_spk_f, NULL, BE_B200_SPK_F32, n_pre, n_post, 1, d_out, d_ws, ws_bytes, st));
    FETCH(n_post); ok &= close_enough(got, ref_sc, n_post, "scatter, float events");
    /* 3. bit-packed and compact containers, produced by the library's own encoders */
    BE(be_b200_compact_1d(d_spk_pre, BE_B200_SPK_BOOL, n_pre, 0, d_ids, d_ids + n_pre, d_packed, d_ws, ws_bytes, st));
    BE(be_b200_binary_fcnmv(d_w, BE_B200_F32, 0, d_idx, n_pre, n_conn, d_packed, NULL, BE_B200_SPK_PACKED, n_pre, n_post, 1, d_out, d_ws, ws_bytes, st));
    FETCH(n_post); ok &= close_enough(got, ref_sc, n_post, "scatter, BitPackedBinary");
    BE(be_b200_binary_fcnmv(d_w, BE_B200_F32, 0, d_idx, n_pre, n_conn, d_packed, d_ids, BE_B200_SPK_COMPACT, n_pre, n_post, 1, d_out, d_ws, ws_bytes, st));
    FETCH(n_post); ok &= close_enough(got, ref_sc, n_post, "scatter, CompactBinary");
    /* 4. homogeneous weight */
    BE(be_b200_binary_fcnmv(d_wh, BE_B200_F32, 1, d_idx, n_pre, n_conn, d_spk_pre, NULL, BE_B200_SPK_BOOL, n_pre, n_post, 1, d_out, d_ws, ws_bytes, st));
    FETCH(n_post); ok &= close_enough(got, ref_homo, n_post, "scatter, homogeneous weight");
    /* 5. gather (FixedNumPerPre @ event_array): deterministic, run twice and compare bit for bit */
    BE(be_b200_binary_fcnmv(d_w, BE_B200_F32, 0, d_idx, n_pre, n_conn, d_spk_post, NULL, BE_B200_SPK_BOOL, n_post, n_pre, 0, d_out, d_ws, ws_bytes, st));
    FETCH(n_pre); ok &= close_enough(got, ref_ga, n_pre, "gather, bool events");
    float* again = (float*)malloc(sizeof(float) * n_pre);
    memcpy(again, got, sizeof(float) * n_pre);
    BE(be_b200_binary_fcnmv(d_w, BE_B200_F32, 0, d_idx, n_pre, n_conn, d_spk_post, NULL, BE_B200_SPK_BOOL, n_post, n_pre, 0, d_out, d_ws, ws_bytes, st));
    FETCH(n_pre);
    if (memcmp(again, got, sizeof(float) * n_pre) != 0) { fprintf(stderr, "gather is not bit-reproducible\n"); ok = 0; }
    else printf("  %-44s ok\n", "gather, second run identical bit for bit");
    /* 6. error behaviour: a NULL structure pointer is refused with a message, nothing is launched */
    if (be_b200_binary_fcnmv(d_w, BE_B200_F32, 0, NULL, n_pre, n_conn, d_spk_pre, NULL, BE_B200_SPK_BOOL, n_pre, n_post, 1, d_out, d_ws, ws_bytes, st) == BE_B200_OK) {
        fprintf(stderr, "NULL indices were accepted\n"); ok = 0;
    } else printf("  %-44s ok (\"%s\")\n", "NULL indices refused", be_b200_last_error());
    printf(ok ? "c_abi_host: all cases match the oracle\n" : "c_abi_host: MISMATCH\n");
    return ok ? 0 : 1;
}
