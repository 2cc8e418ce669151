// microbench_segread.cu -- HBM read bandwidth as a function of the contiguous segment length per row.
// The dense event matmul reads W[k, m] fp32 in tiles of 128 columns: every CTA streams 512-byte row segments at a
// 64 KB stride.  This probe reads the same 1 GiB matrix with segments of 512 B ... 64 KB per row (same bytes in flight
// per CTA: 256 threads x 4 x 16 B = 16 KB per iteration, 3 CTAs/SM) and reports GB/s.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench_segread microbench_segread.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256, 3) segread(const float4* __restrict__ W, int64_t row_f4, int seg_f4, int rows_per_split,
                                                  float* __restrict__ sink) {
    // CTA (x = column segment, y = K split).  One iteration: R = 1024 / seg_f4 rows x seg_f4 float4 = 1024 float4 = 16 KB
    const int R = 1024 / seg_f4;                 // rows per iteration
    const int tr = threadIdx.x / (seg_f4 / 4 > 0 ? seg_f4 / 4 : 1);   // unused when seg_f4 < 4
    (void)tr;
    const int64_t col0 = (int64_t)blockIdx.x * seg_f4;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_split;
    float acc = 0.f;
    for (int r = 0; r < rows_per_split; r += R) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = u * 256 + threadIdx.x;          // element of the 1024-float4 iteration tile
            const int rr = e / seg_f4, cc = e % seg_f4;
            v[u] = __ldg(W + (r0 + r + rr) * row_f4 + col0 + cc);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (acc == 1234.5f) sink[0] = acc;
}

int main() {
    const int64_t K = 16384, M = 16384;
    float* W; float* sink;
    cudaMalloc(&W, K * M * 4); cudaMalloc(&sink, 4);
    cudaMemset(W, 0, K * M * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int segs[] = {128, 256, 512, 1024, 2048, 4096, 16384, 65536};
    for (int splitk : {6, 16}) {
        for (int seg : segs) {
            const int seg_f4 = seg / 16;
            if (1024 % seg_f4) continue;
            int rows_per_split = (int)(K / splitk);
            rows_per_split -= rows_per_split % (1024 / seg_f4 > 0 ? 1024 / seg_f4 : 1);
            if (seg_f4 > 1024) continue;
            dim3 grid((unsigned)(M * 4 / seg), splitk);
            for (int i = 0; i < 3; ++i) segread<<<grid, 256>>>((const float4*)W, M / 4, seg_f4, rows_per_split, sink);
            cudaEventRecord(e0);
            for (int i = 0; i < 10; ++i) segread<<<grid, 256>>>((const float4*)W, M / 4, seg_f4, rows_per_split, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
            const double bytes = (double)rows_per_split * splitk * M * 4;
            printf("split_k %2d  segment %6d B  grid %5u x %2d  %.3f ms  %.0f GB/s  (%s)\n", splitk, seg, grid.x, splitk, ms,
                   bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
