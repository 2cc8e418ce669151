// Micro-benchmark: ceiling of scattered L2 reductions / stores / loads on B200.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_red tools/microbench_red.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t xs(uint32_t x) { x ^= x << 13; x ^= x >> 17; x ^= x << 5; return x; }

template <int MODE>
__global__ void k(float* out, uint32_t n_out, int iters, const int* __restrict__ idx, float* sink) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    float acc = 0.f;
    size_t base = (size_t)(blockIdx.x * blockDim.x + threadIdx.x);
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int i = 0; i < iters; ++i) {
        uint32_t j;
        if (MODE >= 10) j = (uint32_t)idx[base + (size_t)i * stride];   // index from memory (coalesced)
        else { s = xs(s); j = __umulhi(s, n_out); }
        int m = MODE % 10;
        if (m == 0) atomicAdd(out + j, 1.0f);                 // RED.ADD.F32
        else if (m == 1) out[j] = 1.0f;                       // scattered store
        else if (m == 2) acc += __ldg(out + j);               // scattered load
        else if (m == 3) atomicAdd((int*)out + j, 1);         // RED.ADD.S32
        else if (m == 4) { j &= ~1u; asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(out + j), "f"(1.0f), "f"(1.0f) : "memory"); }
        else if (m == 5) { j &= ~3u; asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + j), "f"(1.0f), "f"(1.0f), "f"(1.0f), "f"(1.0f) : "memory"); }
    }
    if (acc == 123.456f) *sink = acc;
}

template <int MODE>
void run(const char* name, size_t n_out, int blocks, int threads, int iters, const int* idx) {
    float *out, *sink;
    cudaMalloc(&out, n_out * 4); cudaMemset(out, 0, n_out * 4); cudaMalloc(&sink, 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<blocks, threads>>>(out, (uint32_t)n_out, iters, idx, sink);
    cudaEventRecord(a);
    k<MODE><<<blocks, threads>>>(out, (uint32_t)n_out, iters, idx, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = (double)blocks * threads * iters;
    printf("%-28s n_out=%9zu (%7.1f MB) grid=%5dx%4d: %8.3f ms  %8.2f Gop/s  err=%s\n", name, n_out, n_out * 4 / 1e6,
           blocks, threads, ms, ops / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(sink);
}

int main() {
    const int blocks = 148 * 8, threads = 256, iters = 256;
    size_t total = (size_t)blocks * threads * iters;
    int* idx; cudaMalloc(&idx, total * 4);
    // fill idx with random values < 1.25M using a kernel-free approach: host fill
    int* h = (int*)malloc(total * 4);
    uint32_t s = 99;
    for (size_t i = 0; i < total; ++i) { s = s * 1664525u + 1013904223u; h[i] = (int)(((uint64_t)s * 1250000u) >> 32); }
    cudaMemcpy(idx, h, total * 4, cudaMemcpyHostToDevice);
    for (size_t n_out : {(size_t)4000, (size_t)1250000, (size_t)10000000, (size_t)64000000}) {
        run<0>("red.f32 random", n_out, blocks, threads, iters, idx);
        run<3>("red.s32 random", n_out, blocks, threads, iters, idx);
        run<1>("st.f32 random", n_out, blocks, threads, iters, idx);
        run<2>("ld.f32 random", n_out, blocks, threads, iters, idx);
        run<4>("red.v2.f32 random", n_out, blocks, threads, iters, idx);
        run<5>("red.v4.f32 random", n_out, blocks, threads, iters, idx);
    }
    run<10>("red.f32 idx-from-mem 1.25M", 1250000, blocks, threads, iters, idx);
    run<0>("red.f32 random 4 CTA/SM", 1250000, 148 * 4, threads, iters * 2, idx);
    run<0>("red.f32 random 1024thr", 1250000, 148 * 2, 1024, iters, idx);
    return 0;
}
