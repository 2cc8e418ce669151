// microbench_fetch.cu -- DRAM bytes fetched per isolated 4-byte load, by load instruction.
// The hetero gather reads one fp32 weight per hit (1 M hits scattered over a 400 MB array) and ncu shows ~110 B of DRAM
// traffic per hit.  Each kernel below issues the same 2 M random, line-distinct 4-byte reads with a different instruction;
// run under `ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum` and divide by 2 M.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench_fetch microbench_fetch.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int kLoads = 1 << 21;

__device__ __forceinline__ uint64_t addr_of(int i, uint64_t n_lines) {
    // distinct 128-byte lines in a scrambled order, one 4-byte word inside each
    const uint64_t line = ((uint64_t)i * 2654435761ull) % n_lines;
    return line * 32 + ((uint64_t)i * 7 & 31);
}

template <int MODE>
__global__ void __launch_bounds__(256) fetch(const float* __restrict__ a, uint64_t n_lines, float* sink) {
    __shared__ float s_buf[256 * 4];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kLoads) return;
    const float* p = a + addr_of(i, n_lines);
    float v = 0.f;
    if (MODE == 0) v = *p;                                                            // ld.global (default .ca)
    if (MODE == 1) v = __ldg(p);                                                      // ld.global.nc
    if (MODE == 2) v = __ldcg(p);                                                     // ld.global.cg
    if (MODE == 3) v = __ldcs(p);                                                     // ld.global.cs
    if (MODE == 4) asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    if (MODE == 5) {                                                                  // explicit evict-first L2 policy
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    }
    if (MODE == 6) {                                                                  // cp.async.ca 4 B
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(s_buf + threadIdx.x);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(p) : "memory");
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
        v = s_buf[threadIdx.x];
    }
    if (MODE == 7) {                                                                  // cp.async.cg 16 B (aligned down)
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(s_buf + threadIdx.x * 4);
        const float* q = (const float*)((uintptr_t)p & ~(uintptr_t)15);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(q) : "memory");
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
        v = s_buf[threadIdx.x * 4];
    }
    if (MODE == 8) asm volatile("ld.global.cv.f32 %0, [%1];" : "=f"(v) : "l"(p));     // ld.global.cv (volatile)
    if (MODE == 9) asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
    if (v == 1234.5f) sink[0] = v;
}

int main(int argc, char** argv) {
    if (argc > 1) {       // optional: cudaLimitMaxL2FetchGranularity in bytes (32 / 64 / 128)
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[1]));
        size_t now = 0; cudaDeviceGetLimit(&now, cudaLimitMaxL2FetchGranularity);
        printf("cudaLimitMaxL2FetchGranularity <- %s: %s, now %zu\n", argv[1], cudaGetErrorString(e), now);
    }
    const uint64_t bytes = 400ull << 20, n_lines = bytes / 128;
    float *a, *sink;
    cudaMalloc(&a, bytes); cudaMalloc(&sink, 4);
    cudaMemset(a, 0, bytes);
    // flush L2 between kernels: touch a 256 MB scratch buffer
    char* scratch; cudaMalloc(&scratch, 256u << 20);
    const int blocks = kLoads / 256;
#define RUN(M) cudaMemset(scratch, M, 256u << 20); fetch<M><<<blocks, 256>>>(a, n_lines, sink); cudaDeviceSynchronize();
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
    printf("done: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
