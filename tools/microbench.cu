// Integer-pipe microbenchmarks for the roofline of the distance kernel:
// issue rate of POPC, LOP3, IMAD and of the kernel's own per-word mix on this GPU.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
template <int OP>
__global__ void k(uint32_t* out, uint32_t seed, long long* cycles) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u;
    uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) { acc[i] += __popc(a[i] ^ acc[i]); }                   // 1 POPC + 1 LOP + 1 IADD
            if (OP == 1) { a[i] = (a[i] ^ acc[i]) | (a[(i + 1) & 7] & seed); acc[i] ^= a[i]; }   // LOP3 only
            if (OP == 2) { acc[i] = acc[i] * 0x01000193u + a[i]; }              // IMAD
            if (OP == 3) {                                                      // distance-kernel mix per word pair
                uint32_t b0 = a[(i + 1) & 7], b1 = a[(i + 2) & 7], bv = a[(i + 3) & 7];
                uint32_t u = ((a[i] ^ b0) | (acc[i] ^ b1));
                uint32_t vv = a[(i + 4) & 7] & bv;
                acc[i] += __popc(vv & ~u) + (__popc(vv) << 16);
            }
            if (OP == 4) { acc[i] += __popc(a[i]) ; a[i] += 0x9e3779b9u; }      // POPC with trivial companions
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i] + a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP>
void run(const char* name, double ops_per_inner) {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int blocks = p.multiProcessorCount * 2, threads = 1024;
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, sizeof(uint32_t) * blocks * threads);
    cudaMalloc(&cyc, 8);
    k<OP><<<blocks, threads>>>(out, 12345u, cyc);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, 12345u, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long hc; cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
    double inner = (double)ITERS * 8 * threads * 2;  // per SM (2 blocks of 1024 threads resident)
    printf("{\"bench\": \"%s\", \"ms\": %.4f, \"cycles\": %lld, \"inner_per_clk_per_sm\": %.2f, \"ops_per_clk_per_sm\": %.2f, \"eff_clock_mhz\": %.0f}\n",
           name, ms, hc, inner / hc, inner * ops_per_inner / hc, hc / (ms * 1e3));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("popc+lop+iadd chain", 1);
    run<4>("popc + 2 iadd", 1);
    run<1>("lop3 x2", 2);
    run<2>("imad", 1);
    run<3>("dist mix (4 lop3 + 2 popc + 2 add) per word pair", 1);
    return 0;
}
