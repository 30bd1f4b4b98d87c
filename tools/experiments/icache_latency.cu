// Does a once-per-iteration code body that exceeds the SM's instruction caches slow down a LATENCY-bound kernel (few warps,
// dependent chains, the NJ cluster kernel's regime)?  Body: NB blocks of 256 dependent integer multiply-adds (4 KB of SASS
// each), executed once per outer iteration; with and without a branch between the blocks.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache_latency icache_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int NB, bool BR>
__global__ void __launch_bounds__(512, 1) k(unsigned int* out, int reps, long long* cyc, unsigned int sel) {
    unsigned int a = threadIdx.x + 1u;
    long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
#pragma unroll
        for (int b = 0; b < NB; b++) {
            if (BR && ((sel >> (b & 31)) & 1u)) { a ^= 0x9e3779b9u; }   // (never taken: sel = 0; keeps the blocks apart as basic blocks)
#pragma unroll
            for (int i = 0; i < 256; i++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(sel + 0x01000193u), "r"(sel + 77u));
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = (t1 - t0) / reps;
}
template <int NB, bool BR>
void run(int grid, int threads, unsigned int* out, long long* cyc) {
    for (int rep = 0; rep < 2; rep++) { k<NB, BR><<<grid, threads>>>(out, 50, cyc, 0u); cudaDeviceSynchronize(); }
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("{\"body_kb\": %d, \"branches\": %d, \"grid\": %d, \"threads\": %d, \"cycles_per_iteration\": %lld, \"cycles_per_instr\": %.2f}\n", NB * 4, (int)BR, grid, threads, h, (double)h / (NB * 256.0));
}
int main() {
    unsigned int* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
    for (int threads : {32, 512}) {
        run<2, false>(16, threads, out, cyc); run<6, false>(16, threads, out, cyc); run<10, false>(16, threads, out, cyc);
        run<16, false>(16, threads, out, cyc); run<24, false>(16, threads, out, cyc); run<32, false>(16, threads, out, cyc);
        run<24, true>(16, threads, out, cyc);
    }
    run<24, false>(112, 512, out, cyc);
    return 0;
}
