// Dependent-chain latency of the warp-level primitives the NJ cluster kernel leans on (one warp, one CTA).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_ops_latency warp_ops_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CHAIN 256
template <int OP>
__global__ void k(unsigned int* out, long long* cyc, int nwarps_active) {
    __shared__ unsigned int sm[1024];
    __shared__ double smd[512];
    const int lane = threadIdx.x & 31;
    unsigned int v = threadIdx.x * 2654435761u + 12345u;
    double d = (double)threadIdx.x * 1.25 + 3.0;
    sm[threadIdx.x & 1023] = (v >> 7) & 1023u;
    smd[threadIdx.x & 511] = d;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < CHAIN; i++) {
        if (OP == 0) v = __reduce_min_sync(0xffffffffu, v) + lane + i;                 // redux.sync.min.u32, full mask
        if (OP == 1) v = __reduce_min_sync(0xffffu << (lane & 16), v) + lane + i;     // half-warp masks
        if (OP == 2) v = __ballot_sync(0xffffffffu, (v >> (i & 7)) & 1u) + lane;       // vote.ballot
        if (OP == 3) v = __shfl_xor_sync(0xffffffffu, v, 1 + (i & 15)) + 1u;           // shfl
        if (OP == 4) v = sm[v & 1023u];                                                // dependent LDS
        if (OP == 5) { d = __shfl_xor_sync(0xffffffffu, d, 16) + d; }                  // 64-bit shuffle + DADD
        if (OP == 6) { d = d / (1.0 + (double)(i + 3)); }                              // fp64 division
        if (OP == 7) { d = fma(d, 1.0000001, 0.5); }                                   // DFMA chain
        if (OP == 8) v = v * 3u + 7u;                                                  // IMAD chain
        if (OP == 9) v = min(v ^ 0x5bd1e995u, v + 77u);                                // LOP + IADD + VIMNMX
        if (OP == 10) { d = smd[((int)d) & 511]; }                                     // LDS.64 + F2I dependent
        if (OP == 11) { unsigned long long e = __double_as_longlong(d); unsigned int hi = (unsigned int)(e >> 32), lo = (unsigned int)e;
                        unsigned int mh = __reduce_min_sync(0xffffffffu, hi); unsigned int ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
                        d = __longlong_as_double(((unsigned long long)mh << 32) | ml) + 1.0; }   // warp_min_u64 pattern
        if (OP == 12) { __syncthreads(); v += 1; }                                     // CTA barrier with nwarps
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) *cyc = (t1 - t0);
    out[blockIdx.x * blockDim.x + threadIdx.x] = v + (unsigned int)d;
}
template <int OP>
void run(const char* name, int threads, unsigned int* out, long long* cyc) {
    for (int r = 0; r < 2; r++) { k<OP><<<1, threads>>>(out, cyc, threads / 32); cudaDeviceSynchronize(); }
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("{\"op\": \"%s\", \"threads\": %d, \"cycles_per_op\": %.1f}\n", name, threads, (double)h / CHAIN);
}
int main() {
    unsigned int* out; long long* cyc;
    cudaMalloc(&out, 1024 * 4); cudaMalloc(&cyc, 8);
    for (int threads : {32, 512}) {
        run<0>("redux.min.u32 full mask", threads, out, cyc);
        run<1>("redux.min.u32 half-warp masks", threads, out, cyc);
        run<2>("vote.ballot", threads, out, cyc);
        run<3>("shfl.bfly b32", threads, out, cyc);
        run<4>("dependent LDS.32", threads, out, cyc);
        run<5>("shfl 64-bit + DADD", threads, out, cyc);
        run<6>("fp64 division", threads, out, cyc);
        run<7>("DFMA", threads, out, cyc);
        run<8>("IMAD", threads, out, cyc);
        run<9>("LOP3 + IADD + VIMNMX", threads, out, cyc);
        run<10>("LDS.64 + F2I", threads, out, cyc);
        run<11>("warp_min_u64 (2 redux) + DADD", threads, out, cyc);
        run<12>("__syncthreads", threads, out, cyc);
    }
    return 0;
}
