// Does a once-per-iteration code body larger than the instruction cache slow a low-occupancy kernel down?
// Body: NB blocks of 256 independent-ish FFMAs (4 KB of SASS each) executed once per outer iteration by all warps.
#include <cstdio>
#include <cuda_runtime.h>

template <int NB>
__global__ void __launch_bounds__(1024, 1) k(float* out, int reps, long long* cyc) {
    float a0 = threadIdx.x, a1 = 1.f, a2 = 2.f, a3 = 3.f;
    const float m = 1.0001f;
    long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
#pragma unroll
        for (int b = 0; b < NB * 64; b++) {      // 4 FFMA per step = 64 B; NB * 64 steps = NB * 4 KB
            a0 = a0 * m + a1; a1 = a1 * m + a2; a2 = a2 * m + a3; a3 = a3 * m + a0;
        }
        __syncthreads();
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = (t1 - t0) / reps;
}

template <int NB>
void run(int grid, float* out, long long* cyc) {
    for (int rep = 0; rep < 2; rep++) {
        k<NB><<<grid, 1024>>>(out, 200, cyc);
        cudaDeviceSynchronize();
    }
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("{\"body_kb\": %d, \"grid\": %d, \"cycles_per_iteration\": %lld, \"cycles_per_instr_per_warp\": %.3f}\n", NB * 4, grid, h, (double)h / (NB * 256.0 * 32.0));
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    for (int grid : {16, 148}) {
        run<2>(grid, out, cyc); run<4>(grid, out, cyc); run<6>(grid, out, cyc); run<7>(grid, out, cyc); run<8>(grid, out, cyc);
        run<9>(grid, out, cyc); run<10>(grid, out, cyc); run<12>(grid, out, cyc); run<16>(grid, out, cyc); run<24>(grid, out, cyc);
    }
    return 0;
}
