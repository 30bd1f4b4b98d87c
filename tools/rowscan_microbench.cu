// How fast can ONE 16-CTA cluster read R random rows of a 30000 x 30000 fp64 matrix, each CTA its share?
// Variants of the NJ scan's access pattern (nvcc -arch=sm_100a -O3 -o rsm rowscan_microbench.cu && ./rsm)
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// MODE 0: interleaved 32-column chunks, warp per (row, UC chunks), ld.cg 8 B per lane
// MODE 1: contiguous slice per CTA, warp per (row, UC*32 columns), ld.cg 8 B per lane
// MODE 2: contiguous slice per CTA, 16 B per lane
// MODE 3: contiguous slice per CTA, cp.async.bulk of the whole slice into shared memory (one thread issues), 4 stages
template <int MODE, int UC>
__global__ void __launch_bounds__(1024, 1) k(const double* __restrict__ D, int n, size_t ld, const int* __restrict__ rows, int R, int iters,
                                           double* out, long long* cyc) {
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = cluster.block_rank(), CS = cluster.num_blocks();
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar[4];
    double acc = 0.0;
    const int nch = (n + 31) / 32;
    const int lch = (nch - rank + CS - 1) / CS;
    const int slice = (n + CS - 1) / CS;           // contiguous columns per CTA (modes 1-3)
    const int c0 = rank * slice, c1 = min(n, c0 + slice);
    if (MODE == 3 && tid == 0) {
        for (int s = 0; s < 4; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster.sync();
    long long t0 = clock64();
    unsigned int phase_bits = 0;
    for (int it = 0; it < iters; it++) {
        const int* rr = rows + (size_t)it * R;
        if (MODE == 0) {
            const int parts = (lch + UC - 1) / UC;
            for (int un = w; un < R * parts; un += 32) {
                const int r = __ldcg(&rr[un / parts]);
                const int lw0 = (un % parts) * UC;
                const double* row = D + (size_t)r * ld;
                double v[UC];
#pragma unroll
                for (int q = 0; q < UC; q++) { const int j = ((lw0 + q) * CS + rank) * 32 + lane; v[q] = j < n ? __ldcg(&row[j]) : 0.0; }
#pragma unroll
                for (int q = 0; q < UC; q++) acc += v[q];
            }
        } else if (MODE == 1) {
            const int per = UC * 32;
            const int parts = (c1 - c0 + per - 1) / per;
            for (int un = w; un < R * parts; un += 32) {
                const int r = __ldcg(&rr[un / parts]);
                const int j0 = c0 + (un % parts) * per;
                const double* row = D + (size_t)r * ld;
                double v[UC];
#pragma unroll
                for (int q = 0; q < UC; q++) { const int j = j0 + q * 32 + lane; v[q] = j < c1 ? __ldcg(&row[j]) : 0.0; }
#pragma unroll
                for (int q = 0; q < UC; q++) acc += v[q];
            }
        } else if (MODE == 2) {
            const int per = UC * 64;
            const int parts = (c1 - c0 + per - 1) / per;
            for (int un = w; un < R * parts; un += 32) {
                const int r = __ldcg(&rr[un / parts]);
                const int j0 = c0 + (un % parts) * per;
                const double* row = D + (size_t)r * ld;
                double2 v[UC];
#pragma unroll
                for (int q = 0; q < UC; q++) {
                    const int j = j0 + q * 64 + 2 * lane;
                    v[q] = (j + 1 < c1) ? __ldcg(reinterpret_cast<const double2*>(&row[j])) : make_double2(0, 0);
                }
#pragma unroll
                for (int q = 0; q < UC; q++) acc += v[q].x + v[q].y;
            }
        } else {
            // 4-stage ring of whole slices; thread 0 issues, all threads consume
            const int bytes = (c1 - c0) * 8;               // multiple of 16 when n and slice are even
            const int stage_doubles = slice;
            double* ring = reinterpret_cast<double*>(smem);
            auto issue = [&](int idx) {
                const int s = idx & 3;
                const int r = __ldcg(&rr[idx]);
                const double* src = D + (size_t)r * ld + c0;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ring + (size_t)s * stage_doubles)),
                             "l"(src), "r"(bytes), "r"(smem_u32(&bar[s]))
                             : "memory");
            };
            if (tid == 0) for (int p = 0; p < 3 && p < R; p++) issue(p);
            for (int idx = 0; idx < R; idx++) {
                const int s = idx & 3;
                if (tid == 0 && idx + 3 < R) issue(idx + 3);
                const unsigned int ph = (phase_bits >> s) & 1u;
                unsigned int done;
                do {
                    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar[s])), "r"(ph) : "memory");
                } while (!done);
                phase_bits ^= 1u << s;
                const double* st = ring + (size_t)s * stage_doubles;
                for (int j = tid; j < c1 - c0; j += 1024) acc += st[j];
                __syncthreads();   // stage free before it is refilled
            }
        }
        cluster.sync();
    }
    long long t1 = clock64();
    out[(size_t)rank * 1024 + tid] = acc;
    if (rank == 0 && tid == 0) *cyc = (t1 - t0) / iters;
}

template <int MODE, int UC>
void run(const char* name, const double* D, int n, const int* rows, int R, int iters, double* out, long long* cyc) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16); cfg.blockDim = dim3(1024);
    cfg.dynamicSmemBytes = MODE == 3 ? (size_t)4 * ((n + 15) / 16) * 8 : 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaFuncSetAttribute(k<MODE, UC>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncSetAttribute(k<MODE, UC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
    cudaError_t e = cudaSuccess, e2 = cudaSuccess;
    for (int rep = 0; rep < 2; rep++) {
        e = cudaLaunchKernelEx(&cfg, k<MODE, UC>, D, n, (size_t)n, rows, R, iters, out, cyc);
        e2 = cudaDeviceSynchronize();
    }
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double bytes_per_sm = (double)R * n * 8 / 16;
    printf("{\"variant\": \"%s\", \"uc\": %d, \"n\": %d, \"rows\": %d, \"err\": \"%s/%s\", \"cycles_per_iteration\": %lld, \"bytes_per_clk_per_sm\": %.1f}\n", name, UC, n, R,
           cudaGetErrorString(e), cudaGetErrorString(e2), h, bytes_per_sm / (double)h);
}

int main() {
    const int iters = 100;
    for (int n : {30000, 3008}) {
        const int R = n == 30000 ? 34 : 340;   // 8.2 MB per iteration either way; the 72 MB matrix stays in L2
        double* D; int* rows; double* out; long long* cyc;
        cudaMalloc(&D, (size_t)n * n * 8); cudaMemset(D, 0, (size_t)n * n * 8);
        cudaMalloc(&rows, sizeof(int) * R * iters); cudaMalloc(&out, 16 * 1024 * 8); cudaMalloc(&cyc, 8);
        int* h = new int[R * iters];
        unsigned long long s = 88172645463325252ull;
        for (int i = 0; i < R * iters; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % n); }
        cudaMemcpy(rows, h, sizeof(int) * R * iters, cudaMemcpyHostToDevice);
        run<0, 8>("interleaved chunks, ld.cg f64", D, n, rows, R, iters, out, cyc);
        run<1, 8>("contiguous slice, ld.cg f64", D, n, rows, R, iters, out, cyc);
        if (n % 32 == 0) {   // 16-byte alignment of the slices
            run<2, 4>("contiguous slice, ld.cg f64x2", D, n, rows, R, iters, out, cyc);
        }
        cudaFree(D); cudaFree(rows); cudaFree(out); cudaFree(cyc); delete[] h;
    }
    return 0;
}
