// What does the first global load after a cluster barrier cost?  (nvcc -arch=sm_100a -O3 -o cmb2 cluster_microbench2.cu)
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

template <int MODE>   // 0: cluster.sync, 1: __syncthreads only, 2: cluster.sync, all warps load, 3: relaxed arrive + wait
__global__ void __launch_bounds__(1024, 1) k(const double* __restrict__ buf, size_t page_doubles, long long* out) {
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, rank = cluster.block_rank();
    long long acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double sink = 0;
    for (int rep = 0; rep < 64; rep++) {
        if (MODE == 1) __syncthreads();
        else if (MODE == 3) {
            asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
            asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
        } else cluster.sync();
        if (tid == 0 || MODE == 2) {
            const double* p0 = buf + (size_t)(rep * 37 % 3000) * page_doubles + rank * 512 + (MODE == 2 ? tid : 0);   // a page not touched recently
            long long t0 = clock64();
            double a = __ldcg(p0);
            sink += a;
            long long t1 = clock64();            // (sink dependency forces the wait)
            if (a > 1e30) sink += 1;
            double b = __ldcg(p0 + 4096);        // same page, other line
            sink += b;
            if (b > 1e30) sink += 1;
            long long t2 = clock64();
            double c = __ldcg(p0 + page_doubles * 3001);   // other page
            sink += c;
            if (c > 1e30) sink += 1;
            long long t3 = clock64();
            // issue 8 independent loads (other rows of the same page), then wait
            double v[8];
#pragma unroll
            for (int q = 0; q < 8; q++) v[q] = __ldcg(p0 + 8192 + q * 512);
            long long t4 = clock64();
#pragma unroll
            for (int q = 0; q < 8; q++) sink += v[q];
            if (sink > 1e30) sink += 1;
            long long t5 = clock64();
            if (tid == 0) { acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2; acc[3] += t4 - t3; acc[4] += t5 - t4; }
        }
    }
    if (rank == 0 && tid == 0) { for (int q = 0; q < 5; q++) out[q] = acc[q] / 64; out[7] = (long long)sink; }
    cluster.sync();
}

template <int MODE>
void run(const char* name, const double* buf, size_t pd, long long* out) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16); cfg.blockDim = dim3(1024);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int rep = 0; rep < 2; rep++) {
        cudaMemset(out, 0, 64);
        cudaLaunchKernelEx(&cfg, k<MODE>, buf, pd, out);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[8];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        if (rep) printf("{\"mode\": \"%s\", \"err\": \"%s\", \"first_load\": %lld, \"same_page\": %lld, \"other_page\": %lld, \"issue8\": %lld, \"wait8\": %lld}\n",
                        name, cudaGetErrorString(e), h[0], h[1], h[2], h[3], h[4]);
    }
}

int main() {
    const size_t pd = (2u << 20) / 8;              // doubles per 2 MB page
    const size_t total = pd * 6100;                // 12.2 GB
    double* buf; long long* out;
    if (cudaMalloc(&buf, total * 8) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(buf, 0, total * 8);
    cudaMalloc(&out, 64);
    run<0>("cluster.sync, thread 0 loads", buf, pd, out);
    run<1>("__syncthreads, thread 0 loads", buf, pd, out);
    run<2>("cluster.sync, all threads load", buf, pd, out);
    run<3>("relaxed arrive + wait, thread 0 loads", buf, pd, out);
    return 0;
}
