// Latency microbenchmarks behind the NJ cluster kernel's design (run: nvcc -arch=sm_100a -O3 -o cmb cluster_microbench.cu && ./cmb)
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ double warp_tree_sum(double v) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) v += __shfl_down_sync(0xffffffffu, v, s);
    return v;
}

template <int CT>
__global__ void __launch_bounds__(CT, 1) k(const unsigned long long* __restrict__ chase_small, const unsigned long long* __restrict__ chase_big,
                                          double* scatter, size_t scatter_n, long long* out) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double sh[64];
    __shared__ unsigned long long shq[64];
    const int tid = threadIdx.x, rank = cluster.block_rank(), CS = cluster.num_blocks();
    if (tid < 64) { sh[tid] = tid; shq[tid] = (tid * 7 + 1) % 64; }
    __syncthreads();
    cluster.sync();
    long long t0, t1;
    // 1: back-to-back cluster.sync
    t0 = clock64();
    for (int i = 0; i < 64; i++) cluster.sync();
    t1 = clock64();
    if (rank == 0 && tid == 0) out[0] = (t1 - t0) / 64;
    // 2: __syncthreads
    t0 = clock64();
    for (int i = 0; i < 64; i++) __syncthreads();
    t1 = clock64();
    if (rank == 0 && tid == 0) out[1] = (t1 - t0) / 64;
    // 3: dependent global loads (L2-resident chain, 64 MB), one thread
    if (rank == 0 && tid == 0) {
        unsigned long long p = 0;
        for (int i = 0; i < 32; i++) p = __ldcg(&chase_small[p]);   // warm
        t0 = clock64();
        for (int i = 0; i < 64; i++) p = __ldcg(&chase_small[p]);
        t1 = clock64();
        out[2] = (t1 - t0) / 64; out[20] = (long long)p;
        p = 0;
        t0 = clock64();
        for (int i = 0; i < 64; i++) p = __ldcg(&chase_big[p]);      // 8 GB chain: DRAM + TLB misses
        t1 = clock64();
        out[3] = (t1 - t0) / 64; out[21] = (long long)p;
    }
    cluster.sync();
    // 4: DSMEM dependent reads from the next rank
    if (tid == 0) {
        unsigned long long* rq = cluster.map_shared_rank(shq, (rank + 1) % CS);
        unsigned long long p = 0;
        t0 = clock64();
        for (int i = 0; i < 64; i++) p = rq[p];
        t1 = clock64();
        if (rank == 0) { out[4] = (t1 - t0) / 64; out[22] = (long long)p; }
    }
    cluster.sync();
    // 5: each thread 2 scattered 8-byte stores, then cluster.sync (release has to drain them)
    {
        size_t g = (size_t)rank * CT + tid;
        t0 = clock64();
        for (int rep = 0; rep < 16; rep++) {
            scatter[((g * 2 + 0) * 30011 + rep * 977) % scatter_n] = 1.0;
            scatter[((g * 2 + 1) * 30011 + rep * 977) % scatter_n] = 2.0;
            cluster.sync();
        }
        t1 = clock64();
        if (rank == 0 && tid == 0) out[5] = (t1 - t0) / 16;
    }
    // 6: fp64 division latency, warp tree sum latency (one warp)
    if (rank == 0 && tid < 32) {
        double v = 1.0 + tid, d = 3.0 + tid;
        t0 = clock64();
        for (int i = 0; i < 64; i++) v = v / d + 1.0;
        t1 = clock64();
        if (tid == 0) out[6] = (t1 - t0) / 64;
        sh[tid] = v;
        t0 = clock64();
        for (int i = 0; i < 64; i++) v = warp_tree_sum(v) * 0.5 + tid;
        t1 = clock64();
        if (tid == 0) out[7] = (t1 - t0) / 64;
        sh[tid] += v;
        // 64-bit integer division
        long long a = 123456789012345ll + tid, b = 7 + tid;
        t0 = clock64();
        for (int i = 0; i < 64; i++) a = a / b + 99999999999ll;
        t1 = clock64();
        if (tid == 0) { out[8] = (t1 - t0) / 64; out[23] = a; }
    }
    cluster.sync();
    // 7: all threads: one global load each (coalesced, L2 hit) + cluster.sync  vs  8: fence + global atomic round trip
    {
        t0 = clock64();
        double acc = 0;
        for (int rep = 0; rep < 16; rep++) {
            acc += __ldcg(&scatter[((size_t)rep * 16384 + rank * CT + tid) % scatter_n]);
            cluster.sync();
        }
        t1 = clock64();
        if (rank == 0 && tid == 0) out[9] = (t1 - t0) / 16;
        if (acc == 12345.678) out[24] = 1;
    }
    if (rank == 0 && tid == 0) {
        t0 = clock64();
        unsigned long long p = 0;
        for (int i = 0; i < 32; i++) p += atomicAdd((unsigned long long*)&out[30], 1ull);
        t1 = clock64();
        out[10] = (t1 - t0) / 32; out[25] = (long long)p;
    }
    cluster.sync();
}

template <int CT>
void run(int CS, unsigned long long* cs, unsigned long long* cb, double* sc, size_t scn, long long* out) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS); cfg.blockDim = dim3(CT); cfg.dynamicSmemBytes = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaFuncSetAttribute(k<CT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaMemset(out, 0, 64 * 8);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k<CT>, (const unsigned long long*)cs, (const unsigned long long*)cb, sc, scn, out);
    cudaError_t e2 = cudaDeviceSynchronize();
    long long h[64];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("{\"cluster\": %d, \"threads\": %d, \"err\": \"%s/%s\", \"cluster_sync\": %lld, \"syncthreads\": %lld, \"ldcg_chain_L2\": %lld, \"ldcg_chain_8GB\": %lld, "
           "\"dsmem_chain\": %lld, \"scatter2_plus_sync\": %lld, \"f64_div\": %lld, \"warp_tree_sum\": %lld, \"i64_div\": %lld, \"ld_plus_sync\": %lld, \"atomic_rt\": %lld}\n",
           CS, CT, cudaGetErrorString(e), cudaGetErrorString(e2), h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9], h[10]);
}

int main() {
    size_t small_n = (64u << 20) / 8, big_n = (size_t)(8ull << 30) / 8;
    unsigned long long *cs, *cb; double* sc; long long* out;
    cudaMalloc(&cs, small_n * 8); cudaMalloc(&cb, big_n * 8); cudaMalloc(&out, 64 * 8);
    size_t scn = (size_t)(7ull << 30) / 8;
    cudaMalloc(&sc, scn * 8); cudaMemset(sc, 0, scn * 8);
    // pointer chains: p -> (p * A + B) mod n, written on the host for the small one, sparse for the big one
    {
        unsigned long long* h = (unsigned long long*)malloc(small_n * 8);
        for (size_t i = 0; i < small_n; i++) h[i] = (i * 1000003ull + 12345ull) % small_n;
        cudaMemcpy(cs, h, small_n * 8, cudaMemcpyHostToDevice); free(h);
        // big chain: only the visited entries are written
        cudaMemset(cb, 0, big_n * 8);
        unsigned long long p = 0;
        for (int i = 0; i < 200; i++) { unsigned long long nx = (p * 2654435761ull + 40503ull * (i + 1)) % big_n; cudaMemcpy(cb + p, &nx, 8, cudaMemcpyHostToDevice); p = nx; }
    }
    run<1024>(16, cs, cb, sc, scn, out);
    run<1024>(8, cs, cb, sc, scn, out);
    run<256>(16, cs, cb, sc, scn, out);
    run<1024>(1, cs, cb, sc, scn, out);
    return 0;
}
