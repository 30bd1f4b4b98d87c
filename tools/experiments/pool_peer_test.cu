// Does a stream-ordered pool that is opened to a peer (cudaMemPoolSetAccess) still serve large allocations?
#include <cstdio>
#include <cuda_runtime.h>
#include <thread>
int main() {
    int nd = 0; cudaGetDeviceCount(&nd); if (nd < 2) { printf("need 2 devices\n"); return 0; }
    cudaSetDevice(0); cudaDeviceEnablePeerAccess(1, 0);
    cudaMemPool_t pool; cudaDeviceGetDefaultMemPool(&pool, 1);
    unsigned long long keep = ~0ull; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    cudaMemAccessDesc desc = {}; desc.location.type = cudaMemLocationTypeDevice; desc.location.id = 0; desc.flags = cudaMemAccessFlagsProtReadWrite;
    printf("set access: %s\n", cudaGetErrorString(cudaMemPoolSetAccess(pool, &desc, 1)));
    auto run = [&]() {
        cudaSetDevice(1);
        cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
        size_t sizes[] = {1ull << 20, 1ull << 30, 2634022912ull, 2634022912ull, 7ull << 30, 200ull << 20};
        void* p[6];
        for (int i = 0; i < 6; i++) { cudaError_t e = cudaMallocAsync(&p[i], sizes[i], s); size_t f, t; cudaMemGetInfo(&f, &t); printf("alloc %zu MB: %s (free %zu MB)\n", sizes[i] >> 20, cudaGetErrorString(e), f >> 20); cudaGetLastError(); }
        cudaStreamSynchronize(s);
    };
    std::thread th(run); th.join();
    return 0;
}
