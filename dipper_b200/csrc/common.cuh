// Shared plumbing of libdipper_b200: error handling, the context, timers, PTX helpers.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <string>
#include <atomic>
#include <vector>
#include "../../include/dipper_b200.h"

namespace dipb {

void set_error(const char* fmt, ...);

#define DIPB_CUDA(call)                                                                     \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            dipb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return DIPB_E_CUDA;                                                             \
        }                                                                                   \
    } while (0)

#define DIPB_KERNEL_CHECK(ctx)                                                              \
    do {                                                                                    \
        (ctx)->launches++;                                                                  \
        cudaError_t e__ = cudaGetLastError();                                               \
        if (e__ != cudaSuccess) {                                                           \
            dipb::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return DIPB_E_CUDA;                                                             \
        }                                                                                   \
    } while (0)

constexpr int kNumSMsDefault = 148;

}  // namespace dipb

struct dipb_ctx {
    int device = 0;
    int num_sms = dipb::kNumSMsDefault;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double elapsed[DIPB_T_COUNT];
    uint64_t launches = 0;
    uint64_t nj_rows_scanned = 0, nj_bytes_scanned = 0, nj_iterations = 0;
    // Life time: dipb_init holds one reference, every child handle (msa, mash, matrix, tree, D&C state) one more.
    // dipb_destroy drops the creator's reference; the stream, the events and the struct go away with the LAST
    // reference, so children may be freed after dipb_destroy (in any order) without touching freed memory.
    std::atomic<int> refs{1};
    bool destroyed = false;               // dipb_destroy was called
    std::vector<int32_t> last_clusters;   // test hook storage of dipb_dc_cluster_ids (per context)
    void* stage = nullptr;                // H2D staging block of the aligned uploads (<= 1 GB: kept and reused, a cudaMalloc /
    size_t stage_bytes = 0;               // cudaFree pair per upload costs 10-100 ms on some hosts); freed with the context
    unsigned long long nj_occ_key = 0;    // nj_cluster launch: last (cluster size, threads, shared memory) asked of the
    int nj_occ_clusters = 0;              // occupancy calculator and its answer (the query costs ~1.5 ms per tree)
};

namespace dipb {

inline void ctx_retain(dipb_ctx* c) { c->refs.fetch_add(1, std::memory_order_relaxed); }
void ctx_release(dipb_ctx* c);   // capi.cu: tears the context down when the last reference goes

// RAII-free helpers: time a region on the context's stream with CUDA events.
inline int timer_begin(dipb_ctx* c) {
    DIPB_CUDA(cudaEventRecord(c->ev0, c->stream));
    return 0;
}
inline int timer_end(dipb_ctx* c, int what) {
    DIPB_CUDA(cudaEventRecord(c->ev1, c->stream));
    DIPB_CUDA(cudaEventSynchronize(c->ev1));
    float ms = 0;
    DIPB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->elapsed[what] = ms;
    return 0;
}

// ---- PTX helpers (sm_100a) -------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Large, short-lived device buffers (distance matrix, packed planes, tensor-core operands) come from the
// stream-ordered pool of the context's stream: a caller that builds one tree after another re-uses the pages
// instead of paying cudaMalloc / cudaFree of ~11 GB per tree (measured ~100 ms of the 30 000-tip end-to-end time).
inline cudaError_t pool_alloc(dipb_ctx* c, void** p, size_t bytes) { return cudaMallocAsync(p, bytes, c->stream); }
inline void pool_free(dipb_ctx* c, void* p) { if (p) cudaFreeAsync(p, c->stream); }

// stride-halving tree over 32 lanes; lane 0 holds a[0] of tree32() in the oracle
__device__ __forceinline__ double warp_tree_sum(double v) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) v += __shfl_down_sync(0xffffffffu, v, s);
    return v;
}

}  // namespace dipb

struct dipb_matrix {
    dipb_ctx* ctx = nullptr;
    int n = 0;
    double* d = nullptr;  // n*n row-major fp64, stride n
};
