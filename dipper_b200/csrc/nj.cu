// K4/K5: neighbor joining on a device-resident fp64 matrix (sm_100a).
//
// Replaces NJDeviceArrays::findNeighbourJoiningTree and its kernels calculateU,
// findMinDist (+ thrust::min_element) and updateDisMatrix
// (reference src/neighborJoining.cu:94-271).  Not a port: the reference crosses the
// host<->device boundary ~8 times per iteration; here the whole N-2 iteration loop
// is enqueued on one stream with all state (n, x, y, realID, tree arrays) on the
// device, and the host copies the finished tree back once.
//
// Semantics kept bit-for-bit:
//  * candidate value for ordered pair (i,j): (d[i][j] - U[i]/(n-2)) - U[j]/(n-2);
//  * ties resolved in the reference's scan order (rowblock(i), j mod 256, j, i);
//  * branch lengths, realID bookkeeping and the row/column merge as written there.
// U is summed in one fixed order (blocks of 1024, 32x32 stride-halving trees, blocks
// ascending) instead of the reference's atomicAdd order, so results are reproducible.
#include "common.cuh"
#include "nj.cuh"

namespace dipb {

struct NJState {
    int n;          // active size
    int x, y;       // pair chosen by the last search (x < y)
    double dxy;
    int next_id;
    unsigned int ticket_scan, ticket_upd;
    unsigned long long rows_scanned, iters;
};

struct Cand {
    double v;
    int i, j;
};

__device__ __forceinline__ int rowblock_of(int i, int n) {
    // src/neighborJoining.cu:124-127: 256 row blocks, the first n%256 own one extra row
    int sz = n / 256, rem = n % 256;
    long long split = (long long)(sz + 1) * rem;
    if (i < split) return i / (sz + 1);
    return rem + (int)((i - split) / sz);
}
__device__ __forceinline__ unsigned long long tie_key(int i, int j, int n) {
    return ((unsigned long long)rowblock_of(i, n) << 56) | ((unsigned long long)(j & 255) << 48) |
           ((unsigned long long)j << 24) | (unsigned long long)i;
}
// strict "a comes before b" in the reference's total order
__device__ __forceinline__ bool cand_before(const Cand& a, const Cand& b, int n) {
    if (a.v < b.v) return true;
    if (a.v > b.v) return false;
    if (a.v >= 10000.0) return false;  // two empty slots
    return tie_key(a.i, a.j, n) < tie_key(b.i, b.j, n);
}

// ---- canonical row sums (calculateU :94-115) --------------------------------
constexpr int RS_THREADS = 256;   // 8 warps: warp w reduces chunks w, w + 8, w + 16, w + 24 of every 1024-column block
__global__ void __launch_bounds__(RS_THREADS) nj_rowsum_kernel(const double* __restrict__ D, int n, size_t ld,
                                                               double* __restrict__ U, double* __restrict__ u) {
    // One CTA per row.  Canonical order (shared with the oracle and the cluster kernel's new-row sum): stride-halving tree over
    // each 32-column chunk, the same tree over the 32 chunk sums of a 1024-column block, blocks added in ascending order.
    // The 32 loads of RB blocks are issued before any is reduced, the block trees of a batch run in parallel warps (two CTA
    // barriers per RB blocks instead of two per block), and several small CTAs per SM overlap one row's loads with another's
    // reduction: 2.9 ms (1024 threads, block by block) -> 2.0 ms (batched) -> see DESIGN.md for this form, at 30 000 tips.
    constexpr int RW = RS_THREADS / 32, RK = 32 / RW, RB = 8;
    __shared__ double ws[RB][33];
    __shared__ double sb[RB];
    const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const double* const row = D + (size_t)i * ld;
    const int nblk = (n + 1023) >> 10;
    double acc = 0.0;
    for (int b0 = 0; b0 < nblk; b0 += RB) {
        const int nb = nblk - b0 < RB ? nblk - b0 : RB;
        double v[RB][RK];
#pragma unroll
        for (int q = 0; q < RB; q++)
#pragma unroll
            for (int k = 0; k < RK; k++) {
                const int j = (b0 + q) * 1024 + (w + RW * k) * 32 + lane;
                v[q][k] = (q < nb && j < n) ? __ldcs(&row[j]) : 0.0;
            }
#pragma unroll
        for (int q = 0; q < RB; q++)
#pragma unroll
            for (int k = 0; k < RK; k++) {
                const double t = warp_tree_sum(v[q][k]);
                if (lane == 0) ws[q][w + RW * k] = t;
            }
        __syncthreads();
        if (w < nb) {
            const double g = warp_tree_sum(ws[w][lane]);
            if (lane == 0) sb[w] = g;
        }
        __syncthreads();
        if (tid == 0)
            for (int q = 0; q < nb; q++) acc += sb[q];
    }
    if (tid == 0) {
        U[i] = acc;
        u[i] = acc / (double)(n - 2);
    }
}

// ---- exhaustive search (findMinDist :117-148 + min_element :214) ------------
__global__ void __launch_bounds__(256) nj_scan_kernel(const double* __restrict__ D, size_t ld,
                                                      const double* __restrict__ U, const double* __restrict__ u,
                                                      NJState* st, Cand* block_best, int* realID, int32_t* child0,
                                                      int32_t* child1, double* len0, double* len1, int n_total) {
    const int n = st->n;
    if (n <= 2) return;
    const int tid = threadIdx.x;
    Cand best{10000.0, 0, 0};
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const double ui = u[i];
        const double* row = D + (size_t)i * ld;
        for (int j = tid; j < n; j += 256) {
            double t = row[j] - ui - u[j];
            if (i != j && t <= best.v && t < 10000.0) {
                Cand c{t, i, j};
                if (cand_before(c, best, n)) best = c;
            }
        }
    }
    __shared__ Cand sb[256];
    sb[tid] = best;
    __syncthreads();
    for (int s = 128; s >= 1; s >>= 1) {
        if (tid < s && cand_before(sb[tid + s], sb[tid], n)) sb[tid] = sb[tid + s];
        __syncthreads();
    }
    __shared__ bool is_last;
    if (tid == 0) {
        block_best[blockIdx.x] = sb[0];
        __threadfence();
        unsigned int t = atomicAdd(&st->ticket_scan, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    Cand c{10000.0, 0, 0};
    const volatile Cand* vb = block_best;
    for (int b = tid; b < (int)gridDim.x; b += 256) {
        Cand o{vb[b].v, vb[b].i, vb[b].j};
        if (cand_before(o, c, n)) c = o;
    }
    sb[tid] = c;
    __syncthreads();
    for (int s = 128; s >= 1; s >>= 1) {
        if (tid < s && cand_before(sb[tid + s], sb[tid], n)) sb[tid] = sb[tid + s];
        __syncthreads();
    }
    if (tid == 0) {
        // host step of the reference, :219-237
        int x = sb[0].i, y = sb[0].j;
        if (x > y) { int t = x; x = y; y = t; }
        double dxy = D[(size_t)x * ld + y];
        double blX = (dxy + u[x] - u[y]) * 0.5;
        double blY = dxy - blX;
        if (blX < 0) { blY += blX; blX = 0; }
        if (blY < 0) { blX += blY; blY = 0; }
        int id = st->next_id;
        child0[id - n_total] = realID[x]; len0[id - n_total] = blX;
        child1[id - n_total] = realID[y]; len1[id - n_total] = blY;
        realID[x] = id; realID[y] = realID[n - 1];
        st->next_id = id + 1;
        st->x = x; st->y = y; st->dxy = dxy;
        st->ticket_scan = 0;
        st->rows_scanned += (unsigned long long)n;
        st->iters += 1;
    }
}

// ---- merge update (updateDisMatrix :161-194), deterministic U ---------------
__global__ void __launch_bounds__(1024) nj_update_kernel(double* __restrict__ D, size_t ld, double* __restrict__ U,
                                                         double* __restrict__ u, NJState* st,
                                                         double* __restrict__ partial) {
    const int n = st->n;
    if (n <= 2) return;
    const int x = st->x, y = st->y, last = n - 1;
    const double dxy = st->dxy;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int i = blockIdx.x * 1024 + tid;
    double slot = 0.0;
    if (i < last && i != x) {
        if (i != y) {
            double a = D[(size_t)x * ld + i], b = D[(size_t)y * ld + i];
            double val = (a + b - dxy) * 0.5;
            double far = D[(size_t)last * ld + i];
            U[i] += -a - b + val;
            D[(size_t)x * ld + i] = val;
            D[(size_t)i * ld + x] = val;
            D[(size_t)y * ld + i] = far;
            D[(size_t)i * ld + y] = far;
            slot = val;
        } else {
            // i == y < last: the old last row moves here (:184-192)
            double a = D[(size_t)x * ld + last], b = D[(size_t)y * ld + last];
            double val = (a + b - dxy) * 0.5;
            double uy = U[last];
            uy += -a - b + val;
            U[y] = uy;
            D[(size_t)x * ld + y] = val;
            D[(size_t)y * ld + x] = val;
            slot = val;
        }
    }
    __shared__ double ws[32];
    __shared__ bool is_last;
    double v = warp_tree_sum(slot);
    if (lane == 0) ws[w] = v;
    __syncthreads();
    if (w == 0) {
        double g = warp_tree_sum(ws[lane]);
        if (lane == 0) {
            partial[blockIdx.x] = g;
            __threadfence();
            unsigned int t = atomicAdd(&st->ticket_upd, 1u);
            is_last = (t == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    __shared__ double ux_s;
    if (tid == 0) {
        double acc = 0.0;
        int nb = (last + 1023) / 1024;
        for (int b = 0; b < nb; b++) acc += const_cast<volatile double*>(partial)[b];
        const_cast<volatile double*>(U)[x] = acc;
        ux_s = acc;
        st->n = last;
        st->ticket_upd = 0;
    }
    __syncthreads();
    // u[j] = U[j] / (n' - 2) for the next search
    const int nn = last;
    if (nn > 2) {
        const double den = (double)(nn - 2);
        for (int j = tid; j < nn; j += 1024) {
            double Uj = (j == x) ? ux_s : const_cast<volatile double*>(U)[j];
            u[j] = Uj / den;
        }
    }
}

__global__ void nj_finish_kernel(const double* D, size_t ld, const int* realID, int32_t* child0, int32_t* child1,
                                 double* len0, double* len1, int n_total) {
    // :245-249
    double d = D[1];
    child0[n_total - 2] = realID[0]; len0[n_total - 2] = d * 0.5;
    child1[n_total - 2] = realID[1]; len1[n_total - 2] = d * 0.5;
    (void)ld;
}

__global__ void nj_init_kernel(NJState* st, int* realID, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) realID[i] = i;
    if (i == 0) {
        st->n = n; st->x = 0; st->y = 0; st->dxy = 0; st->next_id = n;
        st->ticket_scan = 0; st->ticket_upd = 0; st->rows_scanned = 0; st->iters = 0;
    }
}

namespace {
struct NJBuffers {      // device temporaries of one dipb_nj call; freed on every path by nj_run
    double *U = nullptr, *u = nullptr, *partial = nullptr, *l0 = nullptr, *l1 = nullptr;
    int* realID = nullptr;
    int32_t *c0 = nullptr, *c1 = nullptr;
    NJState* st = nullptr;
    Cand* bb = nullptr;
};
}  // namespace

static int nj_run_impl(dipb_matrix* m, int algo, const bool auto_algo, NJBuffers& b, int32_t* child0, int32_t* child1, double* len0, double* len1) {
    dipb_ctx* c = m->ctx;
    const int n = m->n;
    const size_t ld = (size_t)n;
    const int scan_grid = c->num_sms * 4;
    const int upd_grid = (n + 1023) / 1024;
    DIPB_CUDA(pool_alloc(c, (void**)&b.U, sizeof(double) * n));
    DIPB_CUDA(pool_alloc(c, (void**)&b.u, sizeof(double) * n));
    DIPB_CUDA(pool_alloc(c, (void**)&b.partial, sizeof(double) * (upd_grid + 1)));
    DIPB_CUDA(pool_alloc(c, (void**)&b.l0, sizeof(double) * n));
    DIPB_CUDA(pool_alloc(c, (void**)&b.l1, sizeof(double) * n));
    DIPB_CUDA(pool_alloc(c, (void**)&b.c0, sizeof(int32_t) * n));
    DIPB_CUDA(pool_alloc(c, (void**)&b.c1, sizeof(int32_t) * n));
    DIPB_CUDA(pool_alloc(c, (void**)&b.realID, sizeof(int) * n));
    DIPB_CUDA(pool_alloc(c, (void**)&b.st, sizeof(NJState)));
    DIPB_CUDA(pool_alloc(c, (void**)&b.bb, sizeof(Cand) * scan_grid));
    double *U = b.U, *u = b.u, *partial = b.partial, *l0 = b.l0, *l1 = b.l1;
    int* realID = b.realID;
    int32_t *c0 = b.c0, *c1 = b.c1;
    NJState* st = b.st;
    Cand* bb = b.bb;
    int rc = timer_begin(c);
    if (rc) return rc;
    nj_init_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(st, realID, n);
    DIPB_KERNEL_CHECK(c);
    int done = 0;
    if (n > 2) {
        nj_rowsum_kernel<<<n, RS_THREADS, 0, c->stream>>>(m->d, n, ld, U, u);
        DIPB_KERNEL_CHECK(c);
        if (algo == DIPB_NJ_CLUSTER) {
            rc = nj_cluster_loop(m, U, u, realID, c0, c1, l0, l1);
            if (rc == DIPB_E_UNSUPPORTED && auto_algo) algo = DIPB_NJ_PRUNED;   // no cluster shape fits this device
            else if (rc) return rc;
            else done = 1;
        }
        if (algo == DIPB_NJ_PRUNED) {
            rc = nj_pruned_loop(m, U, u, partial, st, realID, c0, c1, l0, l1);
            if (rc) return rc;
            done = 1;
        }
        if (!done) {
            for (int it = 0; it < n - 2; it++) {
                nj_scan_kernel<<<scan_grid, 256, 0, c->stream>>>(m->d, ld, U, u, st, bb, realID, c0, c1, l0, l1, n);
                c->launches++;
                nj_update_kernel<<<upd_grid, 1024, 0, c->stream>>>(m->d, ld, U, u, st, partial);
                c->launches++;
            }
            DIPB_CUDA(cudaGetLastError());
        }
    }
    nj_finish_kernel<<<1, 1, 0, c->stream>>>(m->d, ld, realID, c0, c1, l0, l1, n);
    DIPB_KERNEL_CHECK(c);
    rc = timer_end(c, DIPB_T_NJ);
    if (rc) return rc;
    NJState hs;
    DIPB_CUDA(cudaMemcpyAsync(&hs, st, sizeof(hs), cudaMemcpyDeviceToHost, c->stream));
    DIPB_CUDA(cudaMemcpyAsync(child0, c0, sizeof(int32_t) * (n - 1), cudaMemcpyDeviceToHost, c->stream));
    DIPB_CUDA(cudaMemcpyAsync(child1, c1, sizeof(int32_t) * (n - 1), cudaMemcpyDeviceToHost, c->stream));
    DIPB_CUDA(cudaMemcpyAsync(len0, l0, sizeof(double) * (n - 1), cudaMemcpyDeviceToHost, c->stream));
    DIPB_CUDA(cudaMemcpyAsync(len1, l1, sizeof(double) * (n - 1), cudaMemcpyDeviceToHost, c->stream));
    DIPB_CUDA(cudaStreamSynchronize(c->stream));
    if (!done) {
        c->nj_rows_scanned = hs.rows_scanned;
        c->nj_iterations = hs.iters;
        c->nj_bytes_scanned = 0;
    }
    return 0;
}

int nj_run(dipb_matrix* m, int algo, int32_t* child0, int32_t* child1, double* len0, double* len1) {
    dipb_ctx* c = m->ctx;
    const int n = m->n;
    if (n < 2) { set_error("dipb_nj: need at least 2 sequences"); return DIPB_E_ARG; }
    const bool auto_algo = algo == DIPB_NJ_AUTO;
    if (auto_algo) algo = nj_cluster_fits(n) ? DIPB_NJ_CLUSTER : DIPB_NJ_PRUNED;
    if (algo == DIPB_NJ_CLUSTER && !nj_cluster_fits(n)) { set_error("dipb_nj: %d tips do not fit the cluster kernel's shared memory", n); return DIPB_E_ARG; }
    NJBuffers b;
    const int rc = nj_run_impl(m, algo, auto_algo, b, child0, child1, len0, len1);
    if (rc) cudaStreamSynchronize(c->stream);   // a failed run may still have work queued on these buffers
    pool_free(c, b.U); pool_free(c, b.u); pool_free(c, b.partial); pool_free(c, b.l0); pool_free(c, b.l1); pool_free(c, b.c0); pool_free(c, b.c1);
    pool_free(c, b.realID); pool_free(c, b.st); pool_free(c, b.bb);
    return rc;
}

}  // namespace dipb
