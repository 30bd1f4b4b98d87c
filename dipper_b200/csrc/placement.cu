// K6/K7: k-closest phylogenetic placement (K = 5) and add-tips onto a backbone.
//
// Replaces KPlacementDeviceArrays::{findPlacementTree, initializeDeviceArrays, addQuery}
// and the kernels calculateBranchLength, updateTreeStructure, updateClosestNodes,
// buildInitialTree (reference src/placement_close_k.cu:70-990).  Same data structure
// (directed-edge slots head/e/nxt/belong/len + 5 closest leaves per slot) so the exported
// arrays and the Newick text are those of the reference; execution is different:
//  * distance rows come in batches from the tiled MSA / Mash kernels (or straight from a
//    device matrix) instead of one launch + device sync per tip;
//  * one persistent cooperative kernel places a whole batch: per tip all CTAs score the
//    4i-4 LIVE slots (the reference always scores 4N-4) and reduce the argmin, then one
//    CTA splits the edge and runs the closest-leaf update as a level-parallel BFS
//    (the reference uses two <<<1,1>>> kernels and thrust::min_element + D2H per tip);
//  * a reverse-slot table replaces the per-edge adjacency walk of :339-340.
#include <vector>
#include "common.cuh"
#include "msa.cuh"

struct dipb_mash;
extern "C" int dipb_mash_dist_block(dipb_mash* m, int r0, int r1, int ncols, double* d_out, size_t ld);

struct dipb_tree {
    dipb_ctx* ctx = nullptr;
    int n = 0;
    int *head = nullptr, *e = nullptr, *nxt = nullptr, *belong = nullptr, *cid = nullptr, *rev = nullptr;
    double *len = nullptr, *cdis = nullptr;
};

namespace dipb {

constexpr int KC5 = 5;
constexpr int PL_THREADS = 256;

struct PlCand {
    double add;
    double frac;
    int slot;
    int pad;
};

struct PlShared {
    unsigned int bar_counter;
    unsigned int q_tail;
    int idx;   // next free slot
    int pad;
};

__device__ __forceinline__ void pl_grid_barrier(unsigned int* counter, unsigned int nblocks, unsigned int& gen) {
    gen++;
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int target = gen * nblocks;
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int v;
        do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while ((int)(v - target) < 0);
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ void link_slot(int* head, int* e, int* nxt, int* belong, double* len, int slot, int from,
                                          int to, double l) {
    e[slot] = to; len[slot] = l; nxt[slot] = head[from]; head[from] = slot; belong[slot] = from;
}

// insert (d, x) into the 5-entry list of slot s before the first entry with dis > d; true if inserted
__device__ __forceinline__ bool list_insert(double* cdis, int* cid, int s, double d, int x) {
    for (int j = 0; j < KC5; j++) {
        if (cdis[s * KC5 + j] > d) {
            for (int k = KC5 - 1; k > j; k--) {
                cdis[s * KC5 + k] = cdis[s * KC5 + k - 1];
                cid[s * KC5 + k] = cid[s * KC5 + k - 1];
            }
            cdis[s * KC5 + j] = d;
            cid[s * KC5 + j] = x;
            return true;
        }
    }
    return false;
}

// updateClosestNodes (:86-124) as a level-synchronous BFS run by one CTA.  Every slot is
// reached at most once (tree), so processing a level in parallel gives the serial result.
__device__ void bfs_closest(const int* head, const int* nxt, const int* e, const double* len, double* cdis, int* cid,
                            int x, int* q_node, int* q_from, double* q_dis, unsigned int* s_tail) {
    __shared__ unsigned int lo, hi;
    if (threadIdx.x == 0) {
        q_node[0] = x; q_from[0] = -1; q_dis[0] = 0.0;   // intended seed
        lo = 0; hi = 1; *s_tail = 1;
    }
    __syncthreads();
    while (true) {
        const unsigned int l = lo, h = hi;
        if (l >= h) break;
        for (unsigned int t = l + threadIdx.x; t < h; t += blockDim.x) {
            const int node = q_node[t], fb = q_from[t];
            const double d = q_dis[t];
            for (int s = head[node]; s != -1; s = nxt[s]) {
                if (e[s] == fb) continue;
                if (list_insert(cdis, cid, s, d, x)) {
                    unsigned int pos = atomicAdd(s_tail, 1u);
                    q_node[pos] = e[s]; q_from[pos] = node; q_dis[pos] = d + len[s];
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) { lo = h; hi = *s_tail; }
        __syncthreads();
    }
    __syncthreads();
}

// updateTreeStructure (:446-528), one thread
__device__ void split_edge(int* head, int* nxt, int* e, double* len, double* cdis, int* cid, int* belong, int* rev,
                           int eid, double fracLen, double addLen, int placeId, int edgeCount, int node_off) {
    const int middle = placeId + node_off - 1, outside = placeId;
    const int x = belong[eid], y = e[eid];
    const double orig = len[eid];
    const int xe = eid, ye = rev[eid];
    e[xe] = middle; len[xe] = fracLen;
    e[ye] = middle; len[ye] -= fracLen;
    const int c0 = edgeCount, c1 = edgeCount + 1, c2 = edgeCount + 2, c3 = edgeCount + 3;
    link_slot(head, e, nxt, belong, len, c0, middle, x, fracLen);
    for (int k = 0; k < KC5; k++)
        if (cid[ye * KC5 + k] != -1) { cid[c0 * KC5 + k] = cid[ye * KC5 + k]; cdis[c0 * KC5 + k] = cdis[ye * KC5 + k] + orig - fracLen; }
    link_slot(head, e, nxt, belong, len, c1, middle, y, orig - fracLen);
    for (int k = 0; k < KC5; k++)
        if (cid[xe * KC5 + k] != -1) { cid[c1 * KC5 + k] = cid[xe * KC5 + k]; cdis[c1 * KC5 + k] = cdis[xe * KC5 + k] + fracLen; }
    link_slot(head, e, nxt, belong, len, c2, outside, middle, addLen);
    link_slot(head, e, nxt, belong, len, c3, middle, outside, addLen);
    const int src[2] = {c1, c0};
    for (int w = 0; w < 2; w++)
        for (int i = 0; i < KC5; i++) {
            if (cid[src[w] * KC5 + i] == -1) break;
            list_insert(cdis, cid, c3, cdis[src[w] * KC5 + i], cid[src[w] * KC5 + i]);
        }
    rev[xe] = c0; rev[c0] = xe; rev[ye] = c1; rev[c1] = ye; rev[c2] = c3; rev[c3] = c2;
}

// calculateBranchLength (:309-358) for one candidate slot
__device__ __forceinline__ void score_slot(const double* __restrict__ dis, const int* cid, const double* cdis,
                                           const double* len, const int* rev, int q, double& frac, double& add) {
    const int r = __ldcg(&rev[q]);
    double d1 = 0, d2 = 0;
#pragma unroll
    for (int k = 0; k < KC5; k++) {
        int id = __ldcg(&cid[q * KC5 + k]);
        if (id != -1) { double v = dis[id] - __ldcg(&cdis[q * KC5 + k]); if (v > d1) d1 = v; }
    }
#pragma unroll
    for (int k = 0; k < KC5; k++) {
        int id = __ldcg(&cid[r * KC5 + k]);
        if (id != -1) { double v = dis[id] - __ldcg(&cdis[r * KC5 + k]); if (v > d2) d2 = v; }
    }
    const double L = __ldcg(&len[q]);
    double a = (d1 + d2 - L) / 2;
    if (a < 0) a = 0;
    d1 -= a; d2 -= a;
    if (d1 < 0) d1 = 0;
    if (d2 < 0) d2 = 0;
    if (d1 > L) { a += d1 - L; d1 = L; }
    if (d2 > L) { a += d2 - L; d2 = L; }
    const double rest = L - d1 - d2;
    d1 += rest / 2;
    frac = d1; add = a;
}

// Places tips [i0, i1).  dist row of tip i: rows + (i - row_base) * ld.
__global__ void __launch_bounds__(PL_THREADS)
place_batch_kernel(int* head, int* e, int* nxt, int* belong, double* len, int* cid, double* cdis, int* rev,
                   const double* __restrict__ rows, size_t ld, int row_base, int i0, int i1, int node_off, PlShared* ps,
                   PlCand* cta_best, int* q_node, int* q_from, double* q_dis, unsigned int gen0) {
    __shared__ PlCand sb[PL_THREADS / 32];
    __shared__ unsigned int s_tail;
    const int G = gridDim.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    unsigned int gen = gen0;   // the barrier counter keeps counting across launches of one placement run
    for (int i = i0; i < i1; i++) {
        const double* dis = rows + (size_t)(i - row_base) * ld;
        const int nslots = 4 * i - 4;
        // ---- all CTAs: score live slots, first minimum wins (thrust::min_element :807)
        double badd = 2.0, bfrac = 0.0;
        int bslot = 0;   // non-candidates emit (0,0,2); slot 0 is never a candidate
        for (int q = blockIdx.x * PL_THREADS + tid; q < nslots; q += G * PL_THREADS) {
            if (__ldcg(&belong[q]) > __ldcg(&e[q])) {
                double f, a;
                score_slot(dis, cid, cdis, len, rev, q, f, a);
                if (a < badd || (a == badd && q < bslot)) { badd = a; bfrac = f; bslot = q; }
            }
        }
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
            double oa = __shfl_xor_sync(0xffffffffu, badd, s), of = __shfl_xor_sync(0xffffffffu, bfrac, s);
            int os = __shfl_xor_sync(0xffffffffu, bslot, s);
            if (oa < badd || (oa == badd && os < bslot)) { badd = oa; bfrac = of; bslot = os; }
        }
        if (lane == 0) { sb[w].add = badd; sb[w].frac = bfrac; sb[w].slot = bslot; }
        __syncthreads();
        if (tid == 0) {
            PlCand b = sb[0];
            for (int k = 1; k < PL_THREADS / 32; k++)
                if (sb[k].add < b.add || (sb[k].add == b.add && sb[k].slot < b.slot)) b = sb[k];
            cta_best[blockIdx.x] = b;
        }
        pl_grid_barrier(&ps->bar_counter, G, gen);
        // ---- CTA 0: global argmin, split the edge, update the closest lists
        if (blockIdx.x == 0) {
            double a = 1e300, f = 0; int sl = 0x7fffffff;
            for (int b = tid; b < G; b += PL_THREADS) {
                double oa = __ldcg(&cta_best[b].add), of = __ldcg(&cta_best[b].frac);
                int os = __ldcg(&cta_best[b].slot);
                if (oa < a || (oa == a && os < sl)) { a = oa; f = of; sl = os; }
            }
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) {
                double oa = __shfl_xor_sync(0xffffffffu, a, s), of = __shfl_xor_sync(0xffffffffu, f, s);
                int os = __shfl_xor_sync(0xffffffffu, sl, s);
                if (oa < a || (oa == a && os < sl)) { a = oa; f = of; sl = os; }
            }
            if (lane == 0) { sb[w].add = a; sb[w].frac = f; sb[w].slot = sl; }
            __syncthreads();
            if (tid == 0) {
                PlCand b = sb[0];
                for (int k = 1; k < PL_THREADS / 32; k++)
                    if (sb[k].add < b.add || (sb[k].add == b.add && sb[k].slot < b.slot)) b = sb[k];
                if (!(b.add < 2.0)) { b.add = 2.0; b.frac = 0.0; b.slot = 0; }   // the (0,0,2) tuple at position 0 wins
                const int idx = ps->idx;
                split_edge(head, nxt, e, len, cdis, cid, belong, rev, b.slot, b.frac, b.add, i, idx, node_off);
                ps->idx = idx + 4;
                __threadfence_block();
            }
            __syncthreads();
            bfs_closest(head, nxt, e, len, cdis, cid, i, q_node, q_from, q_dis, &s_tail);
        }
        pl_grid_barrier(&ps->bar_counter, G, gen);
    }
}

__global__ void place_init_kernel(int* head, int* e, int* nxt, int* belong, double* len, int* cid, double* cdis, int* rev,
                                  int n) {
    // initialize (:266-289)
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < 2LL * n) head[g] = -1;
    if (g < 8LL * n) {
        e[g] = -1; nxt[g] = -1; belong[g] = -1; len[g] = 2; rev[g] = -1;
        for (int k = 0; k < KC5; k++) { cid[g * KC5 + k] = -1; cdis[g * KC5 + k] = 2; }
    }
}

// buildInitialTree (:530-554) + the two initial closest-leaf updates (:738-751)
__global__ void place_first_two_kernel(int* head, int* e, int* nxt, int* belong, double* len, int* cid, double* cdis,
                                       int* rev, const double* d01p, int node_off, PlShared* ps, int* q_node, int* q_from,
                                       double* q_dis) {
    __shared__ unsigned int s_tail;
    if (threadIdx.x == 0) {
        const double d = d01p[0];
        const int nv = node_off;
        link_slot(head, e, nxt, belong, len, 0, 0, nv, d / 2);
        link_slot(head, e, nxt, belong, len, 1, 1, nv, d / 2);
        link_slot(head, e, nxt, belong, len, 2, nv, 0, d / 2);
        link_slot(head, e, nxt, belong, len, 3, nv, 1, d / 2);
        rev[0] = 2; rev[2] = 0; rev[1] = 3; rev[3] = 1;
        ps->idx = 4;
    }
    __syncthreads();
    bfs_closest(head, nxt, e, len, cdis, cid, 0, q_node, q_from, q_dis, &s_tail);
    bfs_closest(head, nxt, e, len, cdis, cid, 1, q_node, q_from, q_dis, &s_tail);
}

// backbone: reverse-slot table + closest lists of leaves 0..B-1 in order (:241-260)
__global__ void place_backbone_kernel(int* head, int* e, int* nxt, int* belong, double* len, int* cid, double* cdis,
                                      int* rev, int nslots, int B, PlShared* ps, int* q_node, int* q_from, double* q_dis) {
    __shared__ unsigned int s_tail;
    for (int q = threadIdx.x; q < nslots; q += blockDim.x) {
        int x = belong[q], y = e[q], r = head[y];
        while (r != -1 && e[r] != x) r = nxt[r];
        rev[q] = r;
    }
    if (threadIdx.x == 0) ps->idx = nslots;
    __syncthreads();
    for (int i = 0; i < B; i++) bfs_closest(head, nxt, e, len, cdis, cid, i, q_node, q_from, q_dis, &s_tail);
}

static int tree_alloc(dipb_ctx* c, int n, dipb_tree** out) {
    dipb_tree* t = new dipb_tree();
    t->ctx = c; t->n = n;
    size_t N = (size_t)n;
    DIPB_CUDA(cudaMalloc(&t->head, 2 * N * sizeof(int)));
    DIPB_CUDA(cudaMalloc(&t->e, 8 * N * sizeof(int)));
    DIPB_CUDA(cudaMalloc(&t->nxt, 8 * N * sizeof(int)));
    DIPB_CUDA(cudaMalloc(&t->belong, 8 * N * sizeof(int)));
    DIPB_CUDA(cudaMalloc(&t->rev, 8 * N * sizeof(int)));
    DIPB_CUDA(cudaMalloc(&t->len, 8 * N * sizeof(double)));
    DIPB_CUDA(cudaMalloc(&t->cid, 8 * N * KC5 * sizeof(int)));
    DIPB_CUDA(cudaMalloc(&t->cdis, 8 * N * KC5 * sizeof(double)));
    long long total = 8LL * n;
    place_init_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(t->head, t->e, t->nxt, t->belong, t->len, t->cid, t->cdis, t->rev, n);
    DIPB_KERNEL_CHECK(c);
    *out = t;
    return 0;
}

struct RowSource {
    const dipb_dist_source* src;
    int n;
    double* buf = nullptr;
    size_t ld = 0;
    int batch = 0;
    // fill rows [r0, r1) (columns < r1) and return base pointer / row_base
    int fetch(int r0, int r1, const double** rows, int* row_base) {
        if (src->matrix) { *rows = src->matrix->d; *row_base = 0; ld = (size_t)src->matrix->n; return 0; }
        *rows = buf; *row_base = r0;
        if (src->msa) return msa_block(src->msa, src->dist_type, r0, r1, r1, buf, ld);
        return dipb_mash_dist_block(src->mash, r0, r1, r1, buf, ld);
    }
};

static int place_run(dipb_ctx* c, const dipb_dist_source* src, int n, int first_tip, dipb_tree* t, PlShared* ps,
                     int* q_node, int* q_from, double* q_dis) {
    RowSource rs{src, n};
    rs.ld = (size_t)n;
    rs.batch = 512;
    if (!src->matrix) {
        // keep the row buffer around 256 MB
        size_t want = (size_t)rs.batch * n * sizeof(double);
        while (want > (1ull << 28) && rs.batch > 128) { rs.batch /= 2; want /= 2; }
        DIPB_CUDA(cudaMalloc(&rs.buf, (size_t)rs.batch * n * sizeof(double)));
    }
    int G = c->num_sms;
    int per_sm = 0;
    DIPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, place_batch_kernel, PL_THREADS, 0));
    if (per_sm < 1) { set_error("placement kernel does not fit"); return DIPB_E_CUDA; }
    PlCand* cb = nullptr;
    DIPB_CUDA(cudaMalloc(&cb, sizeof(PlCand) * G));
    int rc = 0;
    unsigned int gen0 = 0;
    for (int i0 = first_tip; i0 < n && !rc; i0 += rs.batch) {
        int i1 = i0 + rs.batch < n ? i0 + rs.batch : n;
        const double* rows; int row_base;
        rc = rs.fetch(i0, i1, &rows, &row_base);
        if (rc) break;
        size_t ld = rs.ld;
        int node_off = n;
        void* args[] = {&t->head, &t->e, &t->nxt, &t->belong, &t->len, &t->cid, &t->cdis, &t->rev, &rows, &ld, &row_base,
                        &i0, &i1, &node_off, &ps, &cb, &q_node, &q_from, &q_dis, &gen0};
        cudaError_t e = cudaLaunchCooperativeKernel((void*)place_batch_kernel, dim3(G), dim3(PL_THREADS), args, 0, c->stream);
        if (e != cudaSuccess) { set_error("placement: cooperative launch failed: %s", cudaGetErrorString(e)); rc = DIPB_E_CUDA; break; }
        c->launches++;
        gen0 += 2u * (unsigned int)(i1 - i0);
    }
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (!rc && e != cudaSuccess) { set_error("placement: %s", cudaGetErrorString(e)); rc = DIPB_E_CUDA; }
    cudaFree(cb);
    if (rs.buf) cudaFree(rs.buf);
    return rc;
}

static int check_source(const dipb_dist_source* s, int n) {
    int cnt = (s->msa != nullptr) + (s->mash != nullptr) + (s->matrix != nullptr);
    if (cnt != 1) { set_error("placement: exactly one distance source must be set"); return DIPB_E_ARG; }
    if (s->msa && s->msa->n != n) { set_error("placement: msa holds %d sequences, n = %d", s->msa->n, n); return DIPB_E_ARG; }
    if (s->matrix && s->matrix->n != n) { set_error("placement: matrix is %d x %d, n = %d", s->matrix->n, s->matrix->n, n); return DIPB_E_ARG; }
    return 0;
}

}  // namespace dipb

using namespace dipb;

extern "C" {

int dipb_place_kclosest(dipb_ctx* c, const dipb_dist_source* src, int n, dipb_tree** out) {
    if (!c || !src || !out || n < 2) { set_error("dipb_place_kclosest: bad argument"); return DIPB_E_ARG; }
    int rc = check_source(src, n);
    if (rc) return rc;
    DIPB_CUDA(cudaSetDevice(c->device));
    rc = timer_begin(c);
    if (rc) return rc;
    dipb_tree* t = nullptr;
    rc = tree_alloc(c, n, &t);
    if (rc) return rc;
    PlShared* ps = nullptr;
    int *q_node = nullptr, *q_from = nullptr;
    double *q_dis = nullptr, *row1 = nullptr;
    DIPB_CUDA(cudaMalloc(&ps, sizeof(PlShared)));
    DIPB_CUDA(cudaMemsetAsync(ps, 0, sizeof(PlShared), c->stream));
    DIPB_CUDA(cudaMalloc(&q_node, sizeof(int) * (2 * (size_t)n + 8)));
    DIPB_CUDA(cudaMalloc(&q_from, sizeof(int) * (2 * (size_t)n + 8)));
    DIPB_CUDA(cudaMalloc(&q_dis, sizeof(double) * (2 * (size_t)n + 8)));
    // d(1,0)
    const double* d01 = nullptr;
    if (src->matrix) d01 = src->matrix->d + (size_t)n;   // row 1, column 0
    else {
        DIPB_CUDA(cudaMalloc(&row1, sizeof(double) * n));
        rc = src->msa ? msa_block(src->msa, src->dist_type, 1, 2, 1, row1, (size_t)n) : dipb_mash_dist_block(src->mash, 1, 2, 1, row1, (size_t)n);
        if (rc) return rc;
        d01 = row1;
    }
    place_first_two_kernel<<<1, PL_THREADS, 0, c->stream>>>(t->head, t->e, t->nxt, t->belong, t->len, t->cid, t->cdis, t->rev, d01, n, ps, q_node, q_from, q_dis);
    DIPB_KERNEL_CHECK(c);
    rc = place_run(c, src, n, 2, t, ps, q_node, q_from, q_dis);
    cudaFree(ps); cudaFree(q_node); cudaFree(q_from); cudaFree(q_dis);
    if (row1) cudaFree(row1);
    if (rc) { dipb_tree_free(t); return rc; }
    rc = timer_end(c, DIPB_T_PLACE);
    if (rc) return rc;
    *out = t;
    return 0;
}

int dipb_place_add(dipb_ctx* c, const dipb_dist_source* src, int n, int backbone, const int32_t* h_head, const int32_t* h_e,
                   const int32_t* h_nxt, const int32_t* h_belong, const double* h_len, dipb_tree** out) {
    if (!c || !src || !out || !h_head || !h_e || !h_nxt || !h_belong || !h_len || backbone < 2 || n < backbone) {
        set_error("dipb_place_add: bad argument");
        return DIPB_E_ARG;
    }
    int rc = check_source(src, n);
    if (rc) return rc;
    DIPB_CUDA(cudaSetDevice(c->device));
    rc = timer_begin(c);
    if (rc) return rc;
    dipb_tree* t = nullptr;
    rc = tree_alloc(c, n, &t);
    if (rc) return rc;
    size_t N = (size_t)n;
    DIPB_CUDA(cudaMemcpyAsync(t->head, h_head, 2 * N * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    DIPB_CUDA(cudaMemcpyAsync(t->e, h_e, 8 * N * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    DIPB_CUDA(cudaMemcpyAsync(t->nxt, h_nxt, 8 * N * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    DIPB_CUDA(cudaMemcpyAsync(t->belong, h_belong, 8 * N * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    DIPB_CUDA(cudaMemcpyAsync(t->len, h_len, 8 * N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    PlShared* ps = nullptr;
    int *q_node = nullptr, *q_from = nullptr;
    double* q_dis = nullptr;
    DIPB_CUDA(cudaMalloc(&ps, sizeof(PlShared)));
    DIPB_CUDA(cudaMemsetAsync(ps, 0, sizeof(PlShared), c->stream));
    DIPB_CUDA(cudaMalloc(&q_node, sizeof(int) * (2 * N + 8)));
    DIPB_CUDA(cudaMalloc(&q_from, sizeof(int) * (2 * N + 8)));
    DIPB_CUDA(cudaMalloc(&q_dis, sizeof(double) * (2 * N + 8)));
    // rooted binary backbone: 4B-4 slots (src/placement_close_k.cu:887)
    place_backbone_kernel<<<1, PL_THREADS, 0, c->stream>>>(t->head, t->e, t->nxt, t->belong, t->len, t->cid, t->cdis, t->rev, 4 * backbone - 4, backbone, ps, q_node, q_from, q_dis);
    DIPB_KERNEL_CHECK(c);
    rc = place_run(c, src, n, backbone, t, ps, q_node, q_from, q_dis);
    cudaFree(ps); cudaFree(q_node); cudaFree(q_from); cudaFree(q_dis);
    if (rc) { dipb_tree_free(t); return rc; }
    rc = timer_end(c, DIPB_T_PLACE);
    if (rc) return rc;
    *out = t;
    return 0;
}

int dipb_tree_export(dipb_tree* t, int32_t* head, int32_t* e, int32_t* nxt, int32_t* belong, double* len) {
    if (!t || !head || !e || !nxt || !belong || !len) { set_error("dipb_tree_export: bad argument"); return DIPB_E_ARG; }
    DIPB_CUDA(cudaSetDevice(t->ctx->device));
    size_t N = (size_t)t->n;
    DIPB_CUDA(cudaMemcpy(head, t->head, 2 * N * sizeof(int), cudaMemcpyDeviceToHost));
    DIPB_CUDA(cudaMemcpy(e, t->e, 8 * N * sizeof(int), cudaMemcpyDeviceToHost));
    DIPB_CUDA(cudaMemcpy(nxt, t->nxt, 8 * N * sizeof(int), cudaMemcpyDeviceToHost));
    DIPB_CUDA(cudaMemcpy(belong, t->belong, 8 * N * sizeof(int), cudaMemcpyDeviceToHost));
    DIPB_CUDA(cudaMemcpy(len, t->len, 8 * N * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int dipb_tree_export_closest(dipb_tree* t, int32_t* cid, double* cdis) {
    if (!t || !cid || !cdis) { set_error("dipb_tree_export_closest: bad argument"); return DIPB_E_ARG; }
    DIPB_CUDA(cudaSetDevice(t->ctx->device));
    size_t N = (size_t)t->n;
    DIPB_CUDA(cudaMemcpy(cid, t->cid, 8 * N * KC5 * sizeof(int), cudaMemcpyDeviceToHost));
    DIPB_CUDA(cudaMemcpy(cdis, t->cdis, 8 * N * KC5 * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int dipb_tree_n(const dipb_tree* t) { return t ? t->n : 0; }

void dipb_tree_free(dipb_tree* t) {
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    cudaFree(t->head); cudaFree(t->e); cudaFree(t->nxt); cudaFree(t->belong); cudaFree(t->rev);
    cudaFree(t->len); cudaFree(t->cid); cudaFree(t->cdis);
    delete t;
}

}  // extern "C"
