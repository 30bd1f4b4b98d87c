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

#include "placement_dev.cuh"

namespace dipb {

// updateClosestNodes (:86-124) as a level-synchronous BFS run by one CTA.  Every slot is
// reached at most once (tree), so processing a level in parallel gives the serial result.
// The frontier is tiny (15 queue entries and 6.6 levels per tip at 30 000 tips) and every level is a chain of
// dependent L2 round trips, so the level is kept short: the queue lives in shared memory (global arrays only
// for entries beyond PL_QCAP), the up-to-three slots of a node are handled by three threads (thread k skips k
// links of the adjacency list) instead of one after the other, and the head of a node is prefetched when the
// node is queued.
constexpr int PL_QCAP = 1024;
struct PlQueue {
    int node[PL_QCAP], from[PL_QCAP];
    double dis[PL_QCAP];
};
__device__ void bfs_closest(const int* head, const int* nxt, const int* e, const double* len, double* cdis, int* cid,
                            int x, int* q_node, int* q_from, double* q_dis, unsigned int* s_tail, PlShared* prof,
                            int leaf_limit, int split_node_min) {
    __shared__ unsigned int lo, hi;
    __shared__ PlQueue q;
    if (threadIdx.x == 0) {
        q.node[0] = x; q.from[0] = -1; q.dis[0] = 0.0;   // intended seed
        lo = 0; hi = 1; *s_tail = 1;
    }
    __syncthreads();
    while (true) {
        const unsigned int l = lo, h = hi;
        if (l >= h) break;
        for (unsigned int w = threadIdx.x; w < 3u * (h - l); w += blockDim.x) {
            const unsigned int t = l + w / 3u;
            const int k = (int)(w % 3u);
            const int node = t < PL_QCAP ? q.node[t] : q_node[t], fb = t < PL_QCAP ? q.from[t] : q_from[t];
            const double d = t < PL_QCAP ? q.dis[t] : q_dis[t];
            // The slots of a node never change after it is created: a leaf (id < leaf_limit) has one, an inner node made
            // by split_edge (id >= split_node_min) has head, head - 2, head - 3 (c3, c1, c0); only nodes of a loaded
            // backbone need the linked list.
            int s = head[node];
            if (node < leaf_limit) { if (k > 0) continue; }
            else if (node >= split_node_min) s -= (k == 0 ? 0 : k + 1);
            else for (int j = 0; j < k && s != -1; j++) s = nxt[s];
            if (s == -1) continue;
            // one round trip for everything the slot needs: target node, length, the 5-entry list
            const int to = e[s];
            const double ls = len[s];
            double cd[KC5];
            int ci[KC5];
#pragma unroll
            for (int j = 0; j < KC5; j++) { cd[j] = cdis[s * KC5 + j]; ci[j] = cid[s * KC5 + j]; }
            if (to == fb) continue;
            // list_insert: before the first entry with dis > d
            int at = KC5;
#pragma unroll
            for (int j = KC5 - 1; j >= 0; j--) if (cd[j] > d) at = j;
            if (at < KC5) {
#pragma unroll
                for (int j = KC5 - 1; j >= 0; j--) {
                    if (j > at) { cdis[s * KC5 + j] = cd[j - 1]; cid[s * KC5 + j] = ci[j - 1]; }
                    else if (j == at) { cdis[s * KC5 + j] = d; cid[s * KC5 + j] = x; }
                }
                const unsigned int pos = atomicAdd(s_tail, 1u);
                const double nd = d + ls;
                if (pos < PL_QCAP) { q.node[pos] = to; q.from[pos] = node; q.dis[pos] = nd; }
                else { q_node[pos] = to; q_from[pos] = node; q_dis[pos] = nd; }
                asm volatile("prefetch.global.L1 [%0];" ::"l"(head + to));
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) { lo = h; hi = *s_tail; if (prof) prof->bfs_levels++; }
        __syncthreads();
    }
    if (prof && threadIdx.x == 0) prof->bfs_nodes += *s_tail;
    __syncthreads();
}

// Places tips [i0, i1).  dist row of tip i: rows + (i - row_base) * ld.
__global__ void __launch_bounds__(PL_THREADS)
place_batch_kernel(int* head, int* e, int* nxt, int* belong, double* len, int* cid, double* cdis, int* rev,
                   const double* __restrict__ rows, size_t ld, int row_base, int i0, int i1, int node_off, PlShared* ps,
                   PlCand* cta_best, int* q_node, int* q_from, double* q_dis, unsigned int gen0, int profile, int first_split_tip) {
    __shared__ PlCand sb[PL_THREADS / 32];
    __shared__ unsigned int s_tail;
    const int G = gridDim.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    unsigned int gen = gen0;   // the barrier counter keeps counting across launches of one placement run
    long long tm = clock64();
#define PL_MARK(k)                                                                  \
    do {                                                                            \
        if (profile && blockIdx.x == 0 && tid == 0) {                               \
            const long long now__ = clock64();                                      \
            ps->cyc[k] += (unsigned long long)(now__ - tm);                         \
            tm = now__;                                                             \
        }                                                                           \
    } while (0)
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
        PL_MARK(0);
        pl_grid_barrier(&ps->bar_counter, G, gen);
        PL_MARK(1);
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
            PL_MARK(4);
            if (w == 0) {
                PlCand b = sb[0];
                for (int k = 1; k < PL_THREADS / 32; k++)
                    if (sb[k].add < b.add || (sb[k].add == b.add && sb[k].slot < b.slot)) b = sb[k];
                if (!(b.add < 2.0)) { b.add = 2.0; b.frac = 0.0; b.slot = 0; }   // the (0,0,2) tuple at position 0 wins
                const int idx = ps->idx;
                split_edge_warp(head, nxt, e, len, cdis, cid, belong, rev, b.slot, b.frac, b.add, i, idx, node_off);
                if (lane == 0) ps->idx = idx + 4;
                __threadfence_block();
            }
            __syncthreads();
            PL_MARK(5);
            bfs_closest(head, nxt, e, len, cdis, cid, i, q_node, q_from, q_dis, &s_tail, profile ? ps : nullptr, node_off, first_split_tip + node_off - 1);
        }
        PL_MARK(2);
        pl_grid_barrier(&ps->bar_counter, G, gen);
        PL_MARK(3);
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
    bfs_closest(head, nxt, e, len, cdis, cid, 0, q_node, q_from, q_dis, &s_tail, nullptr, 0, 0x7fffffff);
    bfs_closest(head, nxt, e, len, cdis, cid, 1, q_node, q_from, q_dis, &s_tail, nullptr, 0, 0x7fffffff);
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
    for (int i = 0; i < B; i++) bfs_closest(head, nxt, e, len, cdis, cid, i, q_node, q_from, q_dis, &s_tail, nullptr, 0, 0x7fffffff);
}

static int tree_alloc_impl(dipb_ctx* c, int n, dipb_tree* t) {
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
    return 0;
}
int tree_alloc(dipb_ctx* c, int n, dipb_tree** out) {
    dipb_tree* t = new dipb_tree();
    t->ctx = c; t->n = n;
    ctx_retain(c);
    const int rc = tree_alloc_impl(c, n, t);
    if (rc) { dipb_tree_free(t); return rc; }
    *out = t;
    return 0;
}

int place_scratch_alloc(dipb_ctx* c, int n, PlaceScratch* s) {
    DIPB_CUDA(pool_alloc(c, (void**)&s->ps, sizeof(PlShared)));
    DIPB_CUDA(cudaMemsetAsync(s->ps, 0, sizeof(PlShared), c->stream));
    DIPB_CUDA(pool_alloc(c, (void**)&s->q_node, sizeof(int) * (2 * (size_t)n + 8)));
    DIPB_CUDA(pool_alloc(c, (void**)&s->q_from, sizeof(int) * (2 * (size_t)n + 8)));
    DIPB_CUDA(pool_alloc(c, (void**)&s->q_dis, sizeof(double) * (2 * (size_t)n + 8)));
    return 0;
}
void place_scratch_free(dipb_ctx* c, PlaceScratch* s) {
    pool_free(c, s->ps); pool_free(c, s->q_node); pool_free(c, s->q_from); pool_free(c, s->q_dis);
    *s = PlaceScratch();
}

// rows [r0, r1) x cols [0, r1) of the selected provider into buf (or the matrix itself)
int place_fetch_rows(const dipb_dist_source* src, int r0, int r1, double* buf, size_t ld, const double** rows, int* row_base,
                     size_t* ld_out) {
    if (src->matrix) { *rows = src->matrix->d; *row_base = 0; *ld_out = (size_t)src->matrix->n; return 0; }
    *rows = buf; *row_base = r0; *ld_out = ld;
    if (src->msa) return msa_block(src->msa, src->dist_type, r0, r1, r1, buf, ld);
    return dipb_mash_dist_block(src->mash, r0, r1, r1, buf, ld);
}

// places tips [first_tip, end) onto the tree held in t (arrays sized for n_alloc leaves, internal ids offset by n_alloc)
static int place_run(dipb_ctx* c, const dipb_dist_source* src, int n_alloc, int first_tip, int end, dipb_tree* t,
                     PlaceScratch* sc) {
    int batch = 512;
    double* buf = nullptr;
    const size_t ld = (size_t)end;
    if (!src->matrix) {
        // keep the row buffer around 256 MB
        size_t want = (size_t)batch * ld * sizeof(double);
        while (want > (1ull << 28) && batch > 128) { batch /= 2; want /= 2; }
        DIPB_CUDA(pool_alloc(c, (void**)&buf, (size_t)batch * ld * sizeof(double)));
    }
    int G = c->num_sms;
    int per_sm = 0;
    DIPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, place_batch_kernel, PL_THREADS, 0));
    if (per_sm < 1) { set_error("placement kernel does not fit"); return DIPB_E_CUDA; }
    PlCand* cb = nullptr;
    DIPB_CUDA(pool_alloc(c, (void**)&cb, sizeof(PlCand) * G));
    int rc = 0;
    unsigned int gen0 = 0;
    int profile = getenv("DIPB_PLACE_PROFILE") ? 1 : 0;
    DIPB_CUDA(cudaMemsetAsync(&sc->ps->bar_counter, 0, sizeof(unsigned int), c->stream));
    DIPB_CUDA(cudaMemsetAsync(sc->ps->cyc, 0, sizeof(unsigned long long) * 8, c->stream));
    for (int i0 = first_tip; i0 < end && !rc; i0 += batch) {
        int i1 = i0 + batch < end ? i0 + batch : end;
        const double* rows; int row_base; size_t ldr;
        rc = place_fetch_rows(src, i0, i1, buf, ld, &rows, &row_base, &ldr);
        if (rc) break;
        int node_off = n_alloc;
        void* args[] = {&t->head, &t->e, &t->nxt, &t->belong, &t->len, &t->cid, &t->cdis, &t->rev, &rows, &ldr, &row_base,
                        &i0, &i1, &node_off, &sc->ps, &cb, &sc->q_node, &sc->q_from, &sc->q_dis, &gen0, &profile, &first_tip};
        cudaError_t e = cudaLaunchCooperativeKernel((void*)place_batch_kernel, dim3(G), dim3(PL_THREADS), args, 0, c->stream);
        if (e != cudaSuccess) { set_error("placement: cooperative launch failed: %s", cudaGetErrorString(e)); rc = DIPB_E_CUDA; break; }
        c->launches++;
        gen0 += 2u * (unsigned int)(i1 - i0);
    }
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (!rc && e != cudaSuccess) { set_error("placement: %s", cudaGetErrorString(e)); rc = DIPB_E_CUDA; }
    if (!rc && profile) {
        PlShared hs;
        if (cudaMemcpy(&hs, sc->ps, sizeof(hs), cudaMemcpyDeviceToHost) == cudaSuccess) {
            const double tips = (double)(end - first_tip);
            fprintf(stderr, "[placement] tips %d..%d: cycles per tip (CTA 0): score %.0f, barrier 1 %.0f, argmin %.0f, split %.0f, BFS %.0f, barrier 2 %.0f; BFS levels %.1f, queue entries %.1f per tip\n",
                    first_tip, end, hs.cyc[0] / tips, hs.cyc[1] / tips, hs.cyc[4] / tips, hs.cyc[5] / tips, hs.cyc[2] / tips, hs.cyc[3] / tips, hs.bfs_levels / tips, hs.bfs_nodes / tips);
        }
    }
    pool_free(c, cb);
    if (buf) pool_free(c, buf);
    return rc;
}

int place_from_scratch(dipb_ctx* c, const dipb_dist_source* src, int n_alloc, int end, dipb_tree* t, PlaceScratch* sc) {
    // d(1,0)
    const double* d01 = nullptr;
    double* row1 = nullptr;
    int rc = 0;
    if (src->matrix) d01 = src->matrix->d + (size_t)src->matrix->n;   // row 1, column 0
    else {
        DIPB_CUDA(pool_alloc(c, (void**)&row1, sizeof(double) * 8));
        rc = src->msa ? msa_block(src->msa, src->dist_type, 1, 2, 1, row1, 8) : dipb_mash_dist_block(src->mash, 1, 2, 1, row1, 8);
        if (rc) { pool_free(c, row1); return rc; }
        d01 = row1;
    }
    place_first_two_kernel<<<1, PL_THREADS, 0, c->stream>>>(t->head, t->e, t->nxt, t->belong, t->len, t->cid, t->cdis, t->rev, d01, n_alloc, sc->ps, sc->q_node, sc->q_from, sc->q_dis);
    DIPB_KERNEL_CHECK(c);
    rc = place_run(c, src, n_alloc, 2, end, t, sc);
    if (row1) pool_free(c, row1);
    return rc;
}

int check_source(const dipb_dist_source* s, int n) {
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
    PlaceScratch sc;
    rc = place_scratch_alloc(c, n, &sc);
    if (!rc) rc = place_from_scratch(c, src, n, n, t, &sc);
    place_scratch_free(c, &sc);
    if (!rc) rc = timer_end(c, DIPB_T_PLACE);
    if (rc) { dipb_tree_free(t); return rc; }
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
    // rooted binary backbone: exactly 4B-4 directed slots in use (src/placement_close_k.cu:887)
    for (int q = 0; q < 4 * backbone - 4; q++)
        if (h_belong[q] < 0 || h_e[q] < 0) { set_error("dipb_place_add: backbone slot %d is unused; a rooted binary backbone has 4B-4 = %d slots", q, 4 * backbone - 4); return DIPB_E_ARG; }
    if ((size_t)(4 * backbone - 4) < 8 * (size_t)n && h_belong[4 * backbone - 4] >= 0) { set_error("dipb_place_add: backbone has more than 4B-4 slots (not a rooted binary tree)"); return DIPB_E_ARG; }
    rc = timer_begin(c);
    if (rc) return rc;
    dipb_tree* t = nullptr;
    rc = tree_alloc(c, n, &t);
    if (rc) return rc;
    size_t N = (size_t)n;
    PlaceScratch sc;
    auto body = [&]() -> int {
        DIPB_CUDA(cudaMemcpyAsync(t->head, h_head, 2 * N * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        DIPB_CUDA(cudaMemcpyAsync(t->e, h_e, 8 * N * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        DIPB_CUDA(cudaMemcpyAsync(t->nxt, h_nxt, 8 * N * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        DIPB_CUDA(cudaMemcpyAsync(t->belong, h_belong, 8 * N * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        DIPB_CUDA(cudaMemcpyAsync(t->len, h_len, 8 * N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        int r = place_scratch_alloc(c, n, &sc);
        if (r) return r;
        place_backbone_kernel<<<1, PL_THREADS, 0, c->stream>>>(t->head, t->e, t->nxt, t->belong, t->len, t->cid, t->cdis, t->rev, 4 * backbone - 4, backbone, sc.ps, sc.q_node, sc.q_from, sc.q_dis);
        DIPB_KERNEL_CHECK(c);
        return place_run(c, src, n, backbone, n, t, &sc);
    };
    rc = body();
    place_scratch_free(c, &sc);
    if (!rc) rc = timer_end(c, DIPB_T_PLACE);
    if (rc) { dipb_tree_free(t); return rc; }
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

int dipb_tree_device_arrays(dipb_tree* t, int32_t** head, int32_t** e, int32_t** nxt, int32_t** belong, double** len, int32_t** closest_id,
                            double** closest_dis) {
    if (!t) { set_error("dipb_tree_device_arrays: null tree"); return DIPB_E_ARG; }
    if (head) *head = t->head;
    if (e) *e = t->e;
    if (nxt) *nxt = t->nxt;
    if (belong) *belong = t->belong;
    if (len) *len = t->len;
    if (closest_id) *closest_id = t->cid;
    if (closest_dis) *closest_dis = t->cdis;
    return 0;
}

int dipb_tree_n(const dipb_tree* t) { return t ? t->n : 0; }

void dipb_tree_free(dipb_tree* t) {
    if (!t) return;
    cudaSetDevice(t->ctx->device);
    cudaFree(t->head); cudaFree(t->e); cudaFree(t->nxt); cudaFree(t->belong); cudaFree(t->rev);
    cudaFree(t->len); cudaFree(t->cid); cudaFree(t->cdis);
    ctx_release(t->ctx);
    delete t;
}

}  // extern "C"
