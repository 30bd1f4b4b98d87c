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
                            int leaf_limit, int split_node_min, const PlDirty* dirty = nullptr, const int* rev = nullptr) {
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
            // Tip p >= first_split_tip owns slots 4p-4 .. 4p-1 (c0: middle->x, c1: middle->y, c2: p->middle, c3: middle->p) and
            // made the inner node p + leaf_limit - 1, so its slots follow from the node number: no head[] round trip per
            // BFS level.  Only the first two leaves and the nodes of a loaded backbone go through the lists.
            int s;
            const int first_tip_ = split_node_min == 0x7fffffff ? 0x7fffffff : split_node_min - leaf_limit + 1;
            if (node < leaf_limit) {
                if (k > 0) continue;
                s = node >= first_tip_ ? 4 * node - 2 : head[node];
            } else if (node >= split_node_min) {
                const int p4 = 4 * (node - leaf_limit + 1);
                s = k == 0 ? p4 - 1 : (k == 1 ? p4 - 3 : p4 - 4);
            } else {
                s = head[node];
                for (int j = 0; j < k && s != -1; j++) s = nxt[s];
            }
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
                if (dirty) pl_mark_dirty(*dirty, node > to ? s : rev[s]);   // the edge's candidate slot is the one with belong > e
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

// ---- speculative batches (experiment, DIPB_PLACE_SPEC=1) --------------------------------------------------------------
// The per-tip kernel above synchronises the whole grid twice per tip (3.5 k + 2.2 k cycles of software barrier) around
// 6-8 k cycles of scoring, and then one CTA works alone for 20 k cycles.  Here a batch of T tips is scored at once
// against the tree AS IT IS WHEN THE BATCH STARTS (all SMs, slot data read once per four tips), and then ONE CTA places
// the tips in order without any grid-wide step: the score of a slot only changes when an insertion touches its edge
// (the split edge, the four new slots, the slots whose closest-leaf list the BFS updates: ~20 per tip), those slots are
// collected in a dirty list and re-scored exactly for every later tip of the batch; the best untouched slot of a tip is
// the one the batch scoring found -- unless that very slot has been touched, in which case the batch ends early and a
// new one starts at this tip.  Same winner, same tie-break (smallest slot among equal pendant lengths), same arrays.
constexpr int PSQ = 4;   // tips per CTA in the batch scoring
__global__ void __launch_bounds__(256)
place_spec_score_kernel(const int* __restrict__ e, const int* __restrict__ belong, const double* __restrict__ len,
                        const int* __restrict__ cid, const double* __restrict__ cdis, const int* __restrict__ rev, int nslots,
                        const double* __restrict__ rows, size_t ld, int row_base, int t0, int T, PlCand* __restrict__ part) {
    __shared__ PlCand sb[PSQ][8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int g = blockIdx.x, slice = blockIdx.y, nslices = gridDim.y;
    const double* dis[PSQ];
    double badd[PSQ];
    int bslot[PSQ];
#pragma unroll
    for (int j = 0; j < PSQ; j++) {
        const int t = t0 + min(g * PSQ + j, T - 1);
        dis[j] = rows + (size_t)(t - row_base) * ld;
        badd[j] = 2.0; bslot[j] = 0;
    }
    for (int q = slice * 256 + threadIdx.x; q < nslots; q += nslices * 256) {
        if (!(belong[q] > e[q])) continue;
        const int r = rev[q];
        int iq[KC5], ir[KC5];
        double dq[KC5], dr[KC5];
#pragma unroll
        for (int k = 0; k < KC5; k++) { iq[k] = cid[q * KC5 + k]; dq[k] = cdis[q * KC5 + k]; ir[k] = cid[r * KC5 + k]; dr[k] = cdis[r * KC5 + k]; }
        const double L = len[q];
#pragma unroll
        for (int j = 0; j < PSQ; j++) {
            double d1 = 0, d2 = 0;
#pragma unroll
            for (int k = 0; k < KC5; k++) {
                if (iq[k] != -1) { const double v = dis[j][iq[k]] - dq[k]; if (v > d1) d1 = v; }
                if (ir[k] != -1) { const double v = dis[j][ir[k]] - dr[k]; if (v > d2) d2 = v; }
            }
            double a = (d1 + d2 - L) / 2;
            if (a < 0) a = 0;
            d1 -= a; d2 -= a;
            if (d1 < 0) d1 = 0;
            if (d2 < 0) d2 = 0;
            if (d1 > L) a += d1 - L;
            if (d2 > L) a += d2 - L;
            if (a < badd[j] || (a == badd[j] && q < bslot[j])) { badd[j] = a; bslot[j] = q; }
        }
    }
#pragma unroll
    for (int j = 0; j < PSQ; j++) {
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
            const double oa = __shfl_xor_sync(0xffffffffu, badd[j], s);
            const int os = __shfl_xor_sync(0xffffffffu, bslot[j], s);
            if (oa < badd[j] || (oa == badd[j] && os < bslot[j])) { badd[j] = oa; bslot[j] = os; }
        }
        if (lane == 0) { sb[j][w].add = badd[j]; sb[j][w].slot = bslot[j]; }
    }
    __syncthreads();
    if (threadIdx.x < PSQ && g * PSQ + (int)threadIdx.x < T) {
        const int j = threadIdx.x;
        PlCand b = sb[j][0];
        for (int k = 1; k < 8; k++)
            if (sb[j][k].add < b.add || (sb[j][k].add == b.add && sb[j][k].slot < b.slot)) b = sb[j][k];
        b.frac = 0.0;
        part[(size_t)slice * T + g * PSQ + j] = b;     // (add >= 2: no candidate in this slice)
    }
}

__global__ void __launch_bounds__(PL_THREADS)
place_spec_seq_kernel(int* head, int* e, int* nxt, int* belong, double* len, int* cid, double* cdis, int* rev,
                      const double* __restrict__ rows, size_t ld, int row_base, int t0, int T, int node_off, PlShared* ps,
                      const PlCand* __restrict__ part, int nslices, int* q_node, int* q_from, double* q_dis, PlDirty dirty,
                      int* next_out, int first_split_tip, int profile) {
    __shared__ PlCand sb[PL_THREADS / 32];
    __shared__ PlCand s_clean;
    __shared__ unsigned int s_tail;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    long long tm = clock64();
#define PS_MARK(k)                                                                  \
    do {                                                                            \
        if (profile && tid == 0) {                                                  \
            const long long now__ = clock64();                                      \
            ps->cyc[k] += (unsigned long long)(now__ - tm);                         \
            tm = now__;                                                             \
        }                                                                           \
    } while (0)
    auto better = [](double a, int s, double oa, int os) { return oa < a || (oa == a && os < s); };
    for (int t = t0; t < t0 + T; t++) {
        const double* dis = rows + (size_t)(t - row_base) * ld;
        // best slot of the batch scoring: first minimum over the slices
        if (w == 0) {
            double a = 2.0; int sl = 0;
            for (int s = lane; s < nslices; s += 32) {
                const PlCand c = part[(size_t)s * T + (t - t0)];
                if (c.add < 2.0 && better(a, sl, c.add, c.slot)) { a = c.add; sl = c.slot; }
            }
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) {
                const double oa = __shfl_xor_sync(0xffffffffu, a, s);
                const int os = __shfl_xor_sync(0xffffffffu, sl, s);
                if (better(a, sl, oa, os)) { a = oa; sl = os; }
            }
            if (lane == 0) { s_clean.add = a; s_clean.slot = a < 2.0 ? sl : 0; }
        }
        __syncthreads();
        const int cs = s_clean.slot;
        const unsigned int nd = *reinterpret_cast<volatile unsigned int*>(dirty.count);
        if ((cs != 0 && __ldcg(&dirty.flag[cs]) == dirty.gen) || nd + 64u > (unsigned int)dirty.cap) {
            // the batch's answer for this tip is stale (or the dirty list is about to overflow): a new batch starts here
            if (tid == 0) *next_out = t;
            return;
        }
        PS_MARK(0);
        double badd = 2.0, bfrac = 0.0;
        int bslot = 0;
        if (tid == 0 && cs != 0) {
            double f, a;
            score_slot(dis, cid, cdis, len, rev, cs, f, a);
            if (a < 2.0) { badd = a; bfrac = f; bslot = cs; }
        }
        for (unsigned int k = tid; k < nd; k += PL_THREADS) {
            const int q = __ldcg(&dirty.list[k]);
            if (__ldcg(&belong[q]) > __ldcg(&e[q])) {
                double f, a;
                score_slot(dis, cid, cdis, len, rev, q, f, a);
                if (a < badd || (a == badd && q < bslot)) { badd = a; bfrac = f; bslot = q; }
            }
        }
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
            const double oa = __shfl_xor_sync(0xffffffffu, badd, s), of = __shfl_xor_sync(0xffffffffu, bfrac, s);
            const int os = __shfl_xor_sync(0xffffffffu, bslot, s);
            if (oa < badd || (oa == badd && os < bslot)) { badd = oa; bfrac = of; bslot = os; }
        }
        if (lane == 0) { sb[w].add = badd; sb[w].frac = bfrac; sb[w].slot = bslot; }
        __syncthreads();
        PS_MARK(1);
        if (profile && tid == 0) ps->bfs_nodes += nd;      // (dirty slots re-scored)
        if (w == 0) {
            PlCand b = sb[0];
            for (int k = 1; k < PL_THREADS / 32; k++)
                if (sb[k].add < b.add || (sb[k].add == b.add && sb[k].slot < b.slot)) b = sb[k];
            if (!(b.add < 2.0)) { b.add = 2.0; b.frac = 0.0; b.slot = 0; }   // the (0,0,2) tuple at position 0 wins
            const int idx = ps->idx;
            const int xe = b.slot, ye = rev[xe];
            split_edge_warp(head, nxt, e, len, cdis, cid, belong, rev, b.slot, b.frac, b.add, t, idx, node_off);
            if (lane == 0) {
                ps->idx = idx + 4;
                // the three edges the split made: x - middle (xe / c0), y - middle (ye / c1), tip - middle (c2 / c3);
                // middle is the newest node, so the slots leaving it (c0, c1, c3) are the candidates
                pl_mark_dirty(dirty, idx); pl_mark_dirty(dirty, idx + 1); pl_mark_dirty(dirty, idx + 3);
                pl_mark_dirty(dirty, xe); pl_mark_dirty(dirty, ye);      // (no longer candidates: their stamp ends a batch that chose them)
            }
            __threadfence_block();
        }
        __syncthreads();
        PS_MARK(5);
        bfs_closest(head, nxt, e, len, cdis, cid, t, q_node, q_from, q_dis, &s_tail, nullptr, node_off, first_split_tip + node_off - 1, &dirty, rev);
        __threadfence();
        __syncthreads();
        PS_MARK(2);
    }
    if (tid == 0) *next_out = t0 + T;
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
    // DIPB_PLACE_SPEC=1: speculative batches (see place_spec_seq_kernel) instead of the per-tip cooperative kernel.  Exact
    // (same arrays, tests/test_placement_gpu.py runs both) but NOT faster as it stands, so it is off by default: at 30 000
    // tips 70 % of the 64-tip batches end early (35 tips per batch), the batch scoring costs as much per placed tip as the
    // per-tip scoring it replaces (3.2 us), and the sequential kernel spends 11.7 k cycles per tip re-scoring 388 dirty
    // slots on one SM plus 17.9 k on the split (five returning atomics for the dirty marks) and 21.7 k on the BFS:
    // 29.9 k tips/s against 43-51 k for the per-tip kernel (profiles/r2_placement_experiments.json).
    const char* esp = getenv("DIPB_PLACE_SPEC");
    const bool spec = esp && atoi(esp) != 0;
    const char* etb = getenv("DIPB_PLACE_BATCH");
    const int TB = etb && atoi(etb) > 0 ? atoi(etb) : 64;
    constexpr int DCAP = 1 << 16;
    int *dflag = nullptr, *dlist = nullptr, *dnext = nullptr;
    unsigned int* dcount = nullptr;
    PlCand* part = nullptr;
    int max_slices = 0, dgen = 0;
    long long n_batches = 0, n_early = 0;
    double t_score = 0;
    if (spec) {
        DIPB_CUDA(pool_alloc(c, (void**)&dflag, sizeof(int) * 8 * (size_t)n_alloc));
        DIPB_CUDA(cudaMemsetAsync(dflag, 0, sizeof(int) * 8 * (size_t)n_alloc, c->stream));
        DIPB_CUDA(pool_alloc(c, (void**)&dlist, sizeof(int) * DCAP));
        DIPB_CUDA(pool_alloc(c, (void**)&dcount, sizeof(unsigned int) * 2));
        dnext = reinterpret_cast<int*>(dcount + 1);
        max_slices = 2 * G;
        DIPB_CUDA(pool_alloc(c, (void**)&part, sizeof(PlCand) * (size_t)max_slices * TB));
    }
    for (int i0 = first_tip; i0 < end && !rc; i0 += batch) {
        int i1 = i0 + batch < end ? i0 + batch : end;
        const double* rows; int row_base; size_t ldr;
        rc = place_fetch_rows(src, i0, i1, buf, ld, &rows, &row_base, &ldr);
        if (rc) break;
        int node_off = n_alloc;
        if (!spec) {
            void* args[] = {&t->head, &t->e, &t->nxt, &t->belong, &t->len, &t->cid, &t->cdis, &t->rev, &rows, &ldr, &row_base,
                            &i0, &i1, &node_off, &sc->ps, &cb, &sc->q_node, &sc->q_from, &sc->q_dis, &gen0, &profile, &first_tip};
            cudaError_t e = cudaLaunchCooperativeKernel((void*)place_batch_kernel, dim3(G), dim3(PL_THREADS), args, 0, c->stream);
            if (e != cudaSuccess) { set_error("placement: cooperative launch failed: %s", cudaGetErrorString(e)); rc = DIPB_E_CUDA; break; }
            c->launches++;
            gen0 += 2u * (unsigned int)(i1 - i0);
            continue;
        }
        for (int t0 = i0; t0 < i1 && !rc;) {
            const int T = t0 + TB < i1 ? TB : i1 - t0;
            const int nslots = 4 * t0 - 4;
            const int groups = (T + PSQ - 1) / PSQ;
            // enough CTAs to fill the GPU, at least ~8 candidate slots per thread and slice
            int nslices = (2 * G + groups - 1) / groups;
            const int by_work = (nslots / 2 + 256 * 8 - 1) / (256 * 8);
            if (nslices > by_work) nslices = by_work < 1 ? 1 : by_work;
            if (nslices > max_slices) nslices = max_slices;
            if (profile) cudaEventRecord(c->ev0, c->stream);
            place_spec_score_kernel<<<dim3(groups, nslices), 256, 0, c->stream>>>(t->e, t->belong, t->len, t->cid, t->cdis, t->rev, nslots, rows, ldr, row_base, t0, T, part);
            if (profile) cudaEventRecord(c->ev1, c->stream);
            DIPB_CUDA(cudaMemsetAsync(dcount, 0, sizeof(unsigned int), c->stream));
            PlDirty dirty{dflag, dlist, dcount, ++dgen, DCAP};
            place_spec_seq_kernel<<<1, PL_THREADS, 0, c->stream>>>(t->head, t->e, t->nxt, t->belong, t->len, t->cid, t->cdis, t->rev, rows, ldr, row_base, t0, T,
                                                                   node_off, sc->ps, part, nslices, sc->q_node, sc->q_from, sc->q_dis, dirty, dnext, first_tip, profile);
            c->launches += 2;
            int next = 0;
            if (cudaMemcpyAsync(&next, dnext, sizeof(int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) {
                set_error("placement: batch [%d, %d) failed: %s", t0, t0 + T, cudaGetErrorString(cudaGetLastError()));
                rc = DIPB_E_CUDA;
                break;
            }
            if (next <= t0 || next > t0 + T) { set_error("placement: batch [%d, %d) made no progress (next = %d)", t0, t0 + T, next); rc = DIPB_E_STATE; break; }
            n_batches++;
            if (next < t0 + T) n_early++;
            if (profile) { float ms = 0; cudaEventElapsedTime(&ms, c->ev0, c->ev1); t_score += ms; }
            t0 = next;
        }
    }
    if (spec) {
        if (profile) fprintf(stderr, "[placement] speculative batches of %d tips: %lld batches, %lld ended early (%.1f tips per batch); batch scoring kernels %.1f ms in total\n", TB, n_batches, n_early,
                             n_batches ? (double)(end - first_tip) / (double)n_batches : 0.0, t_score);
        pool_free(c, dflag); pool_free(c, dlist); pool_free(c, dcount); pool_free(c, part);
    }
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (!rc && e != cudaSuccess) { set_error("placement: %s", cudaGetErrorString(e)); rc = DIPB_E_CUDA; }
    if (!rc && profile) {
        PlShared hs;
        if (cudaMemcpy(&hs, sc->ps, sizeof(hs), cudaMemcpyDeviceToHost) == cudaSuccess) {
            const double tips = (double)(end - first_tip);
            if (spec) fprintf(stderr, "[placement] sequential kernel, cycles per tip: batch answer + staleness check %.0f, dirty re-scoring %.0f (%.0f slots), split %.0f, BFS %.0f\n",
                              hs.cyc[0] / tips, hs.cyc[1] / tips, hs.bfs_nodes / tips, hs.cyc[5] / tips, hs.cyc[2] / tips);
            else fprintf(stderr, "[placement] tips %d..%d: cycles per tip (CTA 0): score %.0f, barrier 1 %.0f, argmin %.0f, split %.0f, BFS %.0f, barrier 2 %.0f; BFS levels %.1f, queue entries %.1f per tip\n",
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
