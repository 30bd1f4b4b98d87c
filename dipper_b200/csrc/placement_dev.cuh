// Device-side building blocks of k-closest placement shared by placement.cu and dc.cu.
#pragma once
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
    unsigned long long cyc[6];     // DIPB_PLACE_PROFILE: CTA 0 cycles in scoring, barrier 1, BFS, barrier 2, argmin, split
    unsigned long long bfs_levels, bfs_nodes;
};

// Candidate slots whose score has changed since the current speculative batch started (placement.cu, place_run): a
// generation stamp per slot plus a compact list.
struct PlDirty {
    int* flag;            // [8n] generation of the batch that last touched the slot's edge
    int* list;            // [cap]
    unsigned int* count;
    int gen, cap;
};
__device__ __forceinline__ void pl_mark_dirty(const PlDirty& d, int cand) {
    if (atomicExch(&d.flag[cand], d.gen) != d.gen) {
        const unsigned int pos = atomicAdd(d.count, 1u);
        if (pos < (unsigned int)d.cap) d.list[pos] = cand;
    }
}

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

// updateTreeStructure (:446-528), one thread
__device__ __forceinline__ void split_edge(int* head, int* nxt, int* e, double* len, double* cdis, int* cid, int* belong, int* rev,
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

// updateTreeStructure (:446-528) by one WARP: the single-thread version above is a chain of ~25 dependent L2 round
// trips (12 500 cycles per tip at 30 000 tips); here the two 5-entry lists are loaded by five lanes at once, the scalars
// once, and the merge of the new leaf's list runs in lane 0's registers.  Same writes, same values.
__device__ __forceinline__ void split_edge_warp(int* head, int* nxt, int* e, double* len, double* cdis, int* cid, int* belong, int* rev,
                                                int eid, double fracLen, double addLen, int placeId, int edgeCount, int node_off) {
    const int lane = threadIdx.x & 31;
    const int middle = placeId + node_off - 1, outside = placeId;
    const int xe = eid, ye = rev[eid];                       // (all lanes read the same words: one transaction each)
    const int x = belong[eid], y = e[eid];
    const double orig = len[eid], len_ye = len[ye];
    const int h_mid = head[middle], h_out = head[outside];
    const int c0 = edgeCount, c1 = edgeCount + 1, c2 = edgeCount + 2, c3 = edgeCount + 3;
    int iy = -1, ix = -1;
    double dy = 2.0, dx = 2.0;
    if (lane < KC5) { iy = cid[ye * KC5 + lane]; dy = cdis[ye * KC5 + lane]; ix = cid[xe * KC5 + lane]; dx = cdis[xe * KC5 + lane]; }
    // lists of the two halves of the split edge seen from the new inner node (entries without a leaf keep the initial (2, -1))
    const int c0i = iy, c1i = ix;
    const double c0d = iy != -1 ? dy + orig - fracLen : 2.0, c1d = ix != -1 ? dx + fracLen : 2.0;
    if (lane < KC5) {
        cid[c0 * KC5 + lane] = c0i; cdis[c0 * KC5 + lane] = c0d;
        cid[c1 * KC5 + lane] = c1i; cdis[c1 * KC5 + lane] = c1d;
    }
    // list of the slot that leaves the new leaf: c1's entries, then c0's, inserted in order (list_insert semantics)
    double ld[KC5];
    int li[KC5];
#pragma unroll
    for (int k = 0; k < KC5; k++) { ld[k] = 2.0; li[k] = -1; }
    bool open = true;
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
        open = true;
#pragma unroll
        for (int k = 0; k < KC5; k++) {
            const double d = __shfl_sync(0xffffffffu, pass == 0 ? c1d : c0d, k);
            const int id = __shfl_sync(0xffffffffu, pass == 0 ? c1i : c0i, k);
            if (id == -1) open = false;
            if (open) {
                bool done = false;
#pragma unroll
                for (int j = 0; j < KC5; j++) {
                    if (!done && ld[j] > d) {
#pragma unroll
                        for (int q = KC5 - 1; q > j; q--) { ld[q] = ld[q - 1]; li[q] = li[q - 1]; }
                        ld[j] = d; li[j] = id;
                        done = true;
                    }
                }
            }
        }
    }
    if (lane == 0) {
        e[xe] = middle; len[xe] = fracLen;
        e[ye] = middle; len[ye] = len_ye - fracLen;
        // link_slot x 4 with the list heads kept in registers
        e[c0] = x; len[c0] = fracLen; nxt[c0] = h_mid; belong[c0] = middle;
        e[c1] = y; len[c1] = orig - fracLen; nxt[c1] = c0; belong[c1] = middle;
        e[c2] = middle; len[c2] = addLen; nxt[c2] = h_out; belong[c2] = outside; head[outside] = c2;
        e[c3] = outside; len[c3] = addLen; nxt[c3] = c1; belong[c3] = middle; head[middle] = c3;
#pragma unroll
        for (int k = 0; k < KC5; k++) { cdis[c3 * KC5 + k] = ld[k]; cid[c3 * KC5 + k] = li[k]; }
        rev[xe] = c0; rev[c0] = xe; rev[ye] = c1; rev[c1] = ye; rev[c2] = c3; rev[c3] = c2;
    }
    __syncwarp();
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


struct PlaceScratch {
    PlShared* ps = nullptr;
    int *q_node = nullptr, *q_from = nullptr;
    double* q_dis = nullptr;
};
int tree_alloc(dipb_ctx* c, int n, dipb_tree** out);
int place_scratch_alloc(dipb_ctx* c, int n, PlaceScratch* s);
void place_scratch_free(dipb_ctx* c, PlaceScratch* s);
// builds the 2-leaf tree from d(1,0) and places tips [2, end) (src/placement_close_k.cu:646-854)
int place_from_scratch(dipb_ctx* c, const dipb_dist_source* src, int n_alloc, int end, dipb_tree* t, PlaceScratch* sc);
int check_source(const dipb_dist_source* s, int n);
// rows [r0, r1) x cols [0, r1) of the selected provider into buf (or the matrix itself)
int place_fetch_rows(const dipb_dist_source* src, int r0, int r1, double* buf, size_t ld, const double** rows, int* row_base,
                     size_t* ld_out);

}  // namespace dipb
