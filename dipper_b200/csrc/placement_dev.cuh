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
void place_scratch_free(PlaceScratch* s);
// builds the 2-leaf tree from d(1,0) and places tips [2, end) (src/placement_close_k.cu:646-854)
int place_from_scratch(dipb_ctx* c, const dipb_dist_source* src, int n_alloc, int end, dipb_tree* t, PlaceScratch* sc);
int check_source(const dipb_dist_source* s, int n);

}  // namespace dipb
