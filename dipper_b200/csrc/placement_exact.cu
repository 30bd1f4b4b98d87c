// Exact placement mode (-p 0): replaces PlacementDeviceArrays::findPlacementTree, src/placement.cu:505-789, with its
// kernels initialize / buildInitialTree / updateFromBottomToTop / updateFromTopToBottom / calculateBranchLength /
// updateTreeStructure / updateDfsRk / findEndRk / updateDepth / updateLevelStEd (:118-436) and the thrust min_element,
// reduce and stable_sort_by_key calls between them.
//
// The reference keeps the tree rooted at the first internal node and, for every tip, (1) sweeps the levels bottom-up
// and top-down (2 x depth kernel launches) so that lim[slot] = max over ALL leaves behind the slot's source of
// (distance - path), (2) scores every parent->child slot with those two limits, takes the first minimum of the pendant
// length, splits that edge, and (3) renumbers preorder ranks, bumps the depth of the split child's subtree and re-sorts
// the breadth-first order on the device -- about 2 x depth + 12 launches, 4 device->host copies and one sort per tip.
//
// Here ONE 16-CTA thread-block cluster runs the whole loop for a batch of tips with the tree in (distributed) shared
// memory.  Every node belongs to one thread of one CTA (32-node chunks dealt round-robin over the CTAs) which keeps
// its parent, edge length, depth, preorder rank and subtree size; a level step is "the owners of the nodes at this
// depth compute and PUSH the value into the parent's / children's slot in the owner CTA's shared memory
// (st.shared::cluster)" followed by one hardware cluster barrier, so a tip costs 2 x depth + 2 barriers of ~0.3 us and
// no global-memory round trip.  Preorder ranks / subtree sizes make the rank shift, the ancestor size update and the
// subtree depth bump three independent per-node tests (no sort, no reduction).  The reference slot arrays
// (head / e / nxt / belong / len) are written with the slot numbers the reference would produce (they are arithmetic:
// tip t appends slots 4t-4 .. 4t-1), so the exported tree prints the same Newick text.
//
// Arithmetic is the reference's, expression by expression: the per-edge `lim - len` subtractions happen in the same
// order along every path and max() is exact, so limits, pendant lengths, the argmin (ties -> smallest slot) and all
// branch lengths are bit-identical to the oracle restatement (oracle/dipper_oracle.c: orc_place_exact_all).
// Inputs beyond the cluster's shared memory (49 152 tips) run the same data flow through global memory on the whole grid
// (place_exact_global_kernel).  One deviation: when no candidate has a pendant length < 2 the reference "places" on slot 0 through its (0,0,2)
// default tuple and corrupts its depth table (the no-op swap at :246-249); this kernel reports DIPB_E_UNSUPPORTED.
#include <cooperative_groups.h>
#include <cstdlib>
#include "common.cuh"
#include "placement_dev.cuh"

namespace dipb {

namespace {

constexpr int EX_THREADS = 1024;
constexpr int EX_KM = 3;               // owned nodes of each kind per thread at most (NL <= 3072)
constexpr int EX_BYTES_PER_NODE = 90;
constexpr unsigned long long EX_EMPTY = 0xFFF8DEADBEEF0001ull;   // "no value yet": a NaN payload no arithmetic produces  // per local index: leaf 31 B + internal 59 B of state

struct ExCtl {
    int maxdep;
    int error;
    unsigned long long levels;   // sum over tips of the tree depth (profile)
    unsigned long long cycles;   // CTA 0, whole batch loop
};

struct ExState {   // views into this CTA's dynamic shared memory
    double *l_dn, *l_plen, *i_dn, *i_in0, *i_in1, *i_plen;
    int *l_par, *l_sdn, *l_rank, *i_par, *i_kid0, *i_kid1, *i_sdn, *i_rank, *i_sz;
    unsigned short *l_dep, *i_dep;
    unsigned char *l_cidx, *i_cidx;
};

__device__ __forceinline__ ExState ex_carve(unsigned char* base, int NL) {
    ExState s;
    double* d = reinterpret_cast<double*>(base);
    s.l_dn = d; s.l_plen = d + NL; s.i_dn = d + 2 * NL; s.i_in0 = d + 3 * NL; s.i_in1 = d + 4 * NL; s.i_plen = d + 5 * NL;
    int* q = reinterpret_cast<int*>(d + 6 * NL);
    s.l_par = q; s.l_sdn = q + NL; s.l_rank = q + 2 * NL; s.i_par = q + 3 * NL; s.i_kid0 = q + 4 * NL; s.i_kid1 = q + 5 * NL;
    s.i_sdn = q + 6 * NL; s.i_rank = q + 7 * NL; s.i_sz = q + 8 * NL;
    unsigned short* h = reinterpret_cast<unsigned short*>(q + 9 * NL);
    s.l_dep = h; s.i_dep = h + NL;
    unsigned char* b = reinterpret_cast<unsigned char*>(h + 2 * NL);
    s.l_cidx = b; s.i_cidx = b + NL;
    return s;
}

__device__ __forceinline__ unsigned int ex_peer(const void* p, int rank) {
    const unsigned int a = (unsigned int)__cvta_generic_to_shared(p);
    unsigned int ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    return ra;
}
__device__ __forceinline__ void ex_push_f64(double* p, int rank, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ex_peer(p, rank)), "d"(v) : "memory");
}
__device__ __forceinline__ void ex_push_s32(int* p, int rank, int v) {
    asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(ex_peer(p, rank)), "r"(v) : "memory");
}
__device__ __forceinline__ void ex_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

struct ExBest {
    double add, frac;
    int slot, node;
};
__device__ __forceinline__ bool ex_before(double a_add, int a_slot, double b_add, int b_slot) {
    return a_add < b_add || (a_add == b_add && a_slot < b_slot);   // min_element: first minimum in slot order
}

// calculateBranchLength :153-198 for the parent->child slot of one node (dis1 = limit on the parent side)
__device__ __forceinline__ void ex_score(double dis1, double dis2, double L, int slot, int node, ExBest& b) {
    double a = (dis1 + dis2 - L) / 2;
    if (a < 0) a = 0;
    dis1 -= a; dis2 -= a;
    if (dis1 < 0) dis1 = 0;
    if (dis2 < 0) dis2 = 0;
    if (dis1 > L) { a += dis1 - L; dis1 = L; }
    if (dis2 > L) { a += dis2 - L; dis2 = L; }
    const double rest = L - dis1 - dis2;
    dis1 += rest / 2;
    if (ex_before(a, slot, b.add, b.slot)) { b.add = a; b.frac = dis1; b.slot = slot; b.node = node; }
}

template <int CS, bool FLOW>
__global__ void __launch_bounds__(EX_THREADS, 1)
place_exact_kernel(int* __restrict__ head, int* __restrict__ e, int* __restrict__ nxt, int* __restrict__ belong,
                   double* __restrict__ len, const double* __restrict__ rows, size_t ld, int row_base, int i0, int i1, int N,
                   int NL, const double* __restrict__ d01, uint4* __restrict__ saved, ExCtl* __restrict__ ctl, int /*backoff_ns*/) {
    extern __shared__ __align__(16) unsigned char ex_smem[];
    __shared__ double rec_d[2][CS][3];  // every CTA's best candidate of this tip: add, frac, edge length (double-buffered by tip parity:
    __shared__ int rec_i[2][CS][8];     // slot, node y, parent x, rank, subtree size, depth, child index of y   in FLOW mode a CTA that owns
                                        // no nodes yet can be a whole tip behind its peers)
    __shared__ int s_grow[2];
    __shared__ ExBest s_warp[EX_THREADS / 32];
    constexpr int LOGCS = CS == 16 ? 4 : (CS == 8 ? 3 : (CS == 4 ? 2 : (CS == 2 ? 1 : 0)));
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int rank = (int)cooperative_groups::this_cluster().block_rank();
    const ExState S = ex_carve(ex_smem, NL);
    const size_t state_words = (size_t)NL * EX_BYTES_PER_NODE / 16;
    uint4* sm4 = reinterpret_cast<uint4*>(ex_smem);
    auto cta_of = [&](int id) { return (id >> 5) & (CS - 1); };
    auto loc_of = [&](int id) { return ((id >> (5 + LOGCS)) << 5) | (id & 31); };
    auto id_of = [&](int k) { return ((((k >> 5) << LOGCS) | rank) << 5) | (k & 31); };

    // ---- state: fresh 2-leaf tree (buildInitialTree :253-295) or the previous batch's
    if (i0 == 2) {
        for (int k = tid; k < NL; k += EX_THREADS) {
            S.l_dep[k] = 0xFFFF; S.i_dep[k] = 0xFFFF; S.l_rank[k] = -1; S.i_rank[k] = -1; S.i_sz[k] = 0;
            S.l_par[k] = -1; S.i_par[k] = -1; S.i_kid0[k] = -1; S.i_kid1[k] = -1; S.l_sdn[k] = 0; S.i_sdn[k] = 0;
            S.l_cidx[k] = 0; S.i_cidx[k] = 0;
            const double empty = FLOW ? __longlong_as_double((long long)EX_EMPTY) : 0.0;
            S.l_dn[k] = empty; S.i_dn[k] = empty; S.i_in0[k] = empty; S.i_in1[k] = empty; S.l_plen[k] = 0; S.i_plen[k] = 0;
        }
        __syncthreads();
        if (rank == 0 && tid == 0) {   // nodes 0, 1 (leaves) and N (internal index 0) all live in CTA 0, local 0 / 1 / 0
            const double d = d01[0];
            S.i_dep[0] = 0; S.i_rank[0] = 0; S.i_sz[0] = 3; S.i_kid0[0] = 0; S.i_kid1[0] = 1;
            S.l_par[0] = N; S.l_cidx[0] = 0; S.l_plen[0] = d / 2; S.l_sdn[0] = 2; S.l_rank[0] = 1; S.l_dep[0] = 1;
            S.l_par[1] = N; S.l_cidx[1] = 1; S.l_plen[1] = d / 2; S.l_sdn[1] = 3; S.l_rank[1] = 2; S.l_dep[1] = 1;
            e[0] = N; len[0] = d / 2; nxt[0] = -1; belong[0] = 0; head[0] = 0;
            e[1] = N; len[1] = d / 2; nxt[1] = -1; belong[1] = 1; head[1] = 1;
            e[2] = 0; len[2] = d / 2; nxt[2] = -1; belong[2] = N;
            e[3] = 1; len[3] = d / 2; nxt[3] = 2; belong[3] = N; head[N] = 3;
        }
    } else {
        const uint4* src = saved + (size_t)rank * state_words;
        for (size_t w = tid; w < state_words; w += EX_THREADS) sm4[w] = src[w];
    }
    if (tid < 2) s_grow[tid] = 0;
    int maxdep = i0 == 2 ? 1 : ctl->maxdep;
    bool failed = false;
    unsigned long long levels = 0;
    const long long t_begin = clock64();
    __syncthreads();
    ex_cluster_sync();

    for (int i = i0; i < i1; i++) {
        // distances of this tip to the leaves this thread owns
        const double* row = rows + (size_t)(i - row_base) * ld;
        double myd[EX_KM];
#pragma unroll
        for (int u = 0; u < EX_KM; u++) {
            const int k = tid + u * EX_THREADS;
            myd[u] = 0;
            if (k < NL) { const int lid = id_of(k); if (lid < i) myd[u] = __ldg(&row[lid]); }
        }
        ExBest best;
        best.add = 2.0; best.frac = 0.0; best.slot = 0; best.node = -1;   // the (0,0,2) default tuple
        if constexpr (FLOW) {
            // ---- data flow instead of level steps: a value IS its own arrival flag (slots hold EX_EMPTY, a NaN no arithmetic
            // produces, until the producer's st.shared::cluster lands).  Leaves push their limit at once; an internal node
            // waits for both children (updateFromBottomToTop), pushes upwards, then waits for its parent's value
            // (updateFromTopToBottom), scores its parent->child slot and feeds its children.  No barrier, no fence: the
            // bottom-up / top-down sweeps cost one shared-memory store latency per tree level.
            int st_l[EX_KM], st_i[EX_KM], pending = 0;
            double ra[EX_KM], rb[EX_KM];
#pragma unroll
            for (int u = 0; u < EX_KM; u++) {
                const int k = tid + u * EX_THREADS;
                st_l[u] = 0; st_i[u] = 0; ra[u] = 0; rb[u] = 0;
                if (k < NL) {
                    if (S.l_dep[k] != 0xFFFF) {
                        const int j = S.l_par[k] - N;
                        ex_push_f64((S.l_cidx[k] ? S.i_in1 : S.i_in0) + loc_of(j), cta_of(j), myd[u] - S.l_plen[k]);
                        st_l[u] = 2; pending++;
                    }
                    if (S.i_dep[k] != 0xFFFF) { st_i[u] = 1; pending++; }
                }
            }
            while (pending) {
#pragma unroll
                for (int u = 0; u < EX_KM; u++) {
                    const int k = tid + u * EX_THREADS;
                    if (st_i[u] == 1) {
                        const unsigned long long ua = *reinterpret_cast<volatile unsigned long long*>(S.i_in0 + k);
                        const unsigned long long ub = *reinterpret_cast<volatile unsigned long long*>(S.i_in1 + k);
                        if (ua != EX_EMPTY && ub != EX_EMPTY) {
                            const double a = __longlong_as_double((long long)ua), b = __longlong_as_double((long long)ub);
                            *reinterpret_cast<volatile unsigned long long*>(S.i_in0 + k) = EX_EMPTY;
                            *reinterpret_cast<volatile unsigned long long*>(S.i_in1 + k) = EX_EMPTY;
                            ra[u] = a; rb[u] = b;
                            if (S.i_dep[k] == 0) {   // the root: nothing above it
                                double v0 = 0, v1 = 0;
                                if (b > v0) v0 = b;
                                if (a > v1) v1 = a;
                                const int k0 = S.i_kid0[k], k1 = S.i_kid1[k];
                                if (k0 < N) ex_push_f64(S.l_dn + loc_of(k0), cta_of(k0), v0);
                                else ex_push_f64(S.i_dn + loc_of(k0 - N), cta_of(k0 - N), v0);
                                if (k1 < N) ex_push_f64(S.l_dn + loc_of(k1), cta_of(k1), v1);
                                else ex_push_f64(S.i_dn + loc_of(k1 - N), cta_of(k1 - N), v1);
                                st_i[u] = 0; pending--;
                            } else {
                                double m = 0;
                                if (a > m) m = a;
                                if (b > m) m = b;
                                const int j = S.i_par[k] - N;
                                ex_push_f64((S.i_cidx[k] ? S.i_in1 : S.i_in0) + loc_of(j), cta_of(j), m - S.i_plen[k]);
                                st_i[u] = 2;
                            }
                        }
                    } else if (st_i[u] == 2) {
                        const unsigned long long ud = *reinterpret_cast<volatile unsigned long long*>(S.i_dn + k);
                        if (ud != EX_EMPTY) {
                            *reinterpret_cast<volatile unsigned long long*>(S.i_dn + k) = EX_EMPTY;
                            const double dn = __longlong_as_double((long long)ud), a = ra[u], b = rb[u];
                            double up = 0;
                            if (a > up) up = a;
                            if (b > up) up = b;
                            const double pl = S.i_plen[k];
                            ex_score(dn, up, pl, S.i_sdn[k], N + id_of(k), best);
                            const double basev = dn - pl;
                            double v0 = 0, v1 = 0;
                            if (b > v0) v0 = b;
                            if (a > v1) v1 = a;
                            if (basev > v0) v0 = basev;
                            if (basev > v1) v1 = basev;
                            const int k0 = S.i_kid0[k], k1 = S.i_kid1[k];
                            if (k0 < N) ex_push_f64(S.l_dn + loc_of(k0), cta_of(k0), v0);
                            else ex_push_f64(S.i_dn + loc_of(k0 - N), cta_of(k0 - N), v0);
                            if (k1 < N) ex_push_f64(S.l_dn + loc_of(k1), cta_of(k1), v1);
                            else ex_push_f64(S.i_dn + loc_of(k1 - N), cta_of(k1 - N), v1);
                            st_i[u] = 0; pending--;
                        }
                    }
                    if (st_l[u] == 2) {
                        const unsigned long long ud = *reinterpret_cast<volatile unsigned long long*>(S.l_dn + k);
                        if (ud != EX_EMPTY) {
                            *reinterpret_cast<volatile unsigned long long*>(S.l_dn + k) = EX_EMPTY;
                            ex_score(__longlong_as_double((long long)ud), myd[u], S.l_plen[k], S.l_sdn[k], id_of(k), best);
                            st_l[u] = 0; pending--;
                        }
                    }
                }
            }
        } else {
        // ---- bottom-up (updateFromBottomToTop :298-332): a node's limit towards its parent, pushed to the parent
        for (int lev = maxdep; lev >= 1; lev--) {
#pragma unroll
            for (int u = 0; u < EX_KM; u++) {
                const int k = tid + u * EX_THREADS;
                if (k < NL) {
                    if (S.l_dep[k] == lev) {
                        const double req = myd[u] - S.l_plen[k];
                        const int j = S.l_par[k] - N;
                        ex_push_f64((S.l_cidx[k] ? S.i_in1 : S.i_in0) + loc_of(j), cta_of(j), req);
                    }
                    if (S.i_dep[k] == lev) {
                        double m = 0;
                        const double a = S.i_in0[k], b = S.i_in1[k];
                        if (a > m) m = a;
                        if (b > m) m = b;
                        const double req = m - S.i_plen[k];
                        const int j = S.i_par[k] - N;
                        ex_push_f64((S.i_cidx[k] ? S.i_in1 : S.i_in0) + loc_of(j), cta_of(j), req);
                    }
                }
            }
            ex_cluster_sync();
        }
        // ---- top-down (updateFromTopToBottom :334-366) fused with the scoring of the parent->child slots
        for (int lev = 0; lev <= maxdep; lev++) {
#pragma unroll
            for (int u = 0; u < EX_KM; u++) {
                const int k = tid + u * EX_THREADS;
                if (k < NL) {
                    if (S.l_dep[k] == lev) ex_score(S.l_dn[k], myd[u], S.l_plen[k], S.l_sdn[k], id_of(k), best);
                    if (S.i_dep[k] == lev) {
                        const double a = S.i_in0[k], b = S.i_in1[k];
                        double basev = 0;
                        if (lev > 0) {
                            double up = 0;
                            if (a > up) up = a;
                            if (b > up) up = b;
                            const double dn = S.i_dn[k];
                            ex_score(dn, up, S.i_plen[k], S.i_sdn[k], N + id_of(k), best);
                            basev = dn - S.i_plen[k];
                        }
                        double v0 = 0, v1 = 0;
                        if (b > v0) v0 = b;
                        if (a > v1) v1 = a;
                        if (lev > 0) { if (basev > v0) v0 = basev; if (basev > v1) v1 = basev; }
                        const int k0 = S.i_kid0[k], k1 = S.i_kid1[k];
                        if (k0 < N) ex_push_f64(S.l_dn + loc_of(k0), cta_of(k0), v0);
                        else ex_push_f64(S.i_dn + loc_of(k0 - N), cta_of(k0 - N), v0);
                        if (k1 < N) ex_push_f64(S.l_dn + loc_of(k1), cta_of(k1), v1);
                        else ex_push_f64(S.i_dn + loc_of(k1 - N), cta_of(k1 - N), v1);
                    }
                }
            }
            if (lev < maxdep) ex_cluster_sync();
        }
        }
        levels += (unsigned long long)maxdep;
        // ---- first minimum over the cluster (thrust::min_element :657)
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
            const double oa = __shfl_xor_sync(0xffffffffu, best.add, s), of = __shfl_xor_sync(0xffffffffu, best.frac, s);
            const int os = __shfl_xor_sync(0xffffffffu, best.slot, s), on = __shfl_xor_sync(0xffffffffu, best.node, s);
            if (ex_before(oa, os, best.add, best.slot)) { best.add = oa; best.frac = of; best.slot = os; best.node = on; }
        }
        if (lane == 0) s_warp[wid] = best;
        __syncthreads();
        if (wid == 0) {
            best = s_warp[lane];
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) {
                const double oa = __shfl_xor_sync(0xffffffffu, best.add, s), of = __shfl_xor_sync(0xffffffffu, best.frac, s);
                const int os = __shfl_xor_sync(0xffffffffu, best.slot, s), on = __shfl_xor_sync(0xffffffffu, best.node, s);
                if (ex_before(oa, os, best.add, best.slot)) { best.add = oa; best.frac = of; best.slot = os; best.node = on; }
            }
            // the winner's owner is in this CTA: attach what the update needs
            double pl = 0;
            int px = -1, rk = -1, sz = 0, dp = 0, ci = 0;
            const int y = best.node;
            if (y >= 0) {
                if (y < N) { const int k = loc_of(y); pl = S.l_plen[k]; px = S.l_par[k]; rk = S.l_rank[k]; sz = 1; dp = S.l_dep[k]; ci = S.l_cidx[k]; }
                else { const int k = loc_of(y - N); pl = S.i_plen[k]; px = S.i_par[k]; rk = S.i_rank[k]; sz = S.i_sz[k]; dp = S.i_dep[k]; ci = S.i_cidx[k]; }
            }
            if (lane < CS) {
                ex_push_f64(&rec_d[i & 1][rank][0], lane, best.add); ex_push_f64(&rec_d[i & 1][rank][1], lane, best.frac); ex_push_f64(&rec_d[i & 1][rank][2], lane, pl);
                ex_push_s32(&rec_i[i & 1][rank][0], lane, best.slot); ex_push_s32(&rec_i[i & 1][rank][1], lane, y); ex_push_s32(&rec_i[i & 1][rank][2], lane, px);
                ex_push_s32(&rec_i[i & 1][rank][3], lane, rk); ex_push_s32(&rec_i[i & 1][rank][4], lane, sz); ex_push_s32(&rec_i[i & 1][rank][5], lane, dp);
                ex_push_s32(&rec_i[i & 1][rank][6], lane, ci);
            }
        }
        ex_cluster_sync();
        int w = 0;
        {
            double wa = rec_d[i & 1][0][0];
            int ws = rec_i[i & 1][0][0];
#pragma unroll
            for (int c = 1; c < CS; c++) {
                const double ca = rec_d[i & 1][c][0];
                const int cs = rec_i[i & 1][c][0];
                if (ex_before(ca, cs, wa, ws)) { wa = ca; ws = cs; w = c; }
            }
        }
        const double addLen = rec_d[i & 1][w][0], fracLen = rec_d[i & 1][w][1], pleny = rec_d[i & 1][w][2];
        const int slot = rec_i[i & 1][w][0], y = rec_i[i & 1][w][1], x = rec_i[i & 1][w][2], r = rec_i[i & 1][w][3], szy = rec_i[i & 1][w][4],
                  depy = rec_i[i & 1][w][5], cidxy = rec_i[i & 1][w][6];
        if (y < 0) { failed = true; break; }   // uniform over the cluster
        // ---- update: ranks (updateDfsRk :368-381), ancestors' sizes, subtree depth (findEndRk / updateDepth :384-417)
        bool grow = false;
        if constexpr (!FLOW) {
#pragma unroll
        for (int u = 0; u < EX_KM; u++) {
            const int k = tid + u * EX_THREADS;
            if (k < NL) {
                int rk = S.l_rank[k];
                if (rk >= r) {
                    if (rk < r + szy) { const int d = S.l_dep[k] + 1; S.l_dep[k] = (unsigned short)d; if (d > maxdep) grow = true; }
                    S.l_rank[k] = rk + 2;
                }
                rk = S.i_rank[k];
                if (rk >= r) {
                    if (rk < r + szy) { const int d = S.i_dep[k] + 1; S.i_dep[k] = (unsigned short)d; if (d > maxdep) grow = true; }
                    S.i_rank[k] = rk + 2;
                } else if (rk >= 0 && r < rk + S.i_sz[k]) S.i_sz[k] += 2;
            }
        }
        }
        __syncthreads();
        // ---- split (updateTreeStructure :200-251): middle m between x and y, new leaf i below m
        const int m = i + N - 1, jm = i - 1, c0 = 4 * i - 4;
        if (tid == 0 && cta_of(x - N) == rank) (cidxy ? S.i_kid1 : S.i_kid0)[loc_of(x - N)] = m;
        if (tid == 32) {
            if (y < N) { if (cta_of(y) == rank) { const int k = loc_of(y); S.l_par[k] = m; S.l_cidx[k] = 0; S.l_plen[k] = pleny - fracLen; S.l_sdn[k] = c0 + 1; } }
            else if (cta_of(y - N) == rank) { const int k = loc_of(y - N); S.i_par[k] = m; S.i_cidx[k] = 0; S.i_plen[k] = pleny - fracLen; S.i_sdn[k] = c0 + 1; }
        }
        if (tid == 64 && cta_of(jm) == rank) {
            const int k = loc_of(jm);
            S.i_par[k] = x; S.i_cidx[k] = (unsigned char)cidxy; S.i_kid0[k] = y; S.i_kid1[k] = i; S.i_plen[k] = fracLen; S.i_sdn[k] = slot;
            S.i_rank[k] = r; S.i_sz[k] = szy + 2; S.i_dep[k] = (unsigned short)(FLOW ? 1 : depy);   // FLOW: depth is only a placed / root marker
        }
        if (tid == 96 && cta_of(i) == rank) {
            const int k = loc_of(i);
            S.l_par[k] = m; S.l_cidx[k] = 1; S.l_plen[k] = addLen; S.l_sdn[k] = c0 + 3; S.l_rank[k] = r + 1; S.l_dep[k] = (unsigned short)(depy + 1);
        }
        if (tid == 128 && rank == (i & (CS - 1))) {
            // the reference's slot arrays; slot numbers are arithmetic: x->y is the winning slot, y->x was appended when y was
            // created (leaf t: 4t-2, internal node of tip t: 4t-4, the first two leaves: 0 and 1)
            const int xe = slot, ye = y < N ? (y < 2 ? y : 4 * y - 2) : 4 * (y - N);
            const int c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3;
            e[xe] = m; len[xe] = fracLen;
            e[ye] = m; len[ye] = pleny - fracLen;
            e[c0] = x; len[c0] = fracLen; nxt[c0] = -1; belong[c0] = m;
            e[c1] = y; len[c1] = pleny - fracLen; nxt[c1] = c0; belong[c1] = m;
            e[c2] = m; len[c2] = addLen; nxt[c2] = -1; belong[c2] = i; head[i] = c2;
            e[c3] = i; len[c3] = addLen; nxt[c3] = c1; belong[c3] = m; head[m] = c3;
        }
        // ---- did the tree get deeper?  (level mode only; in FLOW mode the local __syncthreads is all the next tip needs:
        // peers' early pushes land in slots that are EX_EMPTY again, and nothing else of a peer is read)
        if constexpr (FLOW) {
            __syncthreads();
        } else {
            const int par = i & 1;
            if (tid == 0) s_grow[par ^ 1] = 0;
            const int g = __syncthreads_or(grow ? 1 : 0);
            if (g && tid < CS) ex_push_s32(&s_grow[par], tid, 1);
            ex_cluster_sync();
            if (s_grow[par]) maxdep++;
        }
    }

    // ---- keep the state for the next batch
    __syncthreads();
    ex_cluster_sync();   // no CTA may leave while peers can still push into its shared memory
    {
        uint4* dst = saved + (size_t)rank * state_words;
        for (size_t w = tid; w < state_words; w += EX_THREADS) dst[w] = sm4[w];
    }
    if (rank == 0 && tid == 0) {
        ctl->maxdep = maxdep;
        if (failed) ctl->error = 1;
        ctl->levels += levels;
        ctl->cycles += (unsigned long long)(clock64() - t_begin);
    }
}

// ---- data flow, second version (default): leaves are not nodes of the flow at all.  A leaf's limit towards its parent is
// dist - len, and its own score needs only the parent's values, so the PARENT's owner keeps the edge length and slot of
// its leaf children, prefetches their distances at the start of the tip and scores them when its own value from above
// arrives.  Half the nodes, half the pushes and half the polled slots disappear; cherries fire immediately.
// State per internal node: 74 B (3 value slots, own edge length, 2 leaf-child edge lengths, parent, 2 children, own slot,
// 2 leaf-child slots, child index, placed / root flag) -> 49 152 tips in one 16-CTA cluster.
constexpr int F2_BYTES_PER_NODE = 74;

struct F2State {
    double *dn, *in0, *in1, *plen, *klen0, *klen1;
    int *par, *kid0, *kid1, *sdn, *ksdn0, *ksdn1;
    unsigned char *cidx, *flag;   // flag: 0 not placed, 1 placed, 2 root
};
__device__ __forceinline__ F2State f2_carve(unsigned char* base, int NL) {
    F2State s;
    double* d = reinterpret_cast<double*>(base);
    s.dn = d; s.in0 = d + NL; s.in1 = d + 2 * NL; s.plen = d + 3 * NL; s.klen0 = d + 4 * NL; s.klen1 = d + 5 * NL;
    int* q = reinterpret_cast<int*>(d + 6 * NL);
    s.par = q; s.kid0 = q + NL; s.kid1 = q + 2 * NL; s.sdn = q + 3 * NL; s.ksdn0 = q + 4 * NL; s.ksdn1 = q + 5 * NL;
    unsigned char* b = reinterpret_cast<unsigned char*>(q + 6 * NL);
    s.cidx = b; s.flag = b + NL;
    return s;
}

template <int CS>
__global__ void __launch_bounds__(EX_THREADS, 1)
place_exact_flow2_kernel(int* __restrict__ head, int* __restrict__ e, int* __restrict__ nxt, int* __restrict__ belong,
                         double* __restrict__ len, const double* __restrict__ rows, size_t ld, int row_base, int i0, int i1, int N,
                         int NL, const double* __restrict__ d01, uint4* __restrict__ saved, ExCtl* __restrict__ ctl, int backoff_ns) {
    extern __shared__ __align__(16) unsigned char ex_smem[];
    __shared__ double rec_d[2][CS][3];  // add, frac, edge length of the winner's edge (double-buffered by tip parity)
    __shared__ int rec_i[2][CS][4];     // slot, node y, parent x, child index of y
    __shared__ ExBest s_warp[EX_THREADS / 32];
    __shared__ double s_wpl[EX_THREADS / 32];
    __shared__ int s_wx[EX_THREADS / 32], s_wc[EX_THREADS / 32];
    constexpr int LOGCS = CS == 16 ? 4 : (CS == 8 ? 3 : (CS == 4 ? 2 : (CS == 2 ? 1 : 0)));
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int rank = (int)cooperative_groups::this_cluster().block_rank();
    const F2State S = f2_carve(ex_smem, NL);
    const size_t state_words = (size_t)NL * F2_BYTES_PER_NODE / 16;
    uint4* sm4 = reinterpret_cast<uint4*>(ex_smem);
    auto cta_of = [&](int id) { return (id >> 5) & (CS - 1); };
    auto loc_of = [&](int id) { return ((id >> (5 + LOGCS)) << 5) | (id & 31); };
    auto id_of = [&](int k) { return ((((k >> 5) << LOGCS) | rank) << 5) | (k & 31); };
    const double empty = __longlong_as_double((long long)EX_EMPTY);

    if (i0 == 2) {
        for (int k = tid; k < NL; k += EX_THREADS) {
            S.dn[k] = empty; S.in0[k] = empty; S.in1[k] = empty; S.plen[k] = 0; S.klen0[k] = 0; S.klen1[k] = 0;
            S.par[k] = -1; S.kid0[k] = -1; S.kid1[k] = -1; S.sdn[k] = 0; S.ksdn0[k] = 0; S.ksdn1[k] = 0; S.cidx[k] = 0; S.flag[k] = 0;
        }
        __syncthreads();
        if (rank == 0 && tid == 0) {   // buildInitialTree :253-295: root N (internal index 0) with the leaves 0 and 1
            const double d = d01[0];
            S.flag[0] = 2; S.kid0[0] = 0; S.kid1[0] = 1; S.klen0[0] = d / 2; S.klen1[0] = d / 2; S.ksdn0[0] = 2; S.ksdn1[0] = 3;
            e[0] = N; len[0] = d / 2; nxt[0] = -1; belong[0] = 0; head[0] = 0;
            e[1] = N; len[1] = d / 2; nxt[1] = -1; belong[1] = 1; head[1] = 1;
            e[2] = 0; len[2] = d / 2; nxt[2] = -1; belong[2] = N;
            e[3] = 1; len[3] = d / 2; nxt[3] = 2; belong[3] = N; head[N] = 3;
        }
    } else {
        const uint4* src = saved + (size_t)rank * state_words;
        for (size_t w = tid; w < state_words; w += EX_THREADS) sm4[w] = src[w];
    }
    bool failed = false;
    const long long t_begin = clock64();
    __syncthreads();
    ex_cluster_sync();

    for (int i = i0; i < i1; i++) {
        const double* row = rows + (size_t)(i - row_base) * ld;
        ExBest best;
        best.add = 2.0; best.frac = 0.0; best.slot = 0; best.node = -1;   // the (0,0,2) default tuple
        double bpl = 0;          // edge length, parent and child index of this thread's best candidate
        int bx = -1, bc = 0;
        int st[EX_KM], pending = 0;
        double ra[EX_KM], rb[EX_KM], r0[EX_KM], r1[EX_KM];
        bool l0[EX_KM], l1[EX_KM];
#pragma unroll
        for (int u = 0; u < EX_KM; u++) {
            const int k = tid + u * EX_THREADS;
            st[u] = 0; ra[u] = empty; rb[u] = empty; r0[u] = 0; r1[u] = 0; l0[u] = false; l1[u] = false;
            if (k < NL && S.flag[k]) {
                const int k0 = S.kid0[k], k1 = S.kid1[k];
                l0[u] = k0 < N; l1[u] = k1 < N;
                if (l0[u]) { r0[u] = __ldg(&row[k0]); ra[u] = r0[u] - S.klen0[k]; }   // the leaf's limit towards this node (:298-332)
                if (l1[u]) { r1[u] = __ldg(&row[k1]); rb[u] = r1[u] - S.klen1[k]; }
                st[u] = 1; pending++;
            }
        }
        while (pending) {
#pragma unroll
            for (int u = 0; u < EX_KM; u++) {
                const int k = tid + u * EX_THREADS;
                if (st[u] == 1) {
                    if (!l0[u] && __double_as_longlong(ra[u]) == (long long)EX_EMPTY) {
                        const unsigned long long v = *reinterpret_cast<volatile unsigned long long*>(S.in0 + k);
                        if (v != EX_EMPTY) { ra[u] = __longlong_as_double((long long)v); *reinterpret_cast<volatile unsigned long long*>(S.in0 + k) = EX_EMPTY; }
                    }
                    if (!l1[u] && __double_as_longlong(rb[u]) == (long long)EX_EMPTY) {
                        const unsigned long long v = *reinterpret_cast<volatile unsigned long long*>(S.in1 + k);
                        if (v != EX_EMPTY) { rb[u] = __longlong_as_double((long long)v); *reinterpret_cast<volatile unsigned long long*>(S.in1 + k) = EX_EMPTY; }
                    }
                    if (__double_as_longlong(ra[u]) != (long long)EX_EMPTY && __double_as_longlong(rb[u]) != (long long)EX_EMPTY) {
                        if (S.flag[k] == 2) st[u] = 3;   // the root has nothing above it: feed the children now (below)
                        else {
                            double m = 0;
                            if (ra[u] > m) m = ra[u];
                            if (rb[u] > m) m = rb[u];
                            const int j = S.par[k] - N;
                            ex_push_f64((S.cidx[k] ? S.in1 : S.in0) + loc_of(j), cta_of(j), m - S.plen[k]);
                            st[u] = 2;
                        }
                    }
                }
                if (st[u] == 2 || st[u] == 3) {
                    // top-down step (:334-366), kept short: every firing is executed by a whole warp for the one or two lanes
                    // that are ready, so the scoring is left to the dense pass below
                    unsigned long long ud = 0;
                    if (st[u] == 2) ud = *reinterpret_cast<volatile unsigned long long*>(S.dn + k);
                    if (st[u] == 3 || ud != EX_EMPTY) {
                        const double a = ra[u], b = rb[u];
                        double v0 = 0, v1 = 0;
                        if (b > v0) v0 = b;
                        if (a > v1) v1 = a;
                        if (st[u] == 2) {
                            const double basev = __longlong_as_double((long long)ud) - S.plen[k];
                            if (basev > v0) v0 = basev;
                            if (basev > v1) v1 = basev;
                        }
                        if (!l0[u]) { const int j = S.kid0[k] - N; ex_push_f64(S.dn + loc_of(j), cta_of(j), v0); }
                        if (!l1[u]) { const int j = S.kid1[k] - N; ex_push_f64(S.dn + loc_of(j), cta_of(j), v1); }
                        st[u] = st[u] == 3 ? 5 : 4;   // done; 4: the value from above is still in the slot
                        pending--;
                    }
                }
            }
            if (backoff_ns && pending) __nanosleep(backoff_ns);   // polling competes with the peers' incoming stores for this SM's shared memory
        }
        // ---- scoring (calculateBranchLength :153-198), all lanes at once: own parent->node slot and the leaf children's slots
#pragma unroll
        for (int u = 0; u < EX_KM; u++) {
            const int k = tid + u * EX_THREADS;
            if (st[u] >= 4) {
                const double a = ra[u], b = rb[u];
                double v0 = 0, v1 = 0;
                if (b > v0) v0 = b;
                if (a > v1) v1 = a;
                const int me = N + id_of(k);
                if (st[u] == 4) {
                    const double dn = S.dn[k], pl = S.plen[k];
                    S.dn[k] = empty;
                    double up = 0;
                    if (a > up) up = a;
                    if (b > up) up = b;
                    const int before = best.node;
                    ex_score(dn, up, pl, S.sdn[k], me, best);
                    if (best.node != before) { bpl = pl; bx = S.par[k]; bc = S.cidx[k]; }
                    const double basev = dn - pl;
                    if (basev > v0) v0 = basev;
                    if (basev > v1) v1 = basev;
                }
                if (l0[u]) {
                    const int before = best.node;
                    ex_score(v0, r0[u], S.klen0[k], S.ksdn0[k], S.kid0[k], best);
                    if (best.node != before) { bpl = S.klen0[k]; bx = me; bc = 0; }
                }
                if (l1[u]) {
                    const int before = best.node;
                    ex_score(v1, r1[u], S.klen1[k], S.ksdn1[k], S.kid1[k], best);
                    if (best.node != before) { bpl = S.klen1[k]; bx = me; bc = 1; }
                }
            }
        }
        // ---- first minimum over the cluster (thrust::min_element :657)
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
            const double oa = __shfl_xor_sync(0xffffffffu, best.add, s), of = __shfl_xor_sync(0xffffffffu, best.frac, s), op = __shfl_xor_sync(0xffffffffu, bpl, s);
            const int os = __shfl_xor_sync(0xffffffffu, best.slot, s), on = __shfl_xor_sync(0xffffffffu, best.node, s);
            const int ox = __shfl_xor_sync(0xffffffffu, bx, s), oc = __shfl_xor_sync(0xffffffffu, bc, s);
            if (ex_before(oa, os, best.add, best.slot)) { best.add = oa; best.frac = of; best.slot = os; best.node = on; bpl = op; bx = ox; bc = oc; }
        }
        if (lane == 0) { s_warp[wid] = best; s_wpl[wid] = bpl; s_wx[wid] = bx; s_wc[wid] = bc; }
        __syncthreads();
        if (wid == 0) {
            best = s_warp[lane]; bpl = s_wpl[lane]; bx = s_wx[lane]; bc = s_wc[lane];
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) {
                const double oa = __shfl_xor_sync(0xffffffffu, best.add, s), of = __shfl_xor_sync(0xffffffffu, best.frac, s), op = __shfl_xor_sync(0xffffffffu, bpl, s);
                const int os = __shfl_xor_sync(0xffffffffu, best.slot, s), on = __shfl_xor_sync(0xffffffffu, best.node, s);
                const int ox = __shfl_xor_sync(0xffffffffu, bx, s), oc = __shfl_xor_sync(0xffffffffu, bc, s);
                if (ex_before(oa, os, best.add, best.slot)) { best.add = oa; best.frac = of; best.slot = os; best.node = on; bpl = op; bx = ox; bc = oc; }
            }
            if (lane < CS) {
                ex_push_f64(&rec_d[i & 1][rank][0], lane, best.add); ex_push_f64(&rec_d[i & 1][rank][1], lane, best.frac); ex_push_f64(&rec_d[i & 1][rank][2], lane, bpl);
                ex_push_s32(&rec_i[i & 1][rank][0], lane, best.slot); ex_push_s32(&rec_i[i & 1][rank][1], lane, best.node);
                ex_push_s32(&rec_i[i & 1][rank][2], lane, bx); ex_push_s32(&rec_i[i & 1][rank][3], lane, bc);
            }
        }
        ex_cluster_sync();
        int w = 0;
        {
            double wa = rec_d[i & 1][0][0];
            int ws = rec_i[i & 1][0][0];
#pragma unroll
            for (int c = 1; c < CS; c++) {
                const double ca = rec_d[i & 1][c][0];
                const int cs = rec_i[i & 1][c][0];
                if (ex_before(ca, cs, wa, ws)) { wa = ca; ws = cs; w = c; }
            }
        }
        const double addLen = rec_d[i & 1][w][0], fracLen = rec_d[i & 1][w][1], pleny = rec_d[i & 1][w][2];
        const int slot = rec_i[i & 1][w][0], y = rec_i[i & 1][w][1], x = rec_i[i & 1][w][2], cidxy = rec_i[i & 1][w][3];
        if (y < 0) { failed = true; break; }   // uniform over the cluster
        // ---- split (updateTreeStructure :200-251): middle m between x and y, new leaf i below m
        const int m = i + N - 1, jm = i - 1, c0 = 4 * i - 4;
        if (tid == 0 && cta_of(x - N) == rank) (cidxy ? S.kid1 : S.kid0)[loc_of(x - N)] = m;
        if (tid == 32 && y >= N && cta_of(y - N) == rank) {
            const int k = loc_of(y - N);
            S.par[k] = m; S.cidx[k] = 0; S.plen[k] = pleny - fracLen; S.sdn[k] = c0 + 1;
        }
        if (tid == 64 && cta_of(jm) == rank) {
            const int k = loc_of(jm);
            S.par[k] = x; S.cidx[k] = (unsigned char)cidxy; S.kid0[k] = y; S.kid1[k] = i; S.plen[k] = fracLen; S.sdn[k] = slot;
            S.klen0[k] = pleny - fracLen; S.ksdn0[k] = c0 + 1; S.klen1[k] = addLen; S.ksdn1[k] = c0 + 3; S.flag[k] = 1;
        }
        if (tid == 128 && rank == (i & (CS - 1))) {
            const int xe = slot, ye = y < N ? (y < 2 ? y : 4 * y - 2) : 4 * (y - N);
            const int c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3;
            e[xe] = m; len[xe] = fracLen;
            e[ye] = m; len[ye] = pleny - fracLen;
            e[c0] = x; len[c0] = fracLen; nxt[c0] = -1; belong[c0] = m;
            e[c1] = y; len[c1] = pleny - fracLen; nxt[c1] = c0; belong[c1] = m;
            e[c2] = m; len[c2] = addLen; nxt[c2] = -1; belong[c2] = i; head[i] = c2;
            e[c3] = i; len[c3] = addLen; nxt[c3] = c1; belong[c3] = m; head[m] = c3;
        }
        __syncthreads();
    }

    __syncthreads();
    ex_cluster_sync();   // no CTA may leave while peers can still push into its shared memory
    {
        uint4* dst = saved + (size_t)rank * state_words;
        for (size_t w = tid; w < state_words; w += EX_THREADS) dst[w] = sm4[w];
    }
    if (rank == 0 && tid == 0) {
        if (failed) ctl->error = 1;
        ctl->cycles += (unsigned long long)(clock64() - t_begin);
    }
}

// ---- beyond one cluster's shared memory (more than 49 152 tips; DIPB_EXACT_GLOBAL=1 forces it): the same data flow over
// internal nodes with the state in global memory and the whole grid (cooperative launch, one 1024-thread CTA per SM).
// A value is still its own arrival flag -- volatile 8-byte stores / loads meet in L2, no fence -- node j belongs to
// thread j mod T, the per-tip scratch of a node (state, the two limits from below) lives in global memory next to it,
// and the one synchronisation per tip is a grid barrier around the argmin exchange.  A hop costs an L2 round trip
// (~2 us) instead of a shared-memory one, so this path is for sizes the cluster cannot hold, not a replacement.
struct G2State {
    double *dn, *in0, *in1, *plen, *klen0, *klen1, *ra, *rb;
    int *par, *kid0, *kid1, *sdn, *ksdn0, *ksdn1;
    unsigned char *cidx, *flag, *st;
};
struct G2Rec { double add, frac, pl; int slot, node, x, c; };

__device__ __forceinline__ unsigned long long g2_poll(const double* p) { return *reinterpret_cast<const volatile unsigned long long*>(p); }
__device__ __forceinline__ void g2_put(double* p, unsigned long long bits) { *reinterpret_cast<volatile unsigned long long*>(p) = bits; }
__device__ __forceinline__ void g2_push(double* p, double v) { g2_put(p, (unsigned long long)__double_as_longlong(v)); }

__global__ void __launch_bounds__(EX_THREADS, 1)
place_exact_global_kernel(int* __restrict__ head, int* __restrict__ e, int* __restrict__ nxt, int* __restrict__ belong,
                          double* __restrict__ len, const double* __restrict__ rows, size_t ld, int row_base, int i0, int i1, int N,
                          const double* __restrict__ d01, G2State S, G2Rec* __restrict__ part, unsigned int* __restrict__ bar,
                          ExCtl* __restrict__ ctl) {
    __shared__ ExBest s_warp[EX_THREADS / 32];
    __shared__ double s_wpl[EX_THREADS / 32];
    __shared__ int s_wx[EX_THREADS / 32], s_wc[EX_THREADS / 32];
    __shared__ G2Rec s_win;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int G = (int)gridDim.x, T = G * EX_THREADS, gtid = (int)blockIdx.x * EX_THREADS + tid;
    unsigned int gen = 0;
    if (i0 == 2) {
        for (int j = gtid; j < N; j += T) {
            g2_put(S.dn + j, EX_EMPTY); g2_put(S.in0 + j, EX_EMPTY); g2_put(S.in1 + j, EX_EMPTY);
            S.plen[j] = 0; S.klen0[j] = 0; S.klen1[j] = 0; S.par[j] = -1; S.kid0[j] = -1; S.kid1[j] = -1;
            S.sdn[j] = 0; S.ksdn0[j] = 0; S.ksdn1[j] = 0; S.cidx[j] = 0; S.flag[j] = 0; S.st[j] = 0;
        }
        if (gtid == 0) {   // buildInitialTree :253-295 (node 0 of the internal nodes is written by its owner, thread 0)
            const double d = d01[0];
            S.flag[0] = 2; S.kid0[0] = 0; S.kid1[0] = 1; S.klen0[0] = d / 2; S.klen1[0] = d / 2; S.ksdn0[0] = 2; S.ksdn1[0] = 3;
            e[0] = N; len[0] = d / 2; nxt[0] = -1; belong[0] = 0; head[0] = 0;
            e[1] = N; len[1] = d / 2; nxt[1] = -1; belong[1] = 1; head[1] = 1;
            e[2] = 0; len[2] = d / 2; nxt[2] = -1; belong[2] = N;
            e[3] = 1; len[3] = d / 2; nxt[3] = 2; belong[3] = N; head[N] = 3;
        }
    }
    pl_grid_barrier(bar, (unsigned int)G, gen);
    bool failed = false;
    const long long t_begin = clock64();
    for (int i = i0; i < i1; i++) {
        const double* row = rows + (size_t)(i - row_base) * ld;
        const int placed = i - 1;   // internal nodes 0 .. i-2
        int pending = 0;
        for (int j = gtid; j < placed; j += T) {
            const int k0 = S.kid0[j], k1 = S.kid1[j];
            S.ra[j] = k0 < N ? __ldg(&row[k0]) - S.klen0[j] : __longlong_as_double((long long)EX_EMPTY);
            S.rb[j] = k1 < N ? __ldg(&row[k1]) - S.klen1[j] : __longlong_as_double((long long)EX_EMPTY);
            S.st[j] = 1;
            pending++;
        }
        while (pending) {
            for (int j = gtid; j < placed; j += T) {
                int st = S.st[j];
                if (st == 0 || st >= 4) continue;
                double a = S.ra[j], b = S.rb[j];
                if (st == 1) {
                    if (__double_as_longlong(a) == (long long)EX_EMPTY) {
                        const unsigned long long v = g2_poll(S.in0 + j);
                        if (v != EX_EMPTY) { a = __longlong_as_double((long long)v); S.ra[j] = a; g2_put(S.in0 + j, EX_EMPTY); }
                    }
                    if (__double_as_longlong(b) == (long long)EX_EMPTY) {
                        const unsigned long long v = g2_poll(S.in1 + j);
                        if (v != EX_EMPTY) { b = __longlong_as_double((long long)v); S.rb[j] = b; g2_put(S.in1 + j, EX_EMPTY); }
                    }
                    if (__double_as_longlong(a) != (long long)EX_EMPTY && __double_as_longlong(b) != (long long)EX_EMPTY) {
                        if (S.flag[j] == 2) st = 3;
                        else {
                            double m = 0;
                            if (a > m) m = a;
                            if (b > m) m = b;
                            const int pj = S.par[j] - N;
                            g2_push((S.cidx[j] ? S.in1 : S.in0) + pj, m - S.plen[j]);
                            st = 2;
                        }
                        S.st[j] = (unsigned char)st;
                    }
                }
                if (st == 2 || st == 3) {
                    unsigned long long ud = 0;
                    if (st == 2) ud = g2_poll(S.dn + j);
                    if (st == 3 || ud != EX_EMPTY) {
                        double v0 = 0, v1 = 0;
                        if (b > v0) v0 = b;
                        if (a > v1) v1 = a;
                        if (st == 2) {
                            const double basev = __longlong_as_double((long long)ud) - S.plen[j];
                            if (basev > v0) v0 = basev;
                            if (basev > v1) v1 = basev;
                        }
                        const int k0 = S.kid0[j], k1 = S.kid1[j];
                        if (k0 >= N) g2_push(S.dn + (k0 - N), v0);
                        if (k1 >= N) g2_push(S.dn + (k1 - N), v1);
                        S.st[j] = (unsigned char)(st == 3 ? 5 : 4);
                        pending--;
                    }
                }
            }
        }
        // ---- scoring (calculateBranchLength :153-198): own parent->node slot and the leaf children's slots
        ExBest best;
        best.add = 2.0; best.frac = 0.0; best.slot = 0; best.node = -1;
        double bpl = 0;
        int bx = -1, bc = 0;
        for (int j = gtid; j < placed; j += T) {
            const int st = S.st[j];
            const double a = S.ra[j], b = S.rb[j];
            double v0 = 0, v1 = 0;
            if (b > v0) v0 = b;
            if (a > v1) v1 = a;
            const int me = N + j;
            if (st == 4) {
                const double dn = __longlong_as_double((long long)g2_poll(S.dn + j)), pl = S.plen[j];
                g2_put(S.dn + j, EX_EMPTY);
                double up = 0;
                if (a > up) up = a;
                if (b > up) up = b;
                const int before = best.node;
                ex_score(dn, up, pl, S.sdn[j], me, best);
                if (best.node != before) { bpl = pl; bx = S.par[j]; bc = S.cidx[j]; }
                const double basev = dn - pl;
                if (basev > v0) v0 = basev;
                if (basev > v1) v1 = basev;
            }
            const int k0 = S.kid0[j], k1 = S.kid1[j];
            if (k0 < N) {
                const int before = best.node;
                ex_score(v0, __ldg(&row[k0]), S.klen0[j], S.ksdn0[j], k0, best);
                if (best.node != before) { bpl = S.klen0[j]; bx = me; bc = 0; }
            }
            if (k1 < N) {
                const int before = best.node;
                ex_score(v1, __ldg(&row[k1]), S.klen1[j], S.ksdn1[j], k1, best);
                if (best.node != before) { bpl = S.klen1[j]; bx = me; bc = 1; }
            }
        }
        // ---- first minimum over the grid (thrust::min_element :657)
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
            const double oa = __shfl_xor_sync(0xffffffffu, best.add, s), of = __shfl_xor_sync(0xffffffffu, best.frac, s), op = __shfl_xor_sync(0xffffffffu, bpl, s);
            const int os = __shfl_xor_sync(0xffffffffu, best.slot, s), on = __shfl_xor_sync(0xffffffffu, best.node, s);
            const int ox = __shfl_xor_sync(0xffffffffu, bx, s), oc = __shfl_xor_sync(0xffffffffu, bc, s);
            if (ex_before(oa, os, best.add, best.slot)) { best.add = oa; best.frac = of; best.slot = os; best.node = on; bpl = op; bx = ox; bc = oc; }
        }
        if (lane == 0) { s_warp[wid] = best; s_wpl[wid] = bpl; s_wx[wid] = bx; s_wc[wid] = bc; }
        __syncthreads();
        if (wid == 0) {
            best = s_warp[lane]; bpl = s_wpl[lane]; bx = s_wx[lane]; bc = s_wc[lane];
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) {
                const double oa = __shfl_xor_sync(0xffffffffu, best.add, s), of = __shfl_xor_sync(0xffffffffu, best.frac, s), op = __shfl_xor_sync(0xffffffffu, bpl, s);
                const int os = __shfl_xor_sync(0xffffffffu, best.slot, s), on = __shfl_xor_sync(0xffffffffu, best.node, s);
                const int ox = __shfl_xor_sync(0xffffffffu, bx, s), oc = __shfl_xor_sync(0xffffffffu, bc, s);
                if (ex_before(oa, os, best.add, best.slot)) { best.add = oa; best.frac = of; best.slot = os; best.node = on; bpl = op; bx = ox; bc = oc; }
            }
            if (lane == 0) {
                G2Rec r; r.add = best.add; r.frac = best.frac; r.pl = bpl; r.slot = best.slot; r.node = best.node; r.x = bx; r.c = bc;
                part[(size_t)(i & 1) * G + blockIdx.x] = r;
            }
        }
        pl_grid_barrier(bar, (unsigned int)G, gen);
        if (wid == 0) {
            G2Rec w; w.add = 2.0; w.frac = 0; w.pl = 0; w.slot = 0; w.node = -1; w.x = -1; w.c = 0;
            for (int c = lane; c < G; c += 32) {
                const G2Rec o = part[(size_t)(i & 1) * G + c];
                if (ex_before(o.add, o.slot, w.add, w.slot)) w = o;
            }
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) {
                G2Rec o;
                o.add = __shfl_xor_sync(0xffffffffu, w.add, s); o.frac = __shfl_xor_sync(0xffffffffu, w.frac, s); o.pl = __shfl_xor_sync(0xffffffffu, w.pl, s);
                o.slot = __shfl_xor_sync(0xffffffffu, w.slot, s); o.node = __shfl_xor_sync(0xffffffffu, w.node, s);
                o.x = __shfl_xor_sync(0xffffffffu, w.x, s); o.c = __shfl_xor_sync(0xffffffffu, w.c, s);
                if (ex_before(o.add, o.slot, w.add, w.slot)) w = o;
            }
            if (lane == 0) s_win = w;
        }
        __syncthreads();
        const double addLen = s_win.add, fracLen = s_win.frac, pleny = s_win.pl;
        const int slot = s_win.slot, y = s_win.node, x = s_win.x, cidxy = s_win.c;
        __syncthreads();   // s_win is rewritten in the next tip
        if (y < 0) { failed = true; break; }
        // ---- split (updateTreeStructure :200-251), every node by its owner thread
        const int m = i + N - 1, jm = i - 1, c0 = 4 * i - 4;
        if ((x - N) % T == gtid) (cidxy ? S.kid1 : S.kid0)[x - N] = m;
        if (y >= N && (y - N) % T == gtid) { const int k = y - N; S.par[k] = m; S.cidx[k] = 0; S.plen[k] = pleny - fracLen; S.sdn[k] = c0 + 1; }
        if (jm % T == gtid) {
            S.par[jm] = x; S.cidx[jm] = (unsigned char)cidxy; S.kid0[jm] = y; S.kid1[jm] = i; S.plen[jm] = fracLen; S.sdn[jm] = slot;
            S.klen0[jm] = pleny - fracLen; S.ksdn0[jm] = c0 + 1; S.klen1[jm] = addLen; S.ksdn1[jm] = c0 + 3; S.flag[jm] = 1;
            const int xe = slot, ye = y < N ? (y < 2 ? y : 4 * y - 2) : 4 * (y - N);
            const int c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3;
            e[xe] = m; len[xe] = fracLen;
            e[ye] = m; len[ye] = pleny - fracLen;
            e[c0] = x; len[c0] = fracLen; nxt[c0] = -1; belong[c0] = m;
            e[c1] = y; len[c1] = pleny - fracLen; nxt[c1] = c0; belong[c1] = m;
            e[c2] = m; len[c2] = addLen; nxt[c2] = -1; belong[c2] = i; head[i] = c2;
            e[c3] = i; len[c3] = addLen; nxt[c3] = c1; belong[c3] = m; head[m] = c3;
        }
    }
    if (gtid == 0) {
        if (failed) ctl->error = 1;
        ctl->cycles += (unsigned long long)(clock64() - t_begin);
    }
}

template <int CS, int MODE>   // MODE 0: level steps, 1: data flow over all nodes, 2: data flow over internal nodes (default)
int ex_launch(dipb_ctx* c, void** args, size_t smem, bool* ok) {
    auto kern = MODE == 2 ? place_exact_flow2_kernel<CS> : (MODE == 1 ? place_exact_kernel<CS, true> : place_exact_kernel<CS, false>);
    *ok = false;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (CS > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS); cfg.blockDim = dim3(EX_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess || nclusters < 1) { cudaGetLastError(); return 0; }
    cudaError_t e = cudaLaunchKernelExC(&cfg, (const void*)kern, args);
    if (e != cudaSuccess) { set_error("exact placement: launch failed: %s", cudaGetErrorString(e)); return DIPB_E_CUDA; }
    *ok = true;
    return 0;
}

// local node capacity per CTA for n tips on CS CTAs, and the dynamic shared memory it needs
inline int ex_local(int n, int CS) { const int chunks = (n + 31) / 32; return ((chunks + CS - 1) / CS) * 32; }
inline size_t ex_smem_bytes(int NL, int mode) { return (size_t)NL * (mode == 2 ? F2_BYTES_PER_NODE : EX_BYTES_PER_NODE); }
constexpr size_t EX_SMEM_MAX = 224u * 1024u;   // 227 KB minus the static records

}  // namespace

static int ex_max_tips(int mode) {
    int nl = (int)(EX_SMEM_MAX / (mode == 2 ? F2_BYTES_PER_NODE : EX_BYTES_PER_NODE) / 32) * 32;
    if (nl > EX_KM * EX_THREADS) nl = EX_KM * EX_THREADS;
    return nl * 16;
}

// tips beyond the cluster's shared memory: state in global memory, whole grid (see place_exact_global_kernel)
static int place_exact_run_global(dipb_ctx* c, const dipb_dist_source* src, int n, dipb_tree* t) {
    const double* d01 = nullptr;
    double* row1 = nullptr;
    int rc = 0;
    if (src->matrix) d01 = src->matrix->d + (size_t)src->matrix->n;
    else {
        DIPB_CUDA(pool_alloc(c, (void**)&row1, sizeof(double) * 8));
        rc = src->msa ? msa_block(src->msa, src->dist_type, 1, 2, 1, row1, 8) : dipb_mash_dist_block(src->mash, 1, 2, 1, row1, 8);
        if (rc) { pool_free(c, row1); return rc; }
        d01 = row1;
    }
    int batch = 512;
    double* buf = nullptr;
    const size_t ld = (size_t)n, N = (size_t)n;
    if (!src->matrix) {
        size_t want = (size_t)batch * ld * sizeof(double);
        while (want > (1ull << 28) && batch > 128) { batch /= 2; want /= 2; }
        DIPB_CUDA(pool_alloc(c, (void**)&buf, (size_t)batch * ld * sizeof(double)));
    } else batch = n;
    G2State S{};
    double* dblk = nullptr;
    int* iblk = nullptr;
    unsigned char* bblk = nullptr;
    DIPB_CUDA(pool_alloc(c, (void**)&dblk, 8 * N * sizeof(double)));
    DIPB_CUDA(pool_alloc(c, (void**)&iblk, 6 * N * sizeof(int)));
    DIPB_CUDA(pool_alloc(c, (void**)&bblk, 3 * N));
    S.dn = dblk; S.in0 = dblk + N; S.in1 = dblk + 2 * N; S.plen = dblk + 3 * N; S.klen0 = dblk + 4 * N; S.klen1 = dblk + 5 * N; S.ra = dblk + 6 * N; S.rb = dblk + 7 * N;
    S.par = iblk; S.kid0 = iblk + N; S.kid1 = iblk + 2 * N; S.sdn = iblk + 3 * N; S.ksdn0 = iblk + 4 * N; S.ksdn1 = iblk + 5 * N;
    S.cidx = bblk; S.flag = bblk + N; S.st = bblk + 2 * N;
    int per_sm = 0;
    DIPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, place_exact_global_kernel, EX_THREADS, 0));
    if (per_sm < 1) { set_error("exact placement (global): kernel does not fit"); return DIPB_E_CUDA; }
    int G = c->num_sms;
    G2Rec* part = nullptr;
    unsigned int* bar = nullptr;
    ExCtl* ctl = nullptr;
    DIPB_CUDA(pool_alloc(c, (void**)&part, 2 * (size_t)G * sizeof(G2Rec)));
    DIPB_CUDA(pool_alloc(c, (void**)&bar, sizeof(unsigned int)));
    DIPB_CUDA(pool_alloc(c, (void**)&ctl, sizeof(ExCtl)));
    DIPB_CUDA(cudaMemsetAsync(ctl, 0, sizeof(ExCtl), c->stream));
    for (int i0 = 2; (i0 < n || i0 == 2) && !rc; i0 += batch) {
        int i1 = i0 + batch < n ? i0 + batch : n;
        const double* rows = nullptr; int row_base = 0; size_t ldr = ld;
        if (i1 > i0) rc = place_fetch_rows(src, i0, i1, buf, ld, &rows, &row_base, &ldr);
        if (rc) break;
        DIPB_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned int), c->stream));
        int Nn = n, i0v = i0;
        void* args[] = {&t->head, &t->e, &t->nxt, &t->belong, &t->len, &rows, &ldr, &row_base, &i0v, &i1, &Nn, &d01, &S, &part, &bar, &ctl};
        cudaError_t e = cudaLaunchCooperativeKernel((void*)place_exact_global_kernel, dim3(G), dim3(EX_THREADS), args, 0, c->stream);
        if (e != cudaSuccess) { set_error("exact placement (global): cooperative launch failed: %s", cudaGetErrorString(e)); rc = DIPB_E_CUDA; break; }
        c->launches++;
    }
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (!rc && e != cudaSuccess) { set_error("exact placement (global): %s", cudaGetErrorString(e)); rc = DIPB_E_CUDA; }
    if (!rc) {
        ExCtl h;
        if (cudaMemcpy(&h, ctl, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("exact placement: control block copy failed"); rc = DIPB_E_CUDA; }
        else if (h.error) {
            set_error("exact placement: a tip has no candidate edge with pendant length < 2 (the reference's (0,0,2) default tuple would win; src/placement.cu:166-170)");
            rc = DIPB_E_UNSUPPORTED;
        } else if (getenv("DIPB_PLACE_PROFILE"))
            fprintf(stderr, "[exact placement] %d tips, global-memory data flow on %d CTAs: %.0f cycles per tip\n", n, G, (double)h.cycles / (n - 2));
    }
    pool_free(c, dblk); pool_free(c, iblk); pool_free(c, bblk); pool_free(c, part); pool_free(c, bar); pool_free(c, ctl);
    if (buf) pool_free(c, buf);
    if (row1) pool_free(c, row1);
    return rc;
}

int place_exact_run(dipb_ctx* c, const dipb_dist_source* src, int n, dipb_tree* t) {
    int CS = 16;
    const char* force = getenv("DIPB_EXACT_CLUSTER");
    if (force && atoi(force) == 8) CS = 8;
    const char* fl = getenv("DIPB_EXACT_FLOW");   // 0: level steps with cluster barriers, 1: data flow over all nodes (kept for comparison)
    const int mode = fl ? (atoi(fl) == 0 ? 0 : (atoi(fl) == 1 ? 1 : 2)) : 2;
    const char* eb = getenv("DIPB_EXACT_BACKOFF");   // ns of __nanosleep between polling passes (default mode only)
    int backoff = eb ? atoi(eb) : 0;
    int NL = ex_local(n, CS);
    const char* eg = getenv("DIPB_EXACT_GLOBAL");   // 1: always the global-memory kernel; 0: never (report the size limit instead)
    const bool too_big = ex_smem_bytes(NL, mode) > EX_SMEM_MAX || NL > EX_KM * EX_THREADS;
    if ((eg && atoi(eg) == 1) || (too_big && !(eg && atoi(eg) == 0))) return place_exact_run_global(c, src, n, t);
    if (too_big) {
        set_error("exact placement: %d tips exceed the shared-memory tree of one %d-CTA cluster (at most %d tips); use -p 1 or -m 3", n, CS,
                  ex_max_tips(mode) / (16 / CS));
        return DIPB_E_UNSUPPORTED;
    }
    // d(1,0) for the 2-leaf tree
    const double* d01 = nullptr;
    double* row1 = nullptr;
    int rc = 0;
    if (src->matrix) d01 = src->matrix->d + (size_t)src->matrix->n;
    else {
        DIPB_CUDA(pool_alloc(c, (void**)&row1, sizeof(double) * 8));
        rc = src->msa ? msa_block(src->msa, src->dist_type, 1, 2, 1, row1, 8) : dipb_mash_dist_block(src->mash, 1, 2, 1, row1, 8);
        if (rc) { pool_free(c, row1); return rc; }
        d01 = row1;
    }
    int batch = 512;
    double* buf = nullptr;
    const size_t ld = (size_t)n;
    if (!src->matrix) {
        size_t want = (size_t)batch * ld * sizeof(double);
        while (want > (1ull << 28) && batch > 128) { batch /= 2; want /= 2; }
        DIPB_CUDA(pool_alloc(c, (void**)&buf, (size_t)batch * ld * sizeof(double)));
    } else batch = n;   // all rows are there: one launch
    uint4* saved = nullptr;
    ExCtl* ctl = nullptr;
    DIPB_CUDA(pool_alloc(c, (void**)&saved, ex_smem_bytes(NL, mode) * 16));
    DIPB_CUDA(pool_alloc(c, (void**)&ctl, sizeof(ExCtl)));
    DIPB_CUDA(cudaMemsetAsync(ctl, 0, sizeof(ExCtl), c->stream));
    for (int i0 = 2; (i0 < n || i0 == 2) && !rc; i0 += batch) {   // (n == 2: one launch that only builds the 2-leaf tree)
        int i1 = i0 + batch < n ? i0 + batch : n;
        const double* rows = nullptr; int row_base = 0; size_t ldr = ld;
        if (i1 > i0) rc = place_fetch_rows(src, i0, i1, buf, ld, &rows, &row_base, &ldr);
        if (rc) break;
        int N = n;
        int i0v = i0;
        void* args[] = {&t->head, &t->e, &t->nxt, &t->belong, &t->len, &rows, &ldr, &row_base, &i0v, &i1, &N, &NL, &d01, &saved, &ctl, &backoff};
        bool ok = false;
        if (CS == 16) {
            rc = mode == 2 ? ex_launch<16, 2>(c, args, ex_smem_bytes(NL, mode), &ok) : (mode == 1 ? ex_launch<16, 1>(c, args, ex_smem_bytes(NL, mode), &ok) : ex_launch<16, 0>(c, args, ex_smem_bytes(NL, mode), &ok));
            if (!rc && !ok && i0 == 2) {   // device cannot co-schedule 16 CTAs: portable cluster size, half the capacity
                CS = 8; NL = ex_local(n, 8);
                if (ex_smem_bytes(NL, mode) > EX_SMEM_MAX || NL > EX_KM * EX_THREADS) { set_error("exact placement: 16-CTA clusters unavailable and %d tips do not fit 8 CTAs", n); rc = DIPB_E_UNSUPPORTED; break; }
                pool_free(c, saved);
                DIPB_CUDA(pool_alloc(c, (void**)&saved, ex_smem_bytes(NL, mode) * 8));
            }
        }
        if (!rc && !ok && CS == 8) rc = mode == 2 ? ex_launch<8, 2>(c, args, ex_smem_bytes(NL, mode), &ok) : (mode == 1 ? ex_launch<8, 1>(c, args, ex_smem_bytes(NL, mode), &ok) : ex_launch<8, false>(c, args, ex_smem_bytes(NL, mode), &ok));
        if (!rc && !ok) { set_error("exact placement: no cluster configuration fits this device"); rc = DIPB_E_UNSUPPORTED; }
        if (!rc) c->launches++;
    }
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (!rc && e != cudaSuccess) { set_error("exact placement: %s", cudaGetErrorString(e)); rc = DIPB_E_CUDA; }
    if (!rc) {
        ExCtl h;
        if (cudaMemcpy(&h, ctl, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("exact placement: control block copy failed"); rc = DIPB_E_CUDA; }
        else if (h.error) {
            set_error("exact placement: a tip has no candidate edge with pendant length < 2 (the reference's (0,0,2) default tuple would win; src/placement.cu:166-170)");
            rc = DIPB_E_UNSUPPORTED;
        } else if (getenv("DIPB_PLACE_PROFILE"))
            fprintf(stderr, "[exact placement] %d tips on a %d-CTA cluster: final depth %d, %.1f levels and %.0f cycles per tip\n", n, CS, h.maxdep,
                    (double)h.levels / (n - 2), (double)h.cycles / (n - 2));
    }
    pool_free(c, saved); pool_free(c, ctl);
    if (buf) pool_free(c, buf);
    if (row1) pool_free(c, row1);
    return rc;
}

}  // namespace dipb

using namespace dipb;

extern "C" int dipb_place_exact_max_tips(void) { return ex_max_tips(2); }

extern "C" int dipb_place_exact(dipb_ctx* c, const dipb_dist_source* src, int n, dipb_tree** out) {
    if (!c || !src || !out || n < 2) { set_error("dipb_place_exact: bad argument"); return DIPB_E_ARG; }
    int rc = check_source(src, n);
    if (rc) return rc;
    DIPB_CUDA(cudaSetDevice(c->device));
    rc = timer_begin(c);
    if (rc) return rc;
    dipb_tree* t = nullptr;
    rc = tree_alloc(c, n, &t);
    if (rc) return rc;
    rc = place_exact_run(c, src, n, t);
    if (rc) { dipb_tree_free(t); return rc; }
    rc = timer_end(c, DIPB_T_PLACE);
    if (rc) return rc;
    *out = t;
    return 0;
}
