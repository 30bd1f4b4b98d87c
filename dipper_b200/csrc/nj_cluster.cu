// K4 (cluster): the bound-pruned exact NJ search of nj_pruned.cu, run by ONE thread-block cluster.
//
// Why one cluster and not the whole GPU.  After pruning an NJ iteration touches little data (~34 rows
// rescanned + 5 row/column updates: ~8 MB at 30 000 tips, shrinking linearly), but it is a chain of four
// dependent steps.  Spread over 148 CTAs each step ends in a software grid barrier through L2 (~2 us) and
// every exchanged scalar is another L2 round trip: 25 us per iteration, the same at 4 000 tips as at
// 30 000 (profiles/r1_nj_phase_cycles.txt).  A 16-CTA cluster synchronises in hardware
// (barrier.cluster, ~0.2 us), exchanges scalars through distributed shared memory, and keeps the per-row
// state (U, u, the new column) in shared memory; 16 SMs still pull ~1.5 TB/s, enough for the scan.
//
// Same algorithm and result as nj_pruned.cu (see its header for the bound): replaces the loop of
// NJDeviceArrays::findNeighbourJoiningTree (src/neighborJoining.cu:196-246) with findMinDist (:117-148),
// thrust::min_element (:214) and updateDisMatrix (:161-194), tie order and U summation order included.
//
// Layout.  Rows are dealt to the CTAs in chunks of 32: chunk w = i / 32 belongs to CTA w % CS, local slot
// (w / CS) * 32 + i % 32.  A CTA owns U, u, the folded new-column value of its rows, and scans its own
// column chunks of every selected row (u of those columns is local).  K (row lower-bound keys) and D live in
// global memory, read with ld.global.cg.
//
// Iteration (4 cluster barriers):
//   D  every CTA reduces the 16 published CTA winners to the same (x, y); rank 0 logs the merge
//   A  owners update rows/columns x, y (move `last` into y), U, u, chunk sums, max u-drift      | barrier
//   B1 U[x] (canonical sum order), C += drift; each CTA re-evaluates its carried candidate pairs  | barrier
//   B2 ub = min over CTAs; owners fold the new column into K and select rows with lb <= ub        | barrier
//   C  every CTA scans its column chunks of the selected rows; publishes its winner to all CTAs   | barrier
#include <cooperative_groups.h>
#include <cstdlib>
#include <vector>
#include "common.cuh"
#include "nj.cuh"
#include "nj_bound.cuh"

namespace cg = cooperative_groups;

namespace dipb {

namespace {

constexpr int MAXW = 32;      // warps per CTA at most
constexpr int UC = 8;         // column chunks per scan unit (loads in flight per lane)
constexpr int CPOOL = 128;    // carried candidate pairs per CTA
constexpr int MAXCS = 16;

struct CRec {                 // a CTA's best candidate of one scan
    double t, d, ui, uj;
    int i, j;
};

struct NJCtl {                // main cluster -> helper clusters doorbell (global memory)
    unsigned int seq;         // number of merges published; 0xffffffff = quit
    int x, y, n;              // the published merge: new node x, slot y (received the old last row when y < n)
    unsigned int done;        // helper CTAs that finished, cumulative
    unsigned int pad[3];
};

struct CStats {
    unsigned long long rows_scanned, iters;
    unsigned long long cyc[24];
};

// lexicographic warp minimum of a u64 through two 32-bit redux ops
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
    const unsigned int hi = (unsigned int)(v >> 32), lo = (unsigned int)v;
    const unsigned int mh = __reduce_min_sync(0xffffffffu, hi);
    const unsigned int ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    return ((unsigned long long)mh << 32) | ml;
}

// distributed-shared-memory load: the same variable in CTA `rank` of the cluster (mapa + ld.shared::cluster)
__device__ __forceinline__ double ld_peer_f64(const double* p, int rank) {
    const unsigned int a = (unsigned int)__cvta_generic_to_shared(p);
    unsigned int ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
    return v;
}

__device__ __forceinline__ double warp_min_f64(double v) { return dec_f64(warp_min_u64(enc_f64(v))); }
__device__ __forceinline__ double warp_max_f64(double v) { return -warp_min_f64(-v); }

// ---- candidate order (reference scan order, nj_bound.cuh) kept OUT of line: exact ties are rare, and the
// kernel must stay small -- every warp walks the whole iteration body once per merge, so a body that
// overflows the 32 KB instruction cache pays an L2 fetch every few instructions.
__device__ __noinline__ bool tie_before(int ia, int ja, int ib, int jb, int n) {
    if (ia == ib) return ((ja & 255) < (jb & 255)) || ((ja & 255) == (jb & 255) && ja < jb);
    return p_tie_key(ia, ja, n) < p_tie_key(ib, jb, n);
}
__device__ __noinline__ int tie_lane(unsigned int tied, int i, int j, int n) {
    const bool in = (tied >> (threadIdx.x & 31)) & 1u;
    const unsigned long long key = in ? p_tie_key(i, j, n) : 0xffffffffffffffffull;
    const unsigned long long km = warp_min_u64(key);
    return __ffs(__ballot_sync(0xffffffffu, key == km)) - 1;
}
// lane holding the best candidate of the warp (t ascending, then reference order), -1 when no lane has one
__device__ __forceinline__ int warp_best_lane(double t, int i, int j, int n) {
    const unsigned long long e = i >= 0 ? enc_f64(t) : 0xffffffffffffffffull;
    const unsigned long long m = warp_min_u64(e);
    if (m == 0xffffffffffffffffull) return -1;
    const unsigned int tied = __ballot_sync(0xffffffffu, e == m);
    if ((tied & (tied - 1u)) == 0u) return __ffs(tied) - 1;
    return tie_lane(tied, i, j, n);
}

}  // namespace

template <int CS, int CT, bool PROF>
__global__ void __launch_bounds__(CT, 1)
nj_cluster_kernel(double* __restrict__ D, size_t ld, const double* __restrict__ U0, const double* __restrict__ u0,
                  unsigned long long* __restrict__ K, int* __restrict__ sel_rows, CStats* stats,
                  int2* __restrict__ log_xy, double2* __restrict__ log_bl, int n_total, int LS, double dmax,
                  NJCtl* ctl, int HC) {
    if (blockIdx.x >= CS) {
        // ---- helper clusters (the other GPCs): transpose rows x and y of each published merge into columns x and
        // y.  These 2n scattered 8-byte stores per merge are request-rate bound on one GPC's L2 port when the
        // main cluster issues them itself (1 us per 1000 tips); spread over the other GPCs they are off the
        // critical path.
        const int hc = (int)blockIdx.x - CS, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
        __shared__ int s_msg[4];
        unsigned int seen = 0;
        for (;;) {
            if (tid == 0) {
                unsigned int q;
                for (;;) {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(q) : "l"(&ctl->seq) : "memory");
                    if (q != seen) break;
                    __nanosleep(64);
                }
                s_msg[0] = (int)q; s_msg[1] = __ldcg(&ctl->x); s_msg[2] = __ldcg(&ctl->y); s_msg[3] = __ldcg(&ctl->n);
            }
            __syncthreads();
            const unsigned int q = (unsigned int)s_msg[0];
            const int x = s_msg[1], y = s_msg[2], n = s_msg[3];
            __syncthreads();
            if (q == 0xffffffffu) return;
            seen = q;
            const int nch = (n + 31) >> 5;
            for (int c = hc + w * HC; c < nch; c += HC * (CT / 32)) {
                const int i = c * 32 + lane;
                if (i < n && i != x && i != y) {
                    D[(size_t)i * ld + x] = __ldcg(&D[(size_t)x * ld + i]);
                    if (y < n) D[(size_t)i * ld + y] = __ldcg(&D[(size_t)y * ld + i]);
                }
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) atomicAdd(&ctl->done, 1u);
        }
    }
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    constexpr int NW = CT / 32;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* U_s = reinterpret_cast<double*>(smem_raw);   // [LS] row sums of owned rows
    double* u_s = U_s + LS;                              // [LS] U / (n - 2)
    double* v_s = u_s + LS;                              // [LS] distance of owned rows to the newest node x
    double* f_s = v_s + LS;                              // [LS] distance of owned rows to the node moved into slot y
    double* cs_s = f_s + LS;                             // [LS / 32] chunk sums of the new column (canonical order)
    // (+ UC * 32 doubles of padding: the scan reads u_s up to UC - 1 chunks past the last owned one)

    __shared__ CRec recs[MAXCS];            // winners published by every CTA of the cluster
    __shared__ CRec wrec[MAXW];
    __shared__ double s_drift, s_ubmin;     // this CTA's max u-drift / best carried candidate (read by peers)
    __shared__ double s_red[MAXW], s_blk[128 + 16];   // s_blk: one sum per 1024-row block (n <= 131 072)
    __shared__ double s_total, s_C, s_ub;
    __shared__ unsigned int s_sel;          // selected-row counter (rank 0's copy is the live one)
    __shared__ int pool_i[CPOOL], pool_j[CPOOL];
    __shared__ double pool_d[CPOOL];        // d of a carried pair never changes while both ends survive
    __shared__ int s_pool_head, s_nsel;
    __shared__ unsigned long long s_cyc[24];   // rank 0, thread 0: cycles per phase (DIPB_NJ_PROFILE)


    // ---- load owned state
    for (int s = tid; s < LS; s += CT) {
        const int i = ((s >> 5) * CS + rank) * 32 + (s & 31);
        U_s[s] = i < n_total ? U0[i] : 0.0;
        u_s[s] = i < n_total ? u0[i] : 0.0;
        v_s[s] = 0.0;
    }
    for (int p = tid; p < CPOOL; p += CT) { pool_i[p] = -1; pool_j[p] = -1; }
    if (tid == 0) { s_pool_head = 0; s_sel = 0; s_drift = -1e300; s_ubmin = 1e300; }
    if (tid < 24) s_cyc[tid] = 0;
    __syncthreads();
    cluster.sync();

    int n = n_total;
    double C = 0.0;
    int x = -1, y = -1;
    double dxy = 0.0;
    bool first = true;
    int iter = 0;
    unsigned long long my_rows = 0;
    unsigned int* sel0 = cluster.map_shared_rank(&s_sel, 0);

    long long tmark = clock64();
#define CL_MARK(k)                                                      \
    do {                                                                \
        if (PROF && rank == 0 && tid == 0) {                            \
            long long now__ = clock64();                                \
            s_cyc[k] += (unsigned long long)(now__ - tmark);            \
            tmark = now__;                                              \
        }                                                               \
    } while (0)

    while (n > 2) {
        double ub = 1e300;
        if (!first) {
            // ------------------------------------------------------------ A: merge update by row owners
            CL_MARK(1);
            const int last = n - 1;
            const double den_new = (double)(n - 3);
            const int nchunk = (last + 31) >> 5;               // chunks holding rows < last
            double dmx = -1e300;
            for (int lw = w; lw * CS + rank < nchunk; lw += NW) {
                const int i = (lw * CS + rank) * 32 + lane;
                const int s = lw * 32 + lane;
                double slot = 0.0;
                if (i < last && i != x) {
                    // slot y receives the node that lived in row `last`: same formulas, sources taken from `last`
                    const bool isy = (i == y);
                    const int src = isy ? last : i;
                    const double a = __ldcg(&D[(size_t)x * ld + src]), b = __ldcg(&D[(size_t)y * ld + src]);
                    const double far = __ldcg(&D[(size_t)last * ld + i]);
                    double Ui = U_s[s], uo = u_s[s];
                    if (isy) {
                        const int lo = (last >> 5) % CS, ls = ((last >> 5) / CS) * 32 + (last & 31);
                        Ui = ld_peer_f64(&U_s[ls], lo);
                        uo = ld_peer_f64(&u_s[ls], lo);
                        K[y] = __ldcg(&K[last]);
                    }
                    const double val = (a + b - dxy) * 0.5;
                    Ui += -a - b + val;
                    U_s[s] = Ui;
                    D[(size_t)x * ld + i] = val;         // rows x and y: coalesced, visible after the next barrier
                    if (isy) D[(size_t)y * ld + x] = val;
                    else D[(size_t)y * ld + i] = far;
                    f_s[s] = far;                        // columns x and y of row i are written during phase C
                    slot = val;
                    if (n > 3) {
                        const double un = Ui / den_new;
                        dmx = fmax(dmx, un - uo);
                        u_s[s] = un;
                    }
                }
                v_s[s] = slot;
                const double csum = warp_tree_sum(slot);
                if (lane == 0) cs_s[lw] = csum;
            }
            CL_MARK(2);
            dmx = warp_max_f64(dmx);
            if (lane == 0) s_red[w] = dmx;
            __syncthreads();
            if (w == 0) {
                const double m = warp_max_f64(lane < NW ? s_red[lane] : -1e300);
                if (lane == 0) s_drift = m;
            }
            CL_MARK(3);
            cluster.sync();
            CL_MARK(4);

            // ------------------------------------------------------------ B1: U[x], drift, carried candidates
            n = last;
            if (n <= 2) {
                // last merge: its column writes are not deferred (nj_finish_kernel reads D[0][1])
                if (rank == 0 && tid < n && tid != x && tid != y) D[(size_t)tid * ld + x] = v_s[tid];
                break;
            }
            if (HC > 0 && rank == 0 && tid == 0) {
                // rows x and y are complete and fenced (every thread ran MEMBAR.GPU before the barrier): ring the helpers
                ctl->x = x; ctl->y = y; ctl->n = n;
                __threadfence();
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&ctl->seq), "r"((unsigned int)iter) : "memory");
            }
            {
                // canonical sum: 1024-row blocks (32 chunk sums, stride-halving tree), blocks ascending
                const int nblk = (last + 1023) >> 10;
                for (int b = w; b < nblk; b += NW) {
                    const int cw = b * 32 + lane;
                    double v = 0.0;
                    if (cw < nchunk) v = ld_peer_f64(&cs_s[cw / CS], cw % CS);
                    v = warp_tree_sum(v);
                    if (lane == 0) s_blk[b] = v;
                }
                double drift = -1e300;
                if (tid < CS) drift = ld_peer_f64(&s_drift, tid);
                if (w == 0) drift = warp_max_f64(drift);
                __syncthreads();
                if (tid == 0) {
                    double acc = 0.0;
                    for (int b0 = 0; b0 < nblk; b0 += 16) {          // ascending order; loads issued ahead of the add chain
                        double v[16];
#pragma unroll
                        for (int q = 0; q < 16; q++) v[q] = s_blk[b0 + q];
#pragma unroll
                        for (int q = 0; q < 16; q++) if (b0 + q < nblk) acc += v[q];
                    }
                    s_total = acc;
                    s_C = C + drift;
                }
                __syncthreads();
            }
            CL_MARK(5);
            const double total = s_total;
            const double ux = total / (double)(n - 2);
            C = s_C;
            if (((x >> 5) % CS) == rank && tid == 0) {
                const int sx = ((x >> 5) / CS) * 32 + (x & 31);
                U_s[sx] = total;
                u_s[sx] = ux;
            }
            // carried candidates of this CTA, re-evaluated exactly with the post-merge u (none touches x or y)
            {
                double pv = 1e300;
                if (tid < CPOOL && pool_i[tid] >= 0) {
                    const int pi = pool_i[tid], pj = pool_j[tid];
                    const double d = pool_d[tid];
                    const double upi = ld_peer_f64(&u_s[((pi >> 5) / CS) * 32 + (pi & 31)], (pi >> 5) % CS);
                    const double upj = ld_peer_f64(&u_s[((pj >> 5) / CS) * 32 + (pj & 31)], (pj >> 5) % CS);
                    pv = (d - upi) - upj;
                }
                if (tid < CPOOL) {
                    pv = warp_min_f64(pv);
                    if (lane == 0) s_red[w] = pv;
                }
                __syncthreads();
                if (tid == 0) {
                    double m = s_red[0];
                    for (int q = 1; q < CPOOL / 32; q++) m = fmin(m, s_red[q]);
                    s_ubmin = m;
                }
            }
            CL_MARK(7);
            cluster.sync();
            CL_MARK(8);

            // ------------------------------------------------------------ B2: upper bound, fold column x, select
            if (w == 0) {
                const double m = warp_min_f64(lane < CS ? ld_peer_f64(&s_ubmin, lane) : 1e300);
                if (lane == 0) s_ub = m;
            }
            __syncthreads();
            ub = s_ub;
            CL_MARK(9);
            {
                const double margin = 1e-9 * (4.0 * dmax + fabs(C));
                const int nch = (n + 31) >> 5;
                for (int lw = w; lw * CS + rank < nch; lw += NW) {
                    const int i = (lw * CS + rank) * 32 + lane;
                    const int s = lw * 32 + lane;
                    bool take = false;
                    if (i < n) {
                        take = (i == x);                          // the new row is always rescanned
                        if (!take) {
                            const unsigned long long kc = enc_f64((v_s[s] - ux) + C);
                            unsigned long long ko = __ldcg(&K[i]);
                            if (kc < ko) { ko = kc; K[i] = kc; }
                            const double lb = (dec_f64(ko) - C) - u_s[s] - margin;
                            take = (ko == 0ull) || !(lb > ub);
                        }
                    }
                    const unsigned int bal = __ballot_sync(0xffffffffu, take);
                    if (bal) {
                        unsigned int base = 0;
                        if (lane == 0) base = atomicAdd(sel0, (unsigned int)__popc(bal));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        if (take) {
                            sel_rows[base + __popc(bal & ((1u << lane) - 1u))] = i;
                            K[i] = 0xffffffffffffffffull;         // reset, the scan lowers it atomically
                        }
                    }
                }
            }
        } else {
            // first search: every row
            const int nch = (n + 31) >> 5;
            for (int lw = w; lw * CS + rank < nch; lw += NW) {
                const int i = (lw * CS + rank) * 32 + lane;
                const bool take = i < n;
                const unsigned int bal = __ballot_sync(0xffffffffu, take);
                unsigned int base = 0;
                if (lane == 0) base = atomicAdd(sel0, (unsigned int)__popc(bal));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (take) {
                    sel_rows[base + __popc(bal & ((1u << lane) - 1u))] = i;
                    K[i] = 0xffffffffffffffffull;
                }
            }
        }
        CL_MARK(10);
        cluster.sync();
        CL_MARK(11);

        // ---------------------------------------------------------------- C: scan own column chunks of the selected rows
        {
            if (tid == 0) s_nsel = (int)*sel0;
            const bool merged = !first;
            const bool ymoved = merged && y < n;               // false when y was the last slot: nothing moved into it
            // Deferred column writes of this merge (D[i][x], D[i][y] for owned rows) are scattered 8-byte stores:
            // ~2 cycles each per SM, and loads queue behind them (a single L2 load took 1600 cycles right after
            // a burst; tools/cluster_microbench.cu).  So each warp slips one chunk (64 stores) behind the loads of
            // each scan unit, and the scan takes columns x and y from v_s / f_s instead of D.
            const int st_nch = (merged && HC == 0) ? (n + 31) >> 5 : 0;   // with helper clusters the main cluster stores no columns
            int st_lw = w;
            auto column_stores = [&]() {
                const int i = (st_lw * CS + rank) * 32 + lane;
                if (i < n && i != x && i != y) {
                    D[(size_t)i * ld + x] = v_s[st_lw * 32 + lane];
                    if (ymoved) D[(size_t)i * ld + y] = f_s[st_lw * 32 + lane];
                }
                st_lw += NW;
            };
            __syncthreads();
            CL_MARK(12);
            const int nsel = s_nsel;
            const int nch = (n + 31) >> 5;
            const int lch = nch > rank ? (nch - rank + CS - 1) / CS : 0;      // local chunks holding columns < n
            const int parts = (lch + UC - 1) / UC;
            const int units = nsel * parts;                                   // <= 131 072 * 32
            // local chunk / lane of columns x and y when this CTA owns them
            const int xlw = (merged && ((x >> 5) % CS) == rank) ? (x >> 5) / CS : -1000000;
            const int ylw = (ymoved && ((y >> 5) % CS) == rank) ? (y >> 5) / CS : -1000000;
            double bt = 1e300, bd = 0.0, bui = 0.0, buj = 0.0;
            int bi = -1, bj = -1;
            int un = w;
            const int pdiv = parts > 0 ? parts : 1;
            int r_next = un < units ? __ldcg(&sel_rows[un / pdiv]) : 0;
            if (PROF && r_next >= 0) CL_MARK(18);
            // one pass = one scan unit (8 column chunks of one selected row) + one chunk of column stores; a warp
            // that has run out of one of the two keeps going with the other (a dead unit loads nothing)
            for (; un < units || st_lw * CS + rank < st_nch; un += NW) {
                const bool live = un < units;
                const int r = live ? r_next : 0;
                const int lw0 = live ? (un % pdiv) * UC : lch;
                if (un + NW < units) r_next = __ldcg(&sel_rows[(un + NW) / pdiv]);
                const double* row = D + (size_t)r * ld;
                double dv[UC];
#pragma unroll
                for (int q = 0; q < UC; q++) {
                    const int j = ((lw0 + q) * CS + rank) * 32 + lane;
                    dv[q] = (live && j < n && j != r) ? __ldcg(&row[j]) : 1e300;   // 1e300: never a candidate
                }
                CL_MARK(6);
                if (st_lw * CS + rank < st_nch) column_stores();
                CL_MARK(14);
                // u[r] and, for rows other than x and y, their fresh distances to x and y (owner's shared memory)
                double rv = 0.0;
                if (lane < 3) {
                    double* src = lane == 0 ? u_s : (lane == 1 ? v_s : f_s);
                    rv = ld_peer_f64(&src[((r >> 5) / CS) * 32 + (r & 31)], (r >> 5) % CS);
                }
                if (PROF && rv > -1.0) CL_MARK(22);
                const double ur = __shfl_sync(0xffffffffu, rv, 0);
                const double vr = __shfl_sync(0xffffffffu, rv, 1);
                const double fr = __shfl_sync(0xffffffffu, rv, 2);
                if (PROF && fr > -1.0) CL_MARK(19);
                const bool patch = merged && r != x && r != y;
                const int xq = (patch && lane == (x & 31)) ? xlw - lw0 : -1;
                const int yq = (patch && lane == (y & 31)) ? ylw - lw0 : -1;
                // A lane's columns of one row differ by multiples of 32 * CS (a multiple of 256), so the reference
                // order within the row is plain ascending j: the first strict minimum is the right one.
                double lm = 1e300, ut = 1e300, ud = 0.0, uuj = 0.0;
                int uq = 0;
#pragma unroll
                for (int q = 0; q < UC; q++) {
                    double d = dv[q];
                    if (q == xq) d = vr;
                    if (q == yq) d = fr;
                    const double uj = u_s[(lw0 + q) * 32 + lane];
                    const double t = (d - ur) - uj;
                    lm = fmin(lm, d - uj);
                    if (t < ut) { ut = t; ud = d; uuj = uj; uq = q; }
                }
                if (ut < 10000.0 && (ut < bt || (ut == bt && bi != r && tie_before(r, ((lw0 + uq) * CS + rank) * 32 + lane, bi, bj, n)))) {
                    bt = ut; bi = r; bj = ((lw0 + uq) * CS + rank) * 32 + lane; bd = ud; bui = ur; buj = uuj;
                }
                if (PROF && bt > -1e300) CL_MARK(20);
                const unsigned long long km = warp_min_u64(lm < 1e299 ? enc_f64(lm + C) : 0xffffffffffffffffull);
                if (lane == 0 && km != 0xffffffffffffffffull) atomicMin(&K[r], km);
                if (lane == 0 && lw0 == 0 && rank == 0) my_rows++;
                if (PROF && km != 1ull) CL_MARK(21);
            }
            CL_MARK(13);
            // warp winner -> CTA winner (reference order), every warp winner also feeds the candidate pool
            {
                const int wl = warp_best_lane(bt, bi, bj, n);
                if (lane == (wl < 0 ? 0 : wl)) { wrec[w].t = bt; wrec[w].i = wl < 0 ? -1 : bi; wrec[w].j = bj; wrec[w].d = bd; wrec[w].ui = bui; wrec[w].uj = buj; }
            }
            __syncthreads();
            if (w == 0) {
                const int src = lane < NW ? lane : 0;
                const int ci = lane < NW ? wrec[src].i : -1;
                const int wl = warp_best_lane(wrec[src].t, ci, wrec[src].j, n);
                if (lane < CS) {
                    CRec* dst = cluster.map_shared_rank(&recs[rank], lane);
                    if (wl >= 0) *dst = wrec[wl]; else dst->i = -1;
                }
            }
        }
        CL_MARK(15);
        cluster.sync();
        CL_MARK(16);

        // ---------------------------------------------------------------- D: pick (identical in every CTA)
        {
            if (w == 0) {
                const int src = lane < CS ? lane : 0;
                const int ci = lane < CS ? recs[src].i : -1;
                const int wl = warp_best_lane(recs[src].t, ci, recs[src].j, n);
                if (lane == 0) wrec[0] = recs[wl < 0 ? 0 : wl];
            }
            // this scan's warp winners (still in wrec[1..], and lane 0's registers for warp 0) go to the pool below
            const int mi = (tid < NW && tid > 0) ? wrec[tid].i : -1, mj = (tid < NW && tid > 0) ? wrec[tid].j : -1;
            const double md = (tid < NW && tid > 0) ? wrec[tid].d : 0.0;
            __syncthreads();
            CL_MARK(17);
            const int wi = wrec[0].i, wj = wrec[0].j;
            const double wd = wrec[0].d, wui = wrec[0].ui, wuj = wrec[0].uj;
            double uxo, uyo;
            if (wi < wj) { x = wi; y = wj; uxo = wui; uyo = wuj; } else { x = wj; y = wi; uxo = wuj; uyo = wui; }
            dxy = wd;
            const int last_ = n - 1;
            for (int p = tid; p < CPOOL; p += CT) {
                const int pi = pool_i[p], pj = pool_j[p];
                if (pi >= 0) {
                    if (pi == x || pi == y || pj == x || pj == y) pool_i[p] = -1;
                    else {
                        if (pi == last_) pool_i[p] = y;
                        if (pj == last_) pool_j[p] = y;
                    }
                }
            }
            __syncthreads();
            if (tid > 0 && tid < NW && mi >= 0 && mi != x && mi != y && mj != x && mj != y) {
                const int slot = (s_pool_head + tid) % CPOOL;
                pool_i[slot] = mi == last_ ? y : mi;
                pool_j[slot] = mj == last_ ? y : mj;
                pool_d[slot] = md;
            }
            if (HC > 0 && tid == 0 && iter > 0) {
                // the next update reads whole rows: the helpers must have finished the columns of the previous merge
                const unsigned int want = (unsigned int)iter * (unsigned int)HC;
                unsigned int dn;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(dn) : "l"(&ctl->done) : "memory");
                } while ((int)(dn - want) < 0);
            }
            __syncthreads();
            if (tid == 0) {
                s_pool_head = (s_pool_head + NW) % CPOOL;
                if (rank == 0) {
                    // host step of the reference, src/neighborJoining.cu:219-237; the realID bookkeeping
                    // (:233-237) is replayed on the host from this log
                    double blX = (dxy + uxo - uyo) * 0.5;
                    double blY = dxy - blX;
                    if (blX < 0) { blY += blX; blX = 0; }
                    if (blY < 0) { blX += blY; blY = 0; }
                    log_xy[iter] = make_int2(x, y);
                    log_bl[iter] = make_double2(blX, blY);
                    s_sel = 0;   // next appended to after two more cluster barriers
                }
            }
            iter++;
        }
        first = false;
    }
    if (rank == 0 && tid == 0) {
        if (HC > 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&ctl->seq), "r"(0xffffffffu) : "memory");
        stats->iters = (unsigned long long)iter;
        for (int k = 0; k < 24; k++) stats->cyc[k] = s_cyc[k];
    }
    if (rank == 0 && lane == 0 && my_rows) atomicAdd(&stats->rows_scanned, my_rows);
    cluster.sync();   // no CTA may exit while peers can still read its shared memory
}

template <int CS, int CT, bool PROF>
static int launch_cluster(dipb_ctx* c, int LS, void** args, int* HC, int max_helper_clusters, bool* ok) {
    const size_t smem = sizeof(double) * ((size_t)4 * LS + LS / 32 + 2 + UC * 32);
    auto kern = nj_cluster_kernel<CS, CT, PROF>;
    *ok = false;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (CS > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS); cfg.blockDim = dim3(CT); cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess || nclusters < 1) { cudaGetLastError(); return 0; }
    // helper clusters spin on a doorbell of the main cluster: only as many as are co-resident with it
    int helpers = nclusters - 1 < max_helper_clusters ? nclusters - 1 : max_helper_clusters;
    if (helpers < 0) helpers = 0;
    *HC = helpers * CS;
    cfg.gridDim = dim3(CS * (1 + helpers));
    cudaError_t e = cudaLaunchKernelExC(&cfg, (const void*)kern, args);
    if (e != cudaSuccess) { set_error("nj_cluster: launch failed: %s", cudaGetErrorString(e)); return DIPB_E_CUDA; }
    *ok = true;
    return 0;
}

bool nj_cluster_fits(int n) {
    // 3 doubles of state per owned row (+ chunk sums) within 200 KB of shared memory per CTA, 8 CTAs at worst
    const int chunks = (n + 31) / 32;
    const int LS = ((chunks + 7) / 8) * 32;
    return n <= 131072 && sizeof(double) * ((size_t)4 * LS + LS / 32 + 2 + UC * 32) <= 200u * 1024u;
}

int nj_cluster_loop(dipb_matrix* m, double* U, double* u, int* realID, int32_t* c0, int32_t* c1, double* l0, double* l1) {
    dipb_ctx* c = m->ctx;
    const int n = m->n;
    unsigned long long* K = nullptr;
    int* sel = nullptr;
    CStats* stats = nullptr;
    int2* log_xy = nullptr;
    double2* log_bl = nullptr;
    DIPB_CUDA(cudaMalloc(&K, sizeof(unsigned long long) * n));
    DIPB_CUDA(cudaMalloc(&sel, sizeof(int) * n));
    DIPB_CUDA(cudaMalloc(&stats, sizeof(CStats)));
    NJCtl* ctl = nullptr;
    DIPB_CUDA(cudaMalloc(&ctl, sizeof(NJCtl)));
    DIPB_CUDA(cudaMemsetAsync(ctl, 0, sizeof(NJCtl), c->stream));
    DIPB_CUDA(cudaMalloc(&log_xy, sizeof(int2) * n));
    DIPB_CUDA(cudaMalloc(&log_bl, sizeof(double2) * n));
    DIPB_CUDA(cudaMemsetAsync(stats, 0, sizeof(CStats), c->stream));
    DIPB_CUDA(cudaMemsetAsync(K, 0, sizeof(unsigned long long) * n, c->stream));
    // scale of the safety margin: twice the largest |u| of the input (as nj_pruned.cu)
    double dmax = 0.0;
    {
        std::vector<double> hu(n);
        DIPB_CUDA(cudaMemcpyAsync(hu.data(), u, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
        DIPB_CUDA(cudaStreamSynchronize(c->stream));
        for (int i = 0; i < n; i++) { double a = hu[i] < 0 ? -hu[i] : hu[i]; if (a == a && a > dmax && a < 1e300) dmax = a; }
        dmax *= 2.0;
    }
    size_t ld = (size_t)n;
    double* Dp = m->d;
    int n_total = n;
    int profile = getenv("DIPB_NJ_PROFILE") ? 1 : 0;
    const int chunks = (n + 31) / 32;
    bool ok = false;
    int rc = 0;
    const char* force = getenv("DIPB_NJ_CLUSTER");   // 8 or 16; default: 16 when the device can co-schedule it
    int want = force ? atoi(force) : 16;
    int used = 0;
    int HC = 0;
    const char* hp = getenv("DIPB_NJ_HELPERS");   // helper clusters for the column stores (default: all that fit, at most 7)
    const int max_helpers = hp ? atoi(hp) : 7;
    if (want >= 16) {
        int LS = ((chunks + 15) / 16) * 32;
        void* args[] = {&Dp, &ld, &U, &u, &K, &sel, &stats, &log_xy, &log_bl, &n_total, &LS, &dmax, &ctl, &HC};
        rc = profile ? launch_cluster<16, 1024, true>(c, LS, args, &HC, max_helpers, &ok) : launch_cluster<16, 1024, false>(c, LS, args, &HC, max_helpers, &ok);
        used = 16;
    }
    if (!rc && !ok) {
        int LS = ((chunks + 7) / 8) * 32;
        void* args[] = {&Dp, &ld, &U, &u, &K, &sel, &stats, &log_xy, &log_bl, &n_total, &LS, &dmax, &ctl, &HC};
        rc = profile ? launch_cluster<8, 1024, true>(c, LS, args, &HC, max_helpers, &ok) : launch_cluster<8, 1024, false>(c, LS, args, &HC, max_helpers, &ok);
        used = 8;
    }
    if (!rc && !ok) { set_error("nj_cluster: no cluster configuration fits this device"); rc = DIPB_E_CUDA; }
    if (rc) return rc;
    c->launches++;
    DIPB_CUDA(cudaStreamSynchronize(c->stream));
    {
        // replay of realID / tree bookkeeping (src/neighborJoining.cu:233-237) from the device log
        const int iters = n - 2;
        std::vector<int2> hxy(iters);
        std::vector<double2> hbl(iters);
        std::vector<int> rid(n), hc0(n), hc1(n);
        std::vector<double> hl0(n), hl1(n);
        DIPB_CUDA(cudaMemcpy(hxy.data(), log_xy, sizeof(int2) * iters, cudaMemcpyDeviceToHost));
        DIPB_CUDA(cudaMemcpy(hbl.data(), log_bl, sizeof(double2) * iters, cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; i++) rid[i] = i;
        int id = n;
        for (int it = 0; it < iters; it++) {
            const int xx = hxy[it].x, yy = hxy[it].y, act = n - it;
            hc0[it] = rid[xx]; hl0[it] = hbl[it].x;
            hc1[it] = rid[yy]; hl1[it] = hbl[it].y;
            rid[xx] = id++; rid[yy] = rid[act - 1];
        }
        DIPB_CUDA(cudaMemcpy(c0, hc0.data(), sizeof(int32_t) * iters, cudaMemcpyHostToDevice));
        DIPB_CUDA(cudaMemcpy(c1, hc1.data(), sizeof(int32_t) * iters, cudaMemcpyHostToDevice));
        DIPB_CUDA(cudaMemcpy(l0, hl0.data(), sizeof(double) * iters, cudaMemcpyHostToDevice));
        DIPB_CUDA(cudaMemcpy(l1, hl1.data(), sizeof(double) * iters, cudaMemcpyHostToDevice));
        DIPB_CUDA(cudaMemcpy(realID, rid.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
    }
    CStats hs;
    DIPB_CUDA(cudaMemcpy(&hs, stats, sizeof(hs), cudaMemcpyDeviceToHost));
    c->nj_rows_scanned = hs.rows_scanned;
    c->nj_iterations = hs.iters;
    c->nj_bytes_scanned = 0;
    if (profile) {
        const char* nm[24] = {"-", "D rest (pool, log)", "A row loop", "A drift reduce", "barrier 1", "B1 canonical sum", "C unit: issue loads", "B1 pool eval",
                              "barrier 2", "B2 ub", "B2 fold+select", "barrier 3", "C col stores+nsel", "C unit loop", "C unit: issue col stores", "C reduce+publish",
                              "barrier 4", "D pick", "C first row index", "C unit: shfl", "C unit: loads+min", "C unit: redux+atomic", "C unit: wait DSMEM", "-"};
        double tot = 0;
        for (int k = 0; k < 24; k++) tot += (double)hs.cyc[k];
        fprintf(stderr, "[nj_cluster] n=%d cluster=%d helper_ctas=%d iters=%llu rows_scanned=%llu (%.1f/iter)\n", n, used, HC, hs.iters, hs.rows_scanned,
                hs.iters ? (double)hs.rows_scanned / hs.iters : 0.0);
        for (int k = 0; k < 24; k++)
            if (hs.cyc[k]) fprintf(stderr, "[nj_cluster]   %-18s %10.0f cyc/iter  %5.1f%%\n", nm[k], hs.iters ? hs.cyc[k] / (double)hs.iters : 0.0,
                    tot > 0 ? 100.0 * hs.cyc[k] / tot : 0.0);
    }
    cudaFree(K); cudaFree(sel); cudaFree(stats); cudaFree(ctl); cudaFree(log_xy); cudaFree(log_bl);
    return 0;
}

}  // namespace dipb
