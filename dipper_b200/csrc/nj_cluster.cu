// K4 (cluster): the bound-pruned exact NJ search of nj_pruned.cu, run by ONE thread-block cluster.
//
// Why one cluster and not the whole GPU.  After pruning an NJ iteration touches little data (~34 rows
// rescanned + 5 row/column updates: ~8 MB at 30 000 tips, shrinking linearly), but it is a chain of
// dependent steps.  Spread over 148 CTAs each step ends in a software grid barrier through L2 and every
// exchanged scalar is another L2 round trip: 25 us per iteration, the same at 4 000 tips as at 30 000.
// A 16-CTA cluster synchronises in hardware (barrier.cluster), exchanges scalars by pushing them into the
// peers' shared memory, and keeps the per-row state (U, u, lower-bound keys, the new column) in shared
// memory.  What the one GPC of the main cluster is bad at -- the 2n scattered 8-byte column stores of every
// merge -- is handed to helper clusters on the other GPCs through a doorbell in global memory.
//
// Same algorithm and result as nj_pruned.cu (see its header for the bound): replaces the loop of
// NJDeviceArrays::findNeighbourJoiningTree (src/neighborJoining.cu:196-246) with findMinDist (:117-148),
// thrust::min_element (:214) and updateDisMatrix (:161-194), tie order and U summation order included.
//
// Layout.  Rows are dealt to the CTAs in chunks of 32: chunk w = i / 32 belongs to CTA w % CS, local slot
// (w / CS) * 32 + i % 32.  A CTA owns U, u, K and the distances of its rows to the two slots the last merge
// rewrote (v: new node x, f: the node moved into y), and scans its own column chunks of every selected row
// (u of those columns is local).  D lives in global memory and is read with ld.global.cg.
//
// Iteration (3 cluster barriers; everything a peer needs after a barrier was pushed into its shared memory
// before it -- measured: pulling the same word from one CTA by 512 warps serialises for ~1500 cycles):
//   D  every WARP reduces the published CTA winners to the same (x, y) and keeps its own candidate-pool slots (no CTA
//      barrier in this stretch); rank 0 logs the merge
//   A  owners update rows x, y (move `last` into y), U (double buffered), u; push chunk sums and max u-drift; in the
//      shadow of the row loads the owners RESOLVE the previous scan (below);
//      the carried candidates are re-evaluated with the post-merge u of their two ends, which every CTA derives
//      itself from the pre-merge U (the other buffer) and rows x, y -- so the upper bound travels with the drift  | barrier
//   B  U[x] (canonical sum order), C += drift, ub = min over CTAs; ring the helpers; owners fold the new column
//      into K and select rows with lb <= ub                                                             | barrier
//   C  every CTA stages the selected rows (index, u, v, f) in shared memory and scans, of its column chunks of
//      them, only the UNITS whose own lower-bound key reaches ub (see Kb below); combines per-row minima in shared
//      memory, writes (key of its minimum, column, key of its runner-up) of every staged row into its SLOT in the
//      row owner's shared memory (one 16-byte DSMEM store, no atomics); publishes its winner             | barrier
// Resolve: the owner of a staged row reads its CS slots locally: the smallest (key, column) becomes the tracked partner
// (K1, a; the exact distance D[r][a] is fetched with phase A's loads), everything else bounds the runner-up K2.
//
// Unit keys (Kb).  A scan unit = UC column chunks of one row in one CTA (UC * 32 columns, UC * 256 bytes of D).
// Kb[row][cta][part] holds, like the row keys, (min over the unit's columns of d - u_j) + C at evaluation time,
// rounded down to fp32: a lower bound that drifts with C.  A selected row is no longer read as a whole (240 KB at
// 30 000 tips): each CTA reads its `parts` keys of the row and loads only the units that can still hold a
// candidate <= ub -- typically the unit of the runner-up that made the row's bound reach ub, 2 KB.  Units that are
// skipped contribute their key to the row's runner-up bound K2.  The new column of a merge (and the column that
// receives the moved last row) is folded into the unit keys of every row by the helper clusters together with the
// column stores (red.global.min), the scan patches those two columns itself for the merge in flight.
#include <cooperative_groups.h>
#include <chrono>
#include <cstdlib>
#include <type_traits>
#include <vector>
#include "common.cuh"
#include "nj.cuh"
#include "nj_bound.cuh"

namespace cg = cooperative_groups;

namespace dipb {

namespace {

constexpr int MAXW = 32;      // warps per CTA at most
constexpr int CPOOL = 64;     // carried candidate pairs per CTA (128: 13.46 us per merge, 64: 13.29, 32: 13.29 at C3)
constexpr int MAXCS = 16;
constexpr int TILE = 128;     // selected rows staged in shared memory at a time
constexpr int MAXPARTS = 32;  // units per row and CTA at most (the staged qualification mask is one word)
constexpr unsigned long long KMAX = 0xffffffffffffffffull;
constexpr unsigned int K32MAX = 0xffffffffu;   // row keys are order-preserving fp32, rounded DOWN (a lower bound stays one);
                                               // 32-bit min is a native shared-memory atomic, local and remote
                                               // (red.shared::cluster.min.u64 assembles but does not take effect on sm_100a)

struct CRec {                 // a CTA's best candidate of one scan
    double t, d, ui, uj;
    int i, j;
};

struct NJCtl {                     // main cluster -> helper clusters doorbell (global memory)
    unsigned long long bell;       // one word, one plain store: [seq:13 | x:17 | y:17 | n:17]; all ones = quit.
                                   // x: new node, y: slot that received the old last row when y < n, seq: merge number
    unsigned int done;             // helper CTAs that finished, cumulative
    unsigned int pad;
    // What the helpers need for the unit-key fold is only known a little later than the bell (after the canonical sum):
    // u of the new node (rounded up), u of the node moved into y (rounded up), drift sum C (rounded down), each as
    // (merge number << 32) | fp32 bits.  A word is its own arrival flag: no fence, no second doorbell.
    unsigned long long pw[3];
};

struct CStats {
    unsigned long long rows_scanned, bytes_scanned, iters, units_scanned;
    unsigned long long cyc[32];
    unsigned long long t_ns, t_cycles;   // whole main loop: globaltimer ns and SM cycles (their ratio is the SM clock)
};

// lexicographic warp minimum of a u64 through two 32-bit redux ops
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
    const unsigned int hi = (unsigned int)(v >> 32), lo = (unsigned int)v;
    const unsigned int mh = __reduce_min_sync(0xffffffffu, hi);
    const unsigned int ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
    return ((unsigned long long)mh << 32) | ml;
}
// minimum over each HALF of the warp.  redux.sync with two 16-lane masks costs ~175 cycles on sm_100a (the halves are
// serialised on a slow path; tools/experiments/warp_ops_latency.cu), a full-mask one 22: two full-mask ops, each half's
// lanes standing aside in the other's.
__device__ __forceinline__ unsigned int half_min_u32(unsigned int v, int lane) {
    const unsigned int lo = __reduce_min_sync(0xffffffffu, lane < 16 ? v : 0xffffffffu);
    const unsigned int hi = __reduce_min_sync(0xffffffffu, lane < 16 ? 0xffffffffu : v);
    return lane < 16 ? lo : hi;
}
__device__ __forceinline__ double warp_min_f64(double v) { return dec_f64(warp_min_u64(enc_f64(v))); }
__device__ __forceinline__ double warp_max_f64(double v) { return -warp_min_f64(-v); }
// Warp maximum / minimum of values that only steer the pruning (the per-merge drift of u, the upper bound of the carried
// candidates): rounded UP to fp32 first -- larger is the safe side for both -- so that one redux does what the exact fp64
// form needs two of, plus the 64-bit encode / decode.  (Sentinels +-1e300 become +-FLT_MAX / inf, still sentinels.)
__device__ __forceinline__ double warp_max_up32(double v) { return (double)dec_f32(__reduce_max_sync(0xffffffffu, enc_f32(__double2float_ru(v)))); }
__device__ __forceinline__ double warp_min_up32(double v) { return (double)dec_f32(__reduce_min_sync(0xffffffffu, enc_f32(__double2float_ru(v)))); }

// (Row / column indices are never negative where an owner or a slot is derived from them: the casts to unsigned below turn
// the divisions by the cluster size into one shift or mask each -- the signed forms cost ~5 instructions apiece, in every warp.)
// ---- distributed shared memory: the same variable in CTA `rank` of the cluster
__device__ __forceinline__ unsigned int peer_addr(const void* p, int rank) {
    const unsigned int a = (unsigned int)__cvta_generic_to_shared(p);
    unsigned int ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    return ra;
}
__device__ __forceinline__ double ld_peer_f64(const double* p, int rank) {
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(peer_addr(p, rank)) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_peer_u32(const unsigned int* p, int rank) {
    unsigned int v;
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(peer_addr(p, rank)) : "memory");
    return v;
}
__device__ __forceinline__ void st_peer_f64(double* p, int rank, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(peer_addr(p, rank)), "d"(v) : "memory");
}
__device__ __forceinline__ void st_peer_v4(uint4* p, int rank, unsigned int a, unsigned int b, unsigned int c, unsigned int d) {
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(peer_addr(p, rank)), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_peer_s32(int* p, int rank, int v) {
    asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(peer_addr(p, rank)), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_peer_s32(const int* p, int rank) {
    int v;
    asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(peer_addr(p, rank)) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int key_of(double v) { return enc_f32(__double2float_rd(v)); }
// a / d, correctly rounded, from r = RN(1 / d): q = RN(a r), rem = a - q d (exact in the FMA), RN(q + rem r) is the IEEE quotient
// (Markstein 1990; d = n - 3 is a small integer, no overflow or underflow on this path).  Bit-identical to the division the
// reference and the oracle perform, at 3 instructions per row instead of the ~14 of a division; checked against a / d for
// every d <= 140 000 on 56 M operands (tools/experiments/markstein_div.c) and by the bit-exact tree tests.
__device__ __forceinline__ double div_rn(double a, double d, double r) {
    const double q = __dmul_rn(a, r);
    return __fma_rn(__fma_rn(-q, d, a), r, q);
}

// ---- candidate order (reference scan order, nj_bound.cuh), out of line: exact ties are rare
__device__ __noinline__ bool tie_before(int ia, int ja, int ib, int jb, int n) {
    if (ia == ib) return ((ja & 255) < (jb & 255)) || ((ja & 255) == (jb & 255) && ja < jb);
    return p_tie_key(ia, ja, n) < p_tie_key(ib, jb, n);
}
__device__ __noinline__ int tie_lane(unsigned int tied, int i, int j, int n) {
    const bool in = (tied >> (threadIdx.x & 31)) & 1u;
    const unsigned long long key = in ? p_tie_key(i, j, n) : KMAX;
    const unsigned long long km = warp_min_u64(key);
    return __ffs(__ballot_sync(0xffffffffu, key == km)) - 1;
}
// lane holding the best candidate of the warp (t ascending, then reference order), -1 when no lane has one
__device__ __forceinline__ int warp_best_lane(double t, int i, int j, int n) {
    const unsigned long long e = i >= 0 ? enc_f64(t) : KMAX;
    const unsigned long long m = warp_min_u64(e);
    if (m == KMAX) return -1;
    const unsigned int tied = __ballot_sync(0xffffffffu, e == m);
    if ((tied & (tied - 1u)) == 0u) return __ffs(tied) - 1;
    return tie_lane(tied, i, j, n);
}

// bytes of dynamic shared memory for LS owned rows per CTA and `chunks` 32-row chunks in total
inline size_t cluster_smem_bytes(int LS, int chunks) {
    return sizeof(double) * ((size_t)6 * LS + chunks + 2) + sizeof(unsigned int) * (size_t)3 * LS + (size_t)TILE * (3 * 8 + 8 + 4 + 4 + 4) +
           (size_t)TILE * MAXCS * 16;   // + the result slots
}

}  // namespace

template <int CS, int CT, int UC, bool PROF>   // UC: column chunks per scan unit (loads in flight per lane)
__global__ void __launch_bounds__(CT, 1)
nj_cluster_kernel(double* __restrict__ D, size_t ld, const double* __restrict__ U0, const double* __restrict__ u0,
                  int* __restrict__ sel_rows, CStats* stats, int2* __restrict__ log_xy, double2* __restrict__ log_bl,
                  int n_total, int LS, double dmax, NJCtl* ctl, int HC, unsigned int* __restrict__ Kb, int PARTS, int dbg, double slack_merges, double* __restrict__ R) {
    if (blockIdx.x >= CS) {
        // ---- helper clusters (the other GPCs): transpose rows x and y of each published merge into columns x and
        // y, and fold those two new columns into the unit keys of every row.  These 2n scattered 8-byte stores (and
        // 2n 4-byte minima) per merge are request-rate bound on one GPC's L2 port when the main cluster issues them
        // itself (~1 us per 1000 tips); spread over the other GPCs they are off the critical path.
        const int hc = (int)blockIdx.x - CS, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
        __shared__ unsigned long long s_msg;
        __shared__ float s_pw[3];
        unsigned long long seen = 0;
        unsigned int hit = 0;
        for (;;) {
            if (tid == 0) {
                unsigned long long q;
                for (;;) {
                    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(q) : "l"(&ctl->bell) : "memory");
                    if (q != seen) break;
                    __nanosleep(64);
                }
                s_msg = q;
            }
            __syncthreads();
            const unsigned long long q = s_msg;
            __syncthreads();
            if (q == KMAX) return;
            seen = q;
            const int x = (int)((q >> 34) & 0x1ffffu), y = (int)((q >> 17) & 0x1ffffu), n = (int)(q & 0x1ffffu);
            hit++;                                                  // merges are published one by one: this is merge number `hit`
            const double* Rx = R + (size_t)(hit & 1) * 2 * ld;      // the new rows x and y as phase A left them (scratch)
            const double* Ry = Rx + ld;
            const bool ymoved = y < n;
            // key rows of the units that hold columns x and y (Kb is [cta][part][row]: this fold is coalesced over the rows)
            const size_t KLD = (size_t)((n_total + 31) & ~31);
            unsigned int* const Kx = Kb + ((size_t)((int)(((unsigned int)x >> 5) % CS)) * PARTS + ((int)(((unsigned int)x >> 5) / CS)) / UC) * KLD;
            unsigned int* const Ky = Kb + ((size_t)((int)(((unsigned int)y >> 5) % CS)) * PARTS + ((int)(((unsigned int)y >> 5) / CS)) / UC) * KLD;
            const int nch = (n + 31) >> 5;
            // 1. rows and columns x, y of D
            double vx[2], fy[2];
            int cnt = 0;
            for (int c = hc + w * HC; c < nch; c += HC * (CT / 32)) {
                const int i = c * 32 + lane;
                double v = 0.0, f = 0.0;
                if (i < n) {
                    if (i != x) { v = __ldcg(&Rx[i]); D[(size_t)x * ld + i] = v; if (!(dbg & 8)) D[(size_t)i * ld + x] = v; }
                    if (ymoved && i != y) { f = __ldcg(&Ry[i]); D[(size_t)y * ld + i] = f; if (!(dbg & 8)) D[(size_t)i * ld + y] = f; }
                }
                if (cnt < 2) { vx[cnt] = v; fy[cnt] = f; }
                cnt++;
            }
            // 2. the two new columns folded into the unit keys of every other row, once the payload has arrived
            if (tid == 0) {
                unsigned long long a0, a1, a2;
                do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a0) : "l"(&ctl->pw[0]) : "memory"); } while ((unsigned int)(a0 >> 32) != hit);
                do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a1) : "l"(&ctl->pw[1]) : "memory"); } while ((unsigned int)(a1 >> 32) != hit);
                do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a2) : "l"(&ctl->pw[2]) : "memory"); } while ((unsigned int)(a2 >> 32) != hit);
                s_pw[0] = __uint_as_float((unsigned int)a0); s_pw[1] = __uint_as_float((unsigned int)a1); s_pw[2] = __uint_as_float((unsigned int)a2);
            }
            __syncthreads();
            const double ux = (double)s_pw[0], uy = (double)s_pw[1], Cm = (double)s_pw[2];
            cnt = 0;
            for (int c = hc + w * HC; c < nch; c += HC * (CT / 32)) {
                const int i = c * 32 + lane;
                if (i < n && i != x && i != y && !(dbg & 4)) {
                    const double v = cnt < 2 ? vx[cnt < 2 ? cnt : 0] : __ldcg(&Rx[i]);
                    atomicMin(&Kx[i], key_of((v - ux) + Cm));
                    if (ymoved) {
                        const double f = cnt < 2 ? fy[cnt < 2 ? cnt : 0] : __ldcg(&Ry[i]);
                        atomicMin(&Ky[i], key_of((f - uy) + Cm));
                    }
                }
                cnt++;
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
    const int chunks_total = (n_total + 31) >> 5;
    // unit keys, [cta][part][row] (row stride KLD): the per-merge fold of a new column into one unit of EVERY row touches
    // one contiguous key row (120 KB at 30 000 tips) instead of one cache line per row (7.7 MB), so the keys stay in L2
    const size_t KLD = (size_t)((n_total + 31) & ~31);
    unsigned int* const Kmine = Kb + (size_t)rank * PARTS * KLD;   // + part * KLD + row

    extern __shared__ __align__(16) unsigned char smem_raw[];
    // result slots of the staged tile: slot_s[k][q] = what CTA q found in its slice of staged row k (key of its minimum, its
    // column, key of its runner-up), written by CTA q into the shared memory of the row's OWNER with one 16-byte store
    uint4* slot_s = reinterpret_cast<uint4*>(smem_raw);
    double* U_s = reinterpret_cast<double*>(smem_raw + (size_t)TILE * MAXCS * 16);   // [2][LS] row sums of owned rows; buffer `cur` is valid, A writes the other
    double* u_s = U_s + 2 * LS;                          // [LS] U / (n - 2)
    double* v_s = u_s + LS;                              // [LS] distance to the newest node x
    double* f_s = v_s + LS;                              // [LS] distance to the node moved into slot y
    double* da_s = f_s + LS;                             // [LS] distance to the tracked best partner a (exact, constant while both live)
    double* cs_all = da_s + LS;                          // [chunks_total] chunk sums of the new column, pushed by the owners
    double* t_u = cs_all + chunks_total + 2;             // staged selected rows: u, v, f, combined minimum, index
    double* t_v = t_u + TILE;
    double* t_f = t_v + TILE;
    unsigned long long* t_best = reinterpret_cast<unsigned long long*>(t_f + TILE);   // (key of the row minimum << 32) | its column
    unsigned int* t_k2 = reinterpret_cast<unsigned int*>(t_best + TILE);              // key of the runner-up
    int* t_row = reinterpret_cast<int*>(t_k2 + TILE);
    unsigned int* t_qm = reinterpret_cast<unsigned int*>(t_row + TILE);               // bit p: unit p of this CTA must be scanned
    // Lower-bound keys of owned rows.  K1 bounds the tracked best partner a_s (whose exact value can be re-evaluated
    // from da_s and u[a]), K2 every other column; both are (value + C) at evaluation time, so `dec(K) - C` stays a
    // lower bound while u drifts.
    unsigned int* K1_s = t_qm + TILE;
    unsigned int* K2_s = K1_s + LS;
    int* a_s = reinterpret_cast<int*>(K2_s + LS);

    __shared__ CRec recs[MAXCS];            // winners published by every CTA of the cluster
    __shared__ CRec wrec[MAXW];
    __shared__ double drift_all[MAXCS], ubmin_all[MAXCS];   // pushed by the peers
    __shared__ double s_red[MAXW], s_red2[MAXW], s_blk[128 + 16];   // s_blk: one sum per 1024-row block (n <= 131 072)
    __shared__ double s_total, s_C, s_ub, s_uy;   // s_uy (rank 0's copy): u of the node moved into slot y, pushed by its owner
    __shared__ unsigned int s_sel;          // selected-row counter (rank 0's copy is the live one)
    __shared__ int pool_i[CPOOL], pool_j[CPOOL];
    __shared__ double pool_d[CPOOL];        // d of a carried pair never changes while both ends survive
    __shared__ double pool_t[CPOOL];        // its value at the last re-evaluation: new candidates replace worse ones only
    __shared__ int s_nsel;
    __shared__ unsigned int s_nunits;          // live scan units of the staged tile
    __shared__ unsigned short s_ulist[TILE * 12];   // (staged row << 5) | unit, in no particular order
    __shared__ unsigned long long s_cyc[32];   // rank 0, thread 0: cycles per phase (DIPB_NJ_PROFILE)

    // ---- load owned state
    for (int s = tid; s < LS; s += CT) {
        const int i = ((s >> 5) * CS + rank) * 32 + (s & 31);
        U_s[s] = i < n_total ? U0[i] : 0.0;
        U_s[LS + s] = 0.0;
        u_s[s] = i < n_total ? u0[i] : 0.0;
        v_s[s] = 0.0; f_s[s] = 0.0;
        K1_s[s] = K32MAX; K2_s[s] = K32MAX; a_s[s] = -1; da_s[s] = 0.0;
    }
    if (tid == 0) t_row[0] = 0;
    for (int p = tid; p < CPOOL; p += CT) { pool_i[p] = -1; pool_j[p] = -1; pool_t[p] = 1e300; }
    if (tid == 0) { s_sel = 0; s_uy = 0.0; }
    if (tid < 32) s_cyc[tid] = 0;
    __syncthreads();
    cluster.sync();

    int n = n_total;
    double C = 0.0;
    int x = -1, y = -1;
    double dxy = 0.0;
    bool first = true;
    int iter = 0;
    int cur = 0;                              // valid U buffer
    unsigned long long my_rows = 0, my_units = 0;
    unsigned int* sel0 = cluster.map_shared_rank(&s_sel, 0);

    // Resolve (runs at the OWNER of a staged row, after the cluster barrier that follows the slot pushes): the row's
    // minimum over the CS slices becomes its tracked partner (K1, a, exact distance da: re-evaluated exactly later instead
    // of rescanning the row while only the bound has drifted), everything else bounds the runner-up K2.  No cross-CTA
    // round trip: the slots are local.  Half a warp per row, lane q reads slot q.  A selection that fitted one tile is
    // resolved in phase A of the NEXT merge, in the shadow of its row loads (`defer`; nothing reads K1 / K2 / a / da before
    // that merge's selection); the tiles of a longer one right after their scan.
    const double* pend_ptr = nullptr;
    int pend_rs = -1;
    unsigned int done_early = 0;
    double rden = 0.0;                        // 1 / (n - 3): reciprocal of the next merge's divisor, computed inside barrier 2
    int pool_head = 0;                        // rotating insertion point of the candidate pool (same in every thread)
    auto resolve = [&](int tn, bool merged, int x, int y, int n, int iter, bool defer) {
        for (int k = w * 2 + (lane >> 4); k - (lane >> 4) < tn; k += 2 * NW) {
            const int hl = lane & 15;
            const bool have = k < tn;
            const int r = have ? t_row[k] : 0;
            const int ro = (int)(((unsigned int)r >> 5) % CS), rs = ((int)(((unsigned int)r >> 5) / CS)) * 32 + (r & 31);
            const bool mine = have && ro == rank;
            uint4 e = make_uint4(K32MAX, 0xffffffffu, K32MAX, 0u);
            if (mine && hl < CS) e = slot_s[k * MAXCS + hl];
            const unsigned int k1m = half_min_u32(e.x, lane);
            const unsigned int j1m = half_min_u32(e.x == k1m ? e.y : 0xffffffffu, lane);
            const bool win = e.x == k1m && e.y == j1m;
            const unsigned int k2m = half_min_u32(win ? e.z : (e.x < e.z ? e.x : e.z), lane);
            if (!mine || hl != 0) continue;
            K1_s[rs] = k1m; K2_s[rs] = k2m;
            if (k1m == K32MAX) continue;                   // nothing scanned: a_s stays -1 (reset by the selection)
            const int j1 = (int)j1m;
            a_s[rs] = j1;
            if (defer) {
                // (phase A: D is complete, the pick has waited for the helpers.)  The first row of a lane is only recorded and
                // loaded behind phase A's row loads; more than one row per lane is rare (> 30 selected rows)
                if (pend_rs < 0) { pend_rs = rs; pend_ptr = &D[(size_t)r * ld + j1]; }
                else da_s[rs] = __ldcg(&D[(size_t)r * ld + j1]);
                continue;
            }
            // between the tiles of a long selection the helpers may still be writing D: columns x and y of an old row come
            // from v / f of the row, rows x and y from the scratch pair
            double da;
            if (merged && r != x && r != y && j1 == x) da = t_v[k];
            else if (merged && r != x && r != y && j1 == y && y < n) da = t_f[k];
            else if (merged && r == x) da = __ldcg(&R[(size_t)(iter & 1) * 2 * ld + j1]);
            else if (merged && r == y && y < n) da = __ldcg(&R[((size_t)(iter & 1) * 2 + 1) * ld + j1]);
            else da = __ldcg(&D[(size_t)r * ld + j1]);
            da_s[rs] = da;
        }
    };

    long long tmark = clock64();
    const long long cyc_begin = tmark;
    unsigned long long ns_begin;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_begin));
#define CL_MARK(k)                                                      \
    do {                                                                \
        if (PROF && rank == 0 && tid == 0) {                            \
            long long now__ = clock64();                                \
            s_cyc[k] += (unsigned long long)(now__ - tmark);            \
            tmark = now__;                                              \
        }                                                               \
    } while (0)

    while (n > 2) {
        double ub = 1e300, ux = 0.0;
        const double margin = 1e-9 * (4.0 * dmax + fabs(C));   // refreshed below once C has moved
        double marg = margin;
        if (!first) {
            // ------------------------------------------------------------ A: merge update by row owners
            CL_MARK(0);
            const int last = n - 1;
            const double den_new = (double)(n - 3);
            const int nchunk = (last + 31) >> 5;               // chunks holding rows < last
            double* const Rx = R + (size_t)(iter & 1) * 2 * ld;  // scratch rows of this merge
            double* const Ry = Rx + ld;
            const double* Uo = U_s + cur * LS;                  // pre-merge sums (read by every CTA through DSMEM)
            double* Un = U_s + (cur ^ 1) * LS;                  // post-merge sums
            // carried candidates of this CTA, re-evaluated exactly with the post-merge u of their ends (none touches x;
            // an end in slot y is the node that lived in `last`): same expressions as the owners use below.  Four candidates
            // per warp, one lane per (candidate, end); the loads are issued here and consumed after the update loop.
            constexpr int PCW = CPOOL / (CT / 32);              // candidates per warp (4 at 1024 threads)
            static_assert(PCW * (CT / 32) == CPOOL && 2 * PCW <= 32, "pool layout: PCW candidates per warp, two lanes each");
            double p_a = 0.0, p_b = 0.0, p_U = 0.0;
            const int pc = w * PCW + (lane >> 1);               // candidate of this lane (lanes 0 .. 2 PCW - 1)
            const int pend = (lane < 2 * PCW && n > 3 && pool_i[pc] >= 0) ? ((lane & 1) ? pool_j[pc] : pool_i[pc]) : -1;
            if (pend >= 0) {
                const int src = pend == y ? last : pend;
                p_a = __ldcg(&D[(size_t)x * ld + src]); p_b = __ldcg(&D[(size_t)y * ld + src]);
                p_U = ld_peer_f64(&Uo[((int)(((unsigned int)src >> 5) / CS)) * 32 + (src & 31)], (int)(((unsigned int)src >> 5) % CS));
            }
            double pend_da = 0.0;
            bool resolved = false;
            double dmx = -1e300;
            // A warp owns chunks lw = w, w + NW, ... (two at 30 000 tips): the loads of up to AB of them are issued before any
            // is consumed -- one exposed memory round trip per merge instead of one per chunk (measured 5.4 k -> cycles below)
            constexpr int AB = CT >= 1024 ? 3 : (CT >= 512 ? 4 : 8);
            for (int lwb = w; lwb * CS + rank < nchunk; lwb += NW * AB) {
                double la[AB], lb[AB], lf[AB];
#pragma unroll
                for (int c = 0; c < AB; c++) {
                    const int lw = lwb + c * NW;
                    const int i = (lw * CS + rank) * 32 + lane;
                    la[c] = lb[c] = lf[c] = 0.0;
                    if (lw * CS + rank < nchunk && i < last && i != x) {
                        const int src = i == y ? last : i;
                        la[c] = __ldcg(&D[(size_t)x * ld + src]); lb[c] = __ldcg(&D[(size_t)y * ld + src]);
                        lf[c] = __ldcg(&D[(size_t)last * ld + i]);
                    }
                }
                if (PROF && lwb == w) { CL_MARK(18); }
                if (!resolved) {
                    // the previous scan's results, in the shadow of the loads just issued (see `resolve`)
                    if (s_nsel <= TILE) resolve(s_nsel, true, -1, -1, n, iter, true);
                    if (pend_rs >= 0) pend_da = __ldcg(pend_ptr);
                    resolved = true;
                }
                if (PROF && lwb == w) { CL_MARK(19); }
#pragma unroll
                for (int c = 0; c < AB; c++) {
                    const int lw = lwb + c * NW;
                    if (!(lw * CS + rank < nchunk)) break;
                    const int i = (lw * CS + rank) * 32 + lane;
                    const int s = lw * 32 + lane;
                    double slot = 0.0;
                    if (i < last && i != x) {
                        // slot y receives the node that lived in row `last`: same formulas, sources taken from `last`
                        const bool isy = (i == y);
                        const double a = la[c], b = lb[c], far = lf[c];
                        double Ui = Uo[s], uo = u_s[s];
                        if (isy) {
                            const int lo = (int)(((unsigned int)last >> 5) % CS), ls = ((int)(((unsigned int)last >> 5) / CS)) * 32 + (last & 31);
                            Ui = ld_peer_f64(&Uo[ls], lo);
                            uo = ld_peer_f64(&u_s[ls], lo);
                        }
                        const double val = (a + b - dxy) * 0.5;
                        Ui += -a - b + val;
                        Un[s] = Ui;
                        f_s[s] = far;
                        // the new rows x and y go to a scratch pair, not into D: phase A must not write D, because every CTA
                        // still reads the pre-merge rows x and y for the ends of its carried candidates; the helpers copy the
                        // scratch rows into D (rows and columns), the scan of this merge reads them from the scratch
                        Rx[i] = val;
                        if (isy) Ry[x] = val; else Ry[i] = far;
                        slot = val;
                        if (n > 3) {
                            const double un = div_rn(Ui, den_new, rden);
                            dmx = fmax(dmx, un - uo);
                            u_s[s] = un;
                            if (isy) st_peer_f64(&s_uy, 0, un);      // the helpers' key fold needs it (rank 0 rings the bell)
                        }
                    } else if (i == x) Un[s] = 0.0;            // replaced by the canonical sum in phase B
                    v_s[s] = slot;
                    // canonical block sum, level 1: this chunk's stride-halving tree, pushed to every CTA
                    const double csum = __shfl_sync(0xffffffffu, warp_tree_sum(slot), 0);
                    if (lane < CS) st_peer_f64(&cs_all[lw * CS + rank], lane, csum);
                }
            }
            CL_MARK(20);
            if (!resolved) {           // (a warp without chunks)
                if (s_nsel <= TILE) resolve(s_nsel, true, -1, -1, n, iter, true);
                if (pend_rs >= 0) pend_da = __ldcg(pend_ptr);
            }
            if (pend_rs >= 0) { da_s[pend_rs] = pend_da; pend_rs = -1; }
            // unit keys of the row that moves from `last` into slot y: every CTA copies its own units
            if (y < last && tid >= CT - 32 && lane < PARTS) Kmine[(size_t)lane * KLD + y] = __ldcg(&Kmine[(size_t)lane * KLD + last]);
            {
                double un = 0.0;
                if (pend >= 0) {
                    const double val = (p_a + p_b - dxy) * 0.5;
                    p_U += -p_a - p_b + val;
                    un = div_rn(p_U, den_new, rden);
                }
                const double uo2 = __shfl_xor_sync(0xffffffffu, un, 1);
                double pv = 1e300;
                if (pend >= 0 && !(lane & 1)) pv = (pool_d[pc] - un) - uo2;
                if (lane < 2 * PCW && !(lane & 1)) pool_t[pc] = pv;
                pv = warp_min_up32(pv);
                if (lane == 0) s_red2[w] = pv;
            }
            dmx = warp_max_up32(dmx);
            if (lane == 0) s_red[w] = dmx;
            CL_MARK(21);
            __syncthreads();
            CL_MARK(22);
            if (w == 0) {
                const double m = warp_max_up32(lane < NW ? s_red[lane] : -1e300);
                const double pm = warp_min_up32(lane < NW ? s_red2[lane] : 1e300);
                if (lane < CS) { st_peer_f64(&drift_all[rank], lane, m); st_peer_f64(&ubmin_all[rank], lane, pm); }
            }
            cur ^= 1;
            CL_MARK(1);
            cluster.sync();
            CL_MARK(2);

            // ------------------------------------------------------------ B: U[x], drift, upper bound, fold, select
            n = last;
            if (n <= 2) {
                // last merge: nj_finish_kernel reads D[0][1] = distance of the two nodes left (slots 0 and 1; x is one of them)
                if (rank == 0 && tid == 0) D[1] = x == 0 ? v_s[1] : v_s[0];
                break;
            }
            if (HC > 0 && rank == 0 && tid == 32) {
                // the scratch rows are complete and fenced (every thread ran MEMBAR.GPU before the barrier), so the bell is
                // one relaxed store; merge numbers differ in the low 13 bits between consecutive merges
                const unsigned long long q = ((unsigned long long)(iter & 0x1fff) << 51) | ((unsigned long long)x << 34) |
                                             ((unsigned long long)y << 17) | (unsigned long long)n;
                asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(&ctl->bell), "l"(q) : "memory");
            }
            {
                // canonical sum, level 2: 1024-row blocks (32 chunk sums, stride-halving tree), then blocks ascending
                const int nblk = (last + 1023) >> 10;
                for (int b = w; b < nblk; b += NW) {
                    const int cw = b * 32 + lane;
                    const double v = warp_tree_sum(cw < nchunk ? cs_all[cw] : 0.0);
                    if (lane == 0) s_blk[b] = v;
                }
                if (w == NW - 1) {
                    // drift and upper bound of the cluster (pushed by every CTA in phase A): one warp, once -- every thread
                    // folding the 16 values itself cost ~800 cycles of issue
                    const double dr = warp_max_up32(lane < CS ? drift_all[lane] : -1e300);
                    const double um = warp_min_up32(lane < CS ? ubmin_all[lane] : 1e300);
                    if (lane == 0) {
                        s_C = C + dr;
                        s_ub = um;
                        if (HC > 0 && rank == 0) {
                            const unsigned long long w2 = ((unsigned long long)(unsigned int)iter << 32) | __float_as_uint(__double2float_rd(C + dr));
                            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(&ctl->pw[2]), "l"(w2) : "memory");
                        }
                    }
                }
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
                    if (HC > 0 && rank == 0) {
                        // what the helpers' key fold needs, each word tagged with the merge number (see NJCtl)
                        const unsigned long long tag = (unsigned long long)(unsigned int)iter << 32;
                        const unsigned long long w0 = tag | __float_as_uint(__double2float_ru(div_rn(acc, (double)(n - 2), rden)));
                        const unsigned long long w1 = tag | __float_as_uint(__double2float_ru(s_uy));
                        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(&ctl->pw[0]), "l"(w0) : "memory");
                        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(&ctl->pw[1]), "l"(w1) : "memory");
                    }
                }
                __syncthreads();
            }
            CL_MARK(3);
            const double total = s_total;
            ux = div_rn(total, (double)(n - 2), rden);   // (correctly rounded, see div_rn)
            C = s_C;
            marg = 1e-9 * (4.0 * dmax + fabs(C));
            if (((int)(((unsigned int)x >> 5) % CS)) == rank && tid == 0) {
                const int sx = ((int)(((unsigned int)x >> 5) / CS)) * 32 + (x & 31);
                U_s[cur * LS + sx] = total;
                u_s[sx] = ux;
            }
            ub = s_ub;
            if (dbg & 2) ub = 1e300;
            CL_MARK(4);
            CL_MARK(5);
            {
                const int nch = (n + 31) >> 5;
                for (int lw = w; lw * CS + rank < nch; lw += NW) {
                    const int i = (lw * CS + rank) * 32 + lane;
                    const int s = lw * 32 + lane;
                    bool take = false;
                    if (i < n) {
                        take = (i == x);                          // the new row is always rescanned
                        if (!take) {
                            unsigned int k1 = K1_s[s], k2 = K2_s[s];
                            int a = a_s[s];
                            if (i == y && y < n) {
                                // slot y now holds the row that lived in `last` (= n): take over its keys and partner
                                const int lo = (int)(((unsigned int)n >> 5) % CS), ls = ((int)(((unsigned int)n >> 5) / CS)) * 32 + (n & 31);
                                k1 = ld_peer_u32(&K1_s[ls], lo); k2 = ld_peer_u32(&K2_s[ls], lo);
                                a = ld_peer_s32(&a_s[ls], lo);
                                da_s[s] = ld_peer_f64(&da_s[ls], lo);
                            }
                            // the tracked partner may have been merged away (x, old y) or moved (last -> y)
                            if (a == x || a == y) { a = -1; k1 = K32MAX; }
                            else if (a == n) a = y;
                            // the new column x joins the untracked columns
                            const unsigned int kc = key_of((v_s[s] - ux) + C);
                            if (kc < k2) k2 = kc;
                            // stage 1: both keys as drifting lower bounds (no memory traffic)
                            const unsigned int km = k1 < k2 ? k1 : k2;
                            take = !(((double)dec_f32(km) - C) - u_s[s] - marg > ub);
                            if (take && a >= 0) {
                                // stage 2: the tracked partner exactly (its d never changes, u[a] from its owner)
                                const double e1 = da_s[s] - ld_peer_f64(&u_s[((int)(((unsigned int)a >> 5) / CS)) * 32 + (a & 31)], (int)(((unsigned int)a >> 5) % CS));
                                k1 = key_of(e1 + C);
                                const double rest = (double)dec_f32(k2) - C;
                                take = !((e1 < rest ? e1 : rest) - u_s[s] - marg > ub);
                            }
                            if (PROF && rank == 0 && take) {
                                // why rows are rescanned (rank 0's rows only): 12 no tracked partner, 13 partner itself is
                                // a contender, 14 the runner-up bound has drifted down to the upper bound
                                const int why = a < 0 ? 12 : ((da_s[s] - ld_peer_f64(&u_s[((int)(((unsigned int)a >> 5) / CS)) * 32 + (a & 31)], (int)(((unsigned int)a >> 5) % CS))) - u_s[s] - marg > ub ? 14 : 13);
                                atomicAdd(&s_cyc[why], 1ull);
                            }
                            if (take) { k1 = K32MAX; k2 = K32MAX; a = -1; }   // reset, the scan lowers them
                            K1_s[s] = k1; K2_s[s] = k2; a_s[s] = a;
                        } else {
                            K1_s[s] = K32MAX; K2_s[s] = K32MAX; a_s[s] = -1;
                        }
                    }
                    const unsigned int bal = __ballot_sync(0xffffffffu, take);
                    if (bal) {
                        unsigned int base = 0;
                        if (lane == 0) base = atomicAdd(sel0, (unsigned int)__popc(bal));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        // a selected row goes, with its u, v, f, straight into every CTA's staging arrays: the scan phase
                        // then starts from local shared memory (pulling the list from rank 0 and the row state from the
                        // owners cost two dependent DSMEM round trips, the first one 16 CTAs deep on one SM).  The warp
                        // pushes its (few) selected rows one after the other, lane q to CTA q: a lane looping over the 16
                        // targets by itself costs the warp 160 instructions per row instead of ~25
                        const unsigned int pos = base + __popc(bal & ((1u << lane) - 1u));
                        const double us = (take && i == x) ? ux : u_s[s], vs = v_s[s], fs = f_s[s];
                        for (unsigned int rem = bal; rem; rem &= rem - 1u) {
                            const int sl = __ffs(rem) - 1;
                            const int ri = __shfl_sync(0xffffffffu, i, sl);
                            const unsigned int rp = __shfl_sync(0xffffffffu, pos, sl);
                            const double ru = __shfl_sync(0xffffffffu, us, sl), rv = __shfl_sync(0xffffffffu, vs, sl), rf = __shfl_sync(0xffffffffu, fs, sl);
                            if (rp < (unsigned int)TILE) {
                                if (lane < CS) {
                                    st_peer_s32(&t_row[rp], lane, ri);
                                    st_peer_f64(&t_u[rp], lane, ru); st_peer_f64(&t_v[rp], lane, rv); st_peer_f64(&t_f[rp], lane, rf);
                                }
                            } else if (lane == 0) sel_rows[rp] = ri;
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
                    const unsigned int pos = base + __popc(bal & ((1u << lane) - 1u));
                    if (pos < (unsigned int)TILE) {
                        const double us = u_s[lw * 32 + lane];
                        for (int q = 0; q < CS; q++) {
                            st_peer_s32(&t_row[pos], q, i);
                            st_peer_f64(&t_u[pos], q, us); st_peer_f64(&t_v[pos], q, 0.0); st_peer_f64(&t_f[pos], q, 0.0);
                        }
                    } else sel_rows[pos] = i;
                }
            }
        }
        CL_MARK(6);
        {
            // barrier 2, split: the reciprocal the next merge divides by (a ~130-cycle dependent chain) is computed between
            // arrive and wait instead of in front of phase A's loads and of the new row's u in phase B
            // (Measured: arrive.relaxed on barriers 2 and 3 saves 0.4 us per merge -- the release is MEMBAR.ALL.GPU + ERRBAR +
            // CGAERRBAR in SASS -- but nothing then orders the DSMEM pushes before the barrier, so it stays a release.)
            auto tok = cluster.barrier_arrive();
            rden = 1.0 / (double)(n - 3);
            cluster.barrier_wait(std::move(tok));
        }
        CL_MARK(7);
        if (PROF && rank == 0 && tid == 0) {
            // probe: how long does one global load of a fixed, L2-resident word take right after the barrier?
            unsigned int pv_;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(pv_) : "l"(&ctl->pad) : "memory");
            asm volatile("" ::"r"(pv_));
            CL_MARK(24);
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(pv_) : "l"(&ctl->pad) : "memory");
            asm volatile("" ::"r"(pv_));
            CL_MARK(25);
        }

        // ---------------------------------------------------------------- C: scan the qualifying units of the selected rows
        {
            if (tid == 0) { s_nsel = (int)*sel0; s_nunits = 0; }
            const bool merged = !first;
            const bool ymoved = merged && y < n;               // false when y was the last slot: nothing moved into it
            // Without helper clusters the main cluster writes columns x and y (and folds them into the unit keys) itself;
            // the stores ride behind the loads of the scan units, one chunk (64 scattered stores) per warp and unit.
            const int st_nch = (merged && HC == 0) ? (n + 31) >> 5 : 0;
            int st_lw = w;
            __syncthreads();
            const int nsel = s_nsel;
            // the unit keys of this half-warp's first staged row: a global round trip, issued before the bookkeeping below
            const int k0 = w * 2 + (lane >> 4);
            unsigned int kv0 = K32MAX;
            if (k0 < nsel && k0 < TILE && (lane & 15) < PARTS) kv0 = __ldcg(&Kmine[(size_t)(lane & 15) * KLD + t_row[k0]]);
            CL_MARK(17);
            // the moved row's unit key of the new column (its other units were copied in phase A; row y is not among the rows
            // the helpers fold).  A global atomic: issued here, not in phase B, where barrier 2's release fence would wait for
            // it; this merge's staging patches column x into row y's keys itself (below)
            if (ymoved && ((int)(((unsigned int)x >> 5) % CS)) == rank && tid == CT - 1)
                atomicMin(&Kmine[(size_t)(((int)(((unsigned int)x >> 5) / CS)) / UC) * KLD + y],
                          key_of((ld_peer_f64(&v_s[((int)(((unsigned int)y >> 5) / CS)) * 32 + (y & 31)], (int)(((unsigned int)y >> 5) % CS)) - ux) + C));
            if (rank == 0 && tid == 0) my_rows += (unsigned long long)nsel;
            const int nch = (n + 31) >> 5;
            const int lch = nch > rank ? (nch - rank + CS - 1) / CS : 0;      // local chunks holding columns < n
            const int parts = (lch + UC - 1) / UC;
            const int pdiv = parts > 0 ? parts : 1;
            // local chunk / lane of columns x and y when this CTA owns them: the scan takes those two columns from
            // v / f of the row, not from D (the helpers may still be writing them)
            const int xlw = (merged && ((int)(((unsigned int)x >> 5) % CS)) == rank) ? (int)(((unsigned int)x >> 5) / CS) : -1000000;
            const int ylw = (ymoved && ((int)(((unsigned int)y >> 5) % CS)) == rank) ? (int)(((unsigned int)y >> 5) / CS) : -1000000;
            const double uy_loc = ylw >= 0 ? u_s[ylw * 32 + (y & 31)] : 0.0;
            const double* const Rxs = R + (size_t)(iter & 1) * 2 * ld;
            const double* const Rys = Rxs + ld;
            // Units are refreshed a little EARLY: a unit whose bound will reach ub within about `slack_merges` merges at
            // the average drift so far is read now, while its row is staged anyway.  Without this every unit drifts to
            // the threshold on its own and costs its row a selection of its own (measured: 129 selected rows per merge
            // instead of 13); with it a row comes back when its runner-up does, as with whole-row rescans.
            const double slack = iter > 0 ? (double)((float)slack_merges * __fdividef((float)fabs(C), (float)iter)) : 0.0;   // (a heuristic: fp32 is plenty; C may be negative: never tighten)
            // (only the helper-less mode folds the new columns here)
            unsigned int* const Kxg = HC > 0 ? Kb : Kb + ((size_t)((x >= 0 ? x >> 5 : 0) % CS) * PARTS + ((x >= 0 ? x >> 5 : 0) / CS) / UC) * KLD;
            unsigned int* const Kyg = HC > 0 ? Kb : Kb + ((size_t)((y >= 0 ? y >> 5 : 0) % CS) * PARTS + ((y >= 0 ? y >> 5 : 0) / CS) / UC) * KLD;
            double bt = 1e300, bd = 0.0, bui = 0.0, buj = 0.0;
            int bi = -1, bj = -1;
            for (int t0 = 0; t0 < nsel || (t0 == 0 && st_nch > 0); t0 += TILE) {
                const int tn = nsel - t0 < TILE ? nsel - t0 : TILE;
                if (t0 > 0) {                       // (the first tile's counter was reset before the barrier above)
                    __syncthreads();
                    if (tid == 0) s_nunits = 0;
                    __syncthreads();
                }
                // stage the tile, one WARP per row (a warp whose 32 lanes gather from 32 different CTAs' shared memory pays for
                // them one after the other: ~1000 cycles per instruction; measured 8 k cycles for 34 rows staged by 34 threads):
                // row index, u, v, f of the row (owner's shared memory), empty combined minimum, and from this CTA's unit keys of
                // the row (lane p: unit p) which units must be read and what the others bound
                for (int k = w * 2 + (lane >> 4); k - (lane >> 4) < tn; k += 2 * NW) {
                    // half warp per row: lane hl of the half handles unit hl (parts <= 12)
                    const int hl = lane & 15;
                    const bool have = k < tn;
                    // the first tile was pushed into t_row / t_u / t_v / t_f by the row owners (phase B); later tiles (first
                    // search, bursts) come from the spill list in global memory and the owners' shared memory
                    int r = 0;
                    if (have) r = t0 == 0 ? t_row[k] : __ldcg(&sel_rows[t0 + k]);
                    const int ro = (int)(((unsigned int)r >> 5) % CS), rs = ((int)(((unsigned int)r >> 5) / CS)) * 32 + (r & 31);
                    const bool all = first || r == x || (dbg & 1);
                    unsigned int kv = K32MAX;
                    if (have && !all && hl < parts) kv = (t0 == 0 && k == k0) ? kv0 : __ldcg(&Kmine[(size_t)hl * KLD + r]);
                    double ur, vr, fr;
                    if (t0 == 0) { ur = have ? t_u[k] : 0.0; vr = have ? t_v[k] : 0.0; fr = have ? t_f[k] : 0.0; }
                    else {
                        double g = 0.0;
                        if (have && hl < 3) g = ld_peer_f64(hl == 0 ? &u_s[rs] : (hl == 1 ? &v_s[rs] : &f_s[rs]), ro);
                        ur = __shfl_sync(0xffffffffu, g, 0, 16); vr = __shfl_sync(0xffffffffu, g, 1, 16); fr = __shfl_sync(0xffffffffu, g, 2, 16);
                    }
                    if (PROF && k == 0) { CL_MARK(15); asm volatile("" ::"r"(kv)); CL_MARK(16); }
                    bool q = have && hl < parts;
                    if (!all) {
                        const bool patch = merged && r != y;        // (row y has no column y)
                        // the merge in flight: the helpers' fold of the two new columns (row y: the owner's atomic) may not have landed
                        if (merged && xlw >= 0 && hl == xlw / UC) { const unsigned int kxv = key_of((vr - ux) + C); if (kxv < kv) kv = kxv; }
                        if (patch && ylw >= 0 && hl == ylw / UC) { const unsigned int kyv = key_of((fr - uy_loc) + C); if (kyv < kv) kv = kyv; }
                        q = q && !(((double)dec_f32(kv) - C) - ur - marg > ub + slack);
                    }
                    const unsigned int qm = (__ballot_sync(0xffffffffu, q) >> (lane & 16)) & 0xffffu;
                    const unsigned int rest = half_min_u32((have && hl < parts && !q) ? kv : K32MAX, lane);
                    unsigned int base = 0;
                    if (have && hl == 0) {
                        if (t0 > 0) { t_row[k] = r; t_u[k] = ur; t_v[k] = vr; t_f[k] = fr; }
                        t_best[k] = KMAX;
                        t_qm[k] = qm;
                        t_k2[k] = rest;
                        if (qm) base = atomicAdd(&s_nunits, (unsigned int)__popc(qm));   // live units go to a compact list
                    }
                    base = __shfl_sync(0xffffffffu, base, 0, 16);
                    if (q) s_ulist[base + __popc(qm & ((1u << hl) - 1u))] = (unsigned short)((k << 5) | hl);
                    if (PROF && k == 0) CL_MARK(23);
                }
                __syncthreads();
                CL_MARK(8);
                const int units = (int)s_nunits;
                // one pass = one scan unit (UC column chunks of one selected row) + one chunk of column stores; a warp
                // that has run out of one of the two keeps going with the other (a dead unit loads nothing)
                // A warp keeps the loads of up to SB list entries in flight before it reduces any of them (one exposed memory
                // round trip per pass; at 512 threads 22 units per CTA and merge are one pass of 16 warps x 2).
                constexpr int SB = CT >= 1024 ? 1 : 2;
                for (int un0 = w; un0 < units || st_lw * CS + rank < st_nch; un0 += NW * SB) {
                    bool live[SB];
                    int uk[SB], upart[SB], ur_[SB];
                    double dv[SB][UC];
#pragma unroll
                    for (int b2 = 0; b2 < SB; b2++) {
                        const int un = un0 + b2 * NW;
                        live[b2] = un < units;
                        const unsigned int ent = live[b2] ? s_ulist[un] : 0u;
                        uk[b2] = (int)(ent >> 5); upart[b2] = (int)(ent & 31u);
                        const int r = t_row[uk[b2]];
                        ur_[b2] = r;
                        const int lw0 = live[b2] ? upart[b2] * UC : lch;
                        // rows x and y of this merge are still on their way into D (helpers): read them from the scratch pair
                        const double* row = (merged && r == x) ? Rxs : ((ymoved && r == y) ? Rys : D + (size_t)r * ld);
                        // (warp-uniform) a unit that lies wholly below n and does not hold the row's own column: bare loads
                        const bool whole = live[b2] && ((lw0 + UC - 1) * CS + rank) * 32 + 31 < n &&
                                           !(((int)(((unsigned int)r >> 5) % CS)) == rank && (unsigned int)(((int)(((unsigned int)r >> 5) / CS)) - lw0) < (unsigned int)UC);
                        if (whole) {
                            const double* const rp = row + (lw0 * CS + rank) * 32 + lane;
#pragma unroll
                            for (int q = 0; q < UC; q++) dv[b2][q] = __ldcg(rp + q * (CS * 32));
                        } else {
#pragma unroll
                            for (int q = 0; q < UC; q++) {
                                const int j = ((lw0 + q) * CS + rank) * 32 + lane;
                                dv[b2][q] = (live[b2] && j < n && j != r) ? __ldcg(&row[j]) : 1e300;   // 1e300: never a candidate
                            }
                        }
                    }
                    if (st_lw * CS + rank < st_nch) {
                        const int i = (st_lw * CS + rank) * 32 + lane;
                        if (i < n && i != x) {
                            const double vx = v_s[st_lw * 32 + lane];
                            D[(size_t)i * ld + x] = vx; D[(size_t)x * ld + i] = vx;
                            if (i != y) atomicMin(&Kxg[i], key_of((vx - ux) + C));
                        }
                        if (ymoved && i < n && i != y) {
                            // (i == x: the moved node's distance to the new node is v of row y, which the scratch row holds)
                            const double fy = i == x ? __ldcg(&Rys[x]) : f_s[st_lw * 32 + lane];
                            D[(size_t)i * ld + y] = fy; D[(size_t)y * ld + i] = fy;
                            if (i != x) atomicMin(&Kyg[i], key_of((fy - ld_peer_f64(&u_s[((int)(((unsigned int)y >> 5) / CS)) * 32 + (y & 31)], (int)(((unsigned int)y >> 5) % CS))) + C));
                        }
                        st_lw += NW;
                    }
#pragma unroll
                    for (int b2 = 0; b2 < SB; b2++) {
                        if (!live[b2]) continue;
                        const int k = uk[b2], part = upart[b2], r = ur_[b2], lw0 = part * UC;
                        const double ur = t_u[k];
                        const bool patch = merged && r != x && r != y;
                        // columns x and y of an old row come from v / f of the row (the helpers may still be writing D); only the
                        // one or two units that hold those columns pay for the test (warp-uniform)
                        const bool near = patch && ((unsigned int)(xlw - lw0) < (unsigned int)UC || (unsigned int)(ylw - lw0) < (unsigned int)UC);
                        const int xq = (near && lane == (x & 31)) ? xlw - lw0 : -1;
                        const int yq = (near && lane == (y & 31)) ? ylw - lw0 : -1;
                        // A lane's columns of one unit differ by multiples of 32 * CS (a multiple of 256), so the reference
                        // order within the unit is plain ascending j: the first strict minimum is the right one.
                        double lm1 = 1e300, lm2 = 1e300, ut = 1e300;   // lm1/lm2: two smallest d - u_j; ut: smallest d - u_r - u_j
                        int uq = 0, lq = 0;
                        const double* const up = u_s + lw0 * 32 + lane;   // (chunks past lch read neighbouring arrays: their d is 1e300)
                        auto unit_pass = [&](auto patched) {
#pragma unroll
                            for (int q = 0; q < UC; q++) {
                                double d = dv[b2][q];
                                if (decltype(patched)::value) {
                                    if (q == xq) d = t_v[k];
                                    if (q == yq) d = t_f[k];
                                }
                                const double uj = up[q * 32];
                                const double t = (d - ur) - uj;
                                const double mv = d - uj;
                                if (mv < lm1) { lm2 = lm1; lm1 = mv; lq = q; } else lm2 = fmin(lm2, mv);
                                if (t < ut) { ut = t; uq = q; }
                            }
                        };
                        if (near) unit_pass(std::true_type{}); else unit_pass(std::false_type{});
                        // (units of one row may reach a lane in any order: ties within the row go through the reference order too)
                        if (ut < 10000.0 && (ut < bt || (ut == bt && tie_before(r, ((lw0 + uq) * CS + rank) * 32 + lane, bi, bj, n)))) {
                            // rare: fetch the winning column's d and u_j again (the loop carries only the value and the index)
                            double ud = 0.0;
#pragma unroll
                            for (int q = 0; q < UC; q++) if (q == uq) ud = dv[b2][q];
                            if (uq == xq) ud = t_v[k];
                            if (uq == yq) ud = t_f[k];
                            bt = ut; bi = r; bj = ((lw0 + uq) * CS + rank) * 32 + lane; bd = ud; bui = ur; buj = up[uq * 32];
                        }
                        // unit minimum and runner-up -> the unit's key, the CTA's (minimum, column) and runner-up of the row.  An
                        // atomic that loses to (or displaces) the standing minimum demotes the loser to the runner-up.
                        const unsigned int key1 = lm1 < 1e299 ? key_of(lm1 + C) : K32MAX;
                        const unsigned int k1w = __reduce_min_sync(0xffffffffu, key1);
                        if (lane == 0) Kmine[(size_t)part * KLD + r] = k1w;
                        if (k1w != K32MAX) {
                            const int wl = __ffs(__ballot_sync(0xffffffffu, key1 == k1w)) - 1;
                            const int jw = __shfl_sync(0xffffffffu, ((lw0 + lq) * CS + rank) * 32 + lane, wl);
                            const unsigned int key2 = lane == wl ? (lm2 < 1e299 ? key_of(lm2 + C) : K32MAX) : key1;
                            const unsigned int k2w = __reduce_min_sync(0xffffffffu, key2);
                            if (lane == 0) {
                                const unsigned long long mine = ((unsigned long long)k1w << 32) | (unsigned int)jw;
                                const unsigned long long old = atomicMin(&t_best[k], mine);
                                const unsigned int demoted = mine < old ? (unsigned int)(old >> 32) : k1w;
                                atomicMin(&t_k2[k], demoted < k2w ? demoted : k2w);
                            }
                        }
                        if (lane == 0) my_units++;
                    }
                }
                __syncthreads();
                CL_MARK(9);
                // helpers' progress, sampled here (well ahead of barrier 3, whose release fence waits for outstanding loads)
                // and looked at in the pick
                if (HC > 0 && lane == 0) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(done_early) : "l"(&ctl->done) : "memory");
                // what this CTA found for each staged row goes to the row owner's slot (one warp per row, see the staging loop)
                for (int k = w; k < tn; k += NW) {
                    if (lane != 0) continue;
                    const unsigned long long best = t_best[k];
                    const int r = t_row[k];
                    st_peer_v4(&slot_s[k * MAXCS + rank], (int)(((unsigned int)r >> 5) % CS), (unsigned int)(best >> 32), (unsigned int)best, t_k2[k], 0u);
                }
                if (nsel > TILE) {
                    // rare (first search, bursts): two extra cluster barriers per tile -- the owners resolve this tile before
                    // anyone pushes the next one
                    cluster.sync();
                    resolve(tn, merged, x, y, n, iter, false);
                    cluster.sync();
                }
            }
            // (Prefetching the next merge's row n - 1 into L2 here measured 0.1 us per merge SLOWER than not doing it.)
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
        CL_MARK(10);
        cluster.sync();
        CL_MARK(11);

        // ---------------------------------------------------------------- D: pick (identical in every CTA)
        {
            // No CTA barrier in this stretch (each costs ~350 cycles here): every warp checks the helpers, finds the winner and
            // keeps its own pool slots by itself, and phase A follows without a join.
            if (HC > 0 && iter > 0 && lane == 0) {
                // the next update reads whole rows: the helpers must have finished the columns of the previous merge.  The
                // counter was sampled before barrier 3 (a global round trip that would otherwise sit here); only if the helpers
                // were not done by then does the lane poll
                const unsigned int want = (unsigned int)iter * (unsigned int)HC;
                unsigned int dn = done_early;
                while ((int)(dn - want) < 0) {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(dn) : "l"(&ctl->done) : "memory");
                }
            }
            __syncwarp();
            CL_MARK(26);
            // the cluster's winner (recs[] is complete after barrier 3)
            const int src = lane < CS ? lane : 0;
            const int ci = lane < CS ? recs[src].i : -1;
            const int wl = warp_best_lane(recs[src].t, ci, recs[src].j, n);
            CL_MARK(29);
            const CRec* const win = &recs[wl < 0 ? 0 : wl];
            const int wi = win->i, wj = win->j;
            const double wd = win->d, wui = win->ui, wuj = win->uj;
            double uxo, uyo;
            if (wi < wj) { x = wi; y = wj; uxo = wui; uyo = wuj; } else { x = wj; y = wi; uxo = wuj; uyo = wui; }
            dxy = wd;
            const int last_ = n - 1;
            {
                // candidate pool: the lane pair that re-evaluates slot pc in phase A also keeps it -- drop a pair that lost an
                // end, rename `last`, then take this scan's winner of warp q (still in wrec[1..NW-1]; slot (pool_head + q) %
                // CPOOL) if it is better than what the slot holds
                constexpr int PCW = CPOOL / NW;
                const int pc = w * PCW + (lane >> 1);
                if (lane < 2 * PCW && !(lane & 1)) {
                    int pi = pool_i[pc], pj = pool_j[pc];
                    if (pi >= 0) {
                        if (pi == x || pi == y || pj == x || pj == y) pi = -1;
                        else {
                            if (pi == last_) pi = y;
                            if (pj == last_) pj = y;
                        }
                    }
                    const int q = (pc - pool_head + CPOOL) % CPOOL;
                    if (q > 0 && q < NW) {
                        const int mi = wrec[q].i, mj = wrec[q].j;
                        const double mt = wrec[q].t;
                        // keep the best: a slot is overwritten only by a candidate that is better than its last evaluation
                        if (mi >= 0 && mi != x && mi != y && mj != x && mj != y && (pi < 0 || mt < pool_t[pc])) {
                            pi = mi == last_ ? y : mi;
                            pj = mj == last_ ? y : mj;
                            pool_d[pc] = wrec[q].d;
                            pool_t[pc] = mt;
                        }
                    }
                    pool_i[pc] = pi; pool_j[pc] = pj;
                }
                __syncwarp();
            }
            pool_head = (pool_head + NW) % CPOOL;
            CL_MARK(30);
            if (rank == 0 && tid == 0) {
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
            CL_MARK(27);
            iter++;
        }
        first = false;
    }
    if (rank == 0 && tid == 0) {
        if (HC > 0) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(&ctl->bell), "l"(KMAX) : "memory");
        stats->iters = (unsigned long long)iter;
        for (int k = 0; k < 32; k++) stats->cyc[k] = s_cyc[k];
        unsigned long long ns_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_end));
        stats->t_ns = ns_end - ns_begin;
        stats->t_cycles = (unsigned long long)(clock64() - cyc_begin);
    }
    if (lane == 0 && (my_units || my_rows)) {
        atomicAdd(&stats->units_scanned, my_units);
        atomicAdd(&stats->bytes_scanned, my_units * (unsigned long long)(UC * 256));
        if (my_rows) atomicAdd(&stats->rows_scanned, my_rows);
    }
    cluster.sync();   // no CTA may exit while peers can still read its shared memory
}

template <int CS, int CT, int UC, bool PROF>
static int launch_cluster(dipb_ctx* c, int n, void** args, int* LS_out, int* HC, int* PARTS, int max_helper_clusters, bool* ok) {
    const int chunks = (n + 31) / 32;
    const int LS = ((chunks + CS - 1) / CS) * 32;
    *PARTS = ((chunks + CS - 1) / CS + UC - 1) / UC;      // scan units per row and CTA
    if (*PARTS > MAXPARTS) { *ok = false; return 0; }
    const size_t smem = cluster_smem_bytes(LS, chunks);
    auto kern = nj_cluster_kernel<CS, CT, UC, PROF>;
    *ok = false;
    *LS_out = LS;
    if (smem > 220u * 1024u) return 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (CS > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS); cfg.blockDim = dim3(CT); cfg.dynamicSmemBytes = smem; cfg.stream = c->stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative;     // helpers spin on the main cluster: all clusters resident or no launch
    at[1].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = 0;
    const unsigned long long occ_key = ((unsigned long long)CS << 56) | ((unsigned long long)CT << 40) | ((unsigned long long)PROF << 39) | (unsigned long long)smem;
    if (c->nj_occ_key == occ_key) nclusters = c->nj_occ_clusters;
    else {
        if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess || nclusters < 1) { cudaGetLastError(); return 0; }
        c->nj_occ_key = occ_key; c->nj_occ_clusters = nclusters;
    }
    // helper clusters spin on a doorbell of the main cluster: only as many as are co-resident with it
    int helpers = nclusters - 1 < max_helper_clusters ? nclusters - 1 : max_helper_clusters;
    if (helpers < 0) helpers = 0;
    cudaError_t e = cudaErrorUnknown;
    for (; helpers >= 0; helpers = helpers > 0 ? 0 : -1) {
        *HC = helpers * CS;
        cfg.gridDim = dim3(CS * (1 + helpers));
        cfg.numAttrs = helpers > 0 ? 2 : 1;
        e = cudaLaunchKernelExC(&cfg, (const void*)kern, args);
        if (e == cudaSuccess) break;
        cudaGetLastError();                        // e.g. the GPU is shared right now: retry without helpers
    }
    if (e != cudaSuccess) { set_error("nj_cluster: launch failed: %s", cudaGetErrorString(e)); return DIPB_E_CUDA; }
    *ok = true;
    return 0;
}

bool nj_cluster_fits(int n) {
    // shared memory of the 16-CTA layout (per-row state of n / 16 rows + the staging tile) within 200 KB; when the
    // device cannot co-schedule 16 CTAs and the 8-CTA layout is too large, nj_cluster_loop reports DIPB_E_UNSUPPORTED
    const int chunks = (n + 31) / 32;
    const int LS = ((chunks + 15) / 16) * 32;
    // (the doorbell packs indices in 17 bits; the unit list of a staged tile holds 12 units per row: 49 152 tips)
    return n < 131072 && cluster_smem_bytes(LS, chunks) <= 200u * 1024u && (LS / 32 + 7) / 8 <= 12;   // (half-warp staging: at most 16 units per row and CTA)
}

int nj_cluster_loop(dipb_matrix* m, double* U, double* u, int* realID, int32_t* c0, int32_t* c1, double* l0, double* l1) {
    dipb_ctx* c = m->ctx;
    const int n = m->n;
    auto t_host0 = std::chrono::steady_clock::now();
    double t_host[5] = {0, 0, 0, 0, 0};   // DIPB_NJ_PROFILE: allocations + margin scale, launch, kernel, replay, frees (ms)
    auto lap = [&](int k) {
        auto now = std::chrono::steady_clock::now();
        t_host[k] += std::chrono::duration<double, std::milli>(now - t_host0).count();
        t_host0 = now;
    };
    int* sel = nullptr;
    CStats* stats = nullptr;
    NJCtl* ctl = nullptr;
    int2* log_xy = nullptr;
    double2* log_bl = nullptr;
    unsigned int* Kb = nullptr;
    double* R = nullptr;              // scratch rows x and y of the merge in flight, double buffered by merge parity
    // unit keys: n rows x (CS * PARTS) keys; CS * PARTS <= chunks / UC + 2 * CS for either cluster size (UC >= 4)
    const size_t kb_bytes = sizeof(unsigned int) * (size_t)((n + 31) & ~31) * ((size_t)((n + 31) / 32) / 4 + 32);
    DIPB_CUDA(pool_alloc(c, (void**)&Kb, kb_bytes));
    DIPB_CUDA(cudaMemsetAsync(Kb, 0xff, kb_bytes, c->stream));
    DIPB_CUDA(pool_alloc(c, (void**)&R, sizeof(double) * 4 * (size_t)n));
    DIPB_CUDA(cudaMemsetAsync(R, 0, sizeof(double) * 4 * (size_t)n, c->stream));
    DIPB_CUDA(pool_alloc(c, (void**)&sel, sizeof(int) * n));
    DIPB_CUDA(pool_alloc(c, (void**)&stats, sizeof(CStats)));
    DIPB_CUDA(pool_alloc(c, (void**)&ctl, sizeof(NJCtl)));
    DIPB_CUDA(pool_alloc(c, (void**)&log_xy, sizeof(int2) * n));
    DIPB_CUDA(pool_alloc(c, (void**)&log_bl, sizeof(double2) * n));
    DIPB_CUDA(cudaMemsetAsync(stats, 0, sizeof(CStats), c->stream));
    DIPB_CUDA(cudaMemsetAsync(ctl, 0, sizeof(NJCtl), c->stream));
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
    const int profile = getenv("DIPB_NJ_PROFILE") ? 1 : 0;
    lap(0);
    const char* force = getenv("DIPB_NJ_CLUSTER");   // 8 or 16; default: 16 when the device can co-schedule it
    const int want = force ? atoi(force) : 16;
    const char* hp = getenv("DIPB_NJ_HELPERS");      // helper clusters for the column stores (default: all that fit, at most 7)
    const int max_helpers = hp ? atoi(hp) : 7;
    int used = 0, HC = 0, LS = 0, PARTS = 0;
    bool ok = false;
    int rc = 0;
    int dbg = getenv("DIPB_NJ_DBG") ? atoi(getenv("DIPB_NJ_DBG")) : 0;   // bit 0: every unit of a selected row is read (full rescans)
    double slack_merges = getenv("DIPB_NJ_SLACK") ? atof(getenv("DIPB_NJ_SLACK")) : 1024.0;
    void* args[] = {&Dp, &ld, &U, &u, &sel, &stats, &log_xy, &log_bl, &n_total, &LS, &dmax, &ctl, &HC, &Kb, &PARTS, &dbg, &slack_merges, &R};
    // 1024 threads, 8 loads in flight per lane: measured best of {512, 1024} x {8, 16} (profiles/r1_nj_cluster_tuning.json)
    const char* e_uc = getenv("DIPB_NJ_UC");
    const int uc = e_uc ? atoi(e_uc) : 8;
    const char* e_ct = getenv("DIPB_NJ_THREADS");    // threads per CTA: 256, 512 (default) or 1024
    const int ct = e_ct ? atoi(e_ct) : 512;
    (void)uc;
    if (want >= 16) {
        if (ct >= 1024) rc = profile ? launch_cluster<16, 1024, 8, true>(c, n, args, &LS, &HC, &PARTS, max_helpers, &ok) : launch_cluster<16, 1024, 8, false>(c, n, args, &LS, &HC, &PARTS, max_helpers, &ok);
        else if (ct >= 512) rc = profile ? launch_cluster<16, 512, 8, true>(c, n, args, &LS, &HC, &PARTS, max_helpers, &ok) : launch_cluster<16, 512, 8, false>(c, n, args, &LS, &HC, &PARTS, max_helpers, &ok);
        else rc = profile ? launch_cluster<16, 256, 8, true>(c, n, args, &LS, &HC, &PARTS, max_helpers, &ok) : launch_cluster<16, 256, 8, false>(c, n, args, &LS, &HC, &PARTS, max_helpers, &ok);
        used = 16;
    }
    if (!rc && !ok) {
        rc = profile ? launch_cluster<8, 1024, 8, true>(c, n, args, &LS, &HC, &PARTS, max_helpers, &ok) : launch_cluster<8, 1024, 8, false>(c, n, args, &LS, &HC, &PARTS, max_helpers, &ok);
        used = 8;
    }
    if (!rc && !ok) { set_error("nj_cluster: no cluster configuration fits this device"); rc = DIPB_E_UNSUPPORTED; }
    if (rc) { pool_free(c, sel); pool_free(c, stats); pool_free(c, ctl); pool_free(c, log_xy); pool_free(c, log_bl); pool_free(c, Kb); pool_free(c, R); return rc; }
    c->launches++;
    lap(1);
    DIPB_CUDA(cudaStreamSynchronize(c->stream));
    lap(2);
    {
        // replay of realID / tree bookkeeping (src/neighborJoining.cu:233-237) from the device log
        const int iters = n - 2;
        std::vector<int2> hxy(iters);
        std::vector<double2> hbl(iters);
        std::vector<int> rid(n), hc0(n), hc1(n);
        std::vector<double> hl0(n), hl1(n);
        DIPB_CUDA(cudaMemcpyAsync(hxy.data(), log_xy, sizeof(int2) * iters, cudaMemcpyDeviceToHost, c->stream));
        DIPB_CUDA(cudaMemcpyAsync(hbl.data(), log_bl, sizeof(double2) * iters, cudaMemcpyDeviceToHost, c->stream));
        DIPB_CUDA(cudaStreamSynchronize(c->stream));
        for (int i = 0; i < n; i++) rid[i] = i;
        int id = n;
        for (int it = 0; it < iters; it++) {
            const int xx = hxy[it].x, yy = hxy[it].y, act = n - it;
            hc0[it] = rid[xx]; hl0[it] = hbl[it].x;
            hc1[it] = rid[yy]; hl1[it] = hbl[it].y;
            rid[xx] = id++; rid[yy] = rid[act - 1];
        }
        // uploads ride the context's stream (nj_finish_kernel reads realID there); the host vectors live until the sync
        DIPB_CUDA(cudaMemcpyAsync(c0, hc0.data(), sizeof(int32_t) * iters, cudaMemcpyHostToDevice, c->stream));
        DIPB_CUDA(cudaMemcpyAsync(c1, hc1.data(), sizeof(int32_t) * iters, cudaMemcpyHostToDevice, c->stream));
        DIPB_CUDA(cudaMemcpyAsync(l0, hl0.data(), sizeof(double) * iters, cudaMemcpyHostToDevice, c->stream));
        DIPB_CUDA(cudaMemcpyAsync(l1, hl1.data(), sizeof(double) * iters, cudaMemcpyHostToDevice, c->stream));
        DIPB_CUDA(cudaMemcpyAsync(realID, rid.data(), sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
        DIPB_CUDA(cudaStreamSynchronize(c->stream));
    }
    lap(3);
    CStats hs;
    DIPB_CUDA(cudaMemcpyAsync(&hs, stats, sizeof(hs), cudaMemcpyDeviceToHost, c->stream));
    DIPB_CUDA(cudaStreamSynchronize(c->stream));
    c->nj_rows_scanned = hs.rows_scanned;
    c->nj_iterations = hs.iters;
    c->nj_bytes_scanned = hs.bytes_scanned;
    if (profile) {
        fprintf(stderr, "[nj_cluster]   stage detail (rank 0 warp 0): to first sync %.0f, key load issued %.0f, key arrived %.0f, unit test + list %.0f cyc/iter (then the CTA barrier)\n",
                hs.cyc[17] / (double)hs.iters, hs.cyc[15] / (double)hs.iters, hs.cyc[16] / (double)hs.iters, hs.cyc[23] / (double)hs.iters);
        fprintf(stderr, "[nj_cluster]   A detail (rank 0 warp 0): loads issued %.0f, resolve %.0f, chunks processed %.0f, pool + drift %.0f, wait for the CTA %.0f cyc/iter\n",
                hs.cyc[18] / (double)hs.iters, hs.cyc[19] / (double)hs.iters, hs.cyc[20] / (double)hs.iters, hs.cyc[21] / (double)hs.iters, hs.cyc[22] / (double)hs.iters);
        fprintf(stderr, "[nj_cluster]   pick detail (rank 0 thread 0): helpers done %.0f, winner of the cluster %.0f, pool slots %.0f, log %.0f cyc/iter (the rest is in 'D pick')\n",
                hs.cyc[26] / (double)hs.iters, hs.cyc[29] / (double)hs.iters, hs.cyc[30] / (double)hs.iters, hs.cyc[27] / (double)hs.iters);
        fprintf(stderr, "[nj_cluster]   probe after barrier 2: first load of a fixed word %.0f, second %.0f cyc\n", hs.cyc[24] / (double)hs.iters, hs.cyc[25] / (double)hs.iters);
        const char* nm[12] = {"D pick + pool", "A update + pool eval + push", "barrier 1", "B canonical sum + bell", "B upper bound", "(unused)",
                              "B fold + select", "barrier 2", "C stage tile + unit keys", "C scan units", "C keys + reduce + publish", "barrier 3"};
        fprintf(stderr, "[nj_cluster] main loop: %.1f ms, %.3f G cycles -> SM clock %.0f MHz while it ran\n", hs.t_ns * 1e-6, hs.t_cycles * 1e-9,
                hs.t_ns ? 1e3 * (double)hs.t_cycles / (double)hs.t_ns : 0.0);
        fprintf(stderr, "[nj_cluster] rescans of rank 0's rows: %llu without a tracked partner, %llu partner is a contender, %llu runner-up bound reached ub\n",
                hs.cyc[12], hs.cyc[13], hs.cyc[14]);
        double tot = 0;
        for (int k = 0; k < 12; k++) tot += (double)hs.cyc[k];
        fprintf(stderr, "[nj_cluster] n=%d cluster=%d helper_ctas=%d iters=%llu rows_selected=%llu (%.1f/iter) units_scanned=%llu (%.1f/iter, %d per full row)\n", n, used, HC, hs.iters,
                hs.rows_scanned, hs.iters ? (double)hs.rows_scanned / hs.iters : 0.0, hs.units_scanned, hs.iters ? (double)hs.units_scanned / hs.iters : 0.0, used * PARTS);
        for (int k = 0; k < 12; k++)
            fprintf(stderr, "[nj_cluster]   %-26s %8.0f cyc/iter  %5.1f%%\n", nm[k], hs.iters ? hs.cyc[k] / (double)hs.iters : 0.0,
                    tot > 0 ? 100.0 * hs.cyc[k] / tot : 0.0);
    }
    pool_free(c, sel); pool_free(c, stats); pool_free(c, ctl); pool_free(c, log_xy); pool_free(c, log_bl); pool_free(c, Kb); pool_free(c, R);
    lap(4);
    if (profile) fprintf(stderr, "[nj_cluster] host ms: alloc+scale %.1f, launch %.1f, kernel wait %.1f, replay %.1f, profile print + frees %.1f\n", t_host[0], t_host[1], t_host[2], t_host[3], t_host[4]);
    return 0;
}

}  // namespace dipb
