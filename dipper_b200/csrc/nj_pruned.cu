// K4 (pruned): exact neighbor-joining search with lower-bound row pruning (sm_100a).
//
// Same result as the exhaustive search in nj.cu / the reference's findMinDist +
// thrust::min_element (src/neighborJoining.cu:117-148,214), including the tie-break,
// but only rows that can still hold the minimum are read.  The reference reads
// n^2 doubles per iteration (72 TB over a 30 000-tip run).
//
// Bound.  For every active row i we keep
//     K[i] >= -inf  with   min_{j != i} (d(i,j) - u_j(now)) >= K[i] - C(now)
// where u_j = U[j]/(n-2) and C accumulates, per merge, max_j (u_j(new) - u_j(old)) over
// surviving columns.  A rescan of row i sets K[i] = min_j (d(i,j) - u_j) + C exactly;
// the column created by a merge is folded into every K explicitly.  Then
//     min_j q(i,j) >= K[i] - C - u_i =: lb_i
// and with ub = the smallest exactly-evaluated candidate we know (each row's last
// argmin partner re-evaluated with the current u, plus the new column), every row with
// lb_i - margin > ub is skipped: all of its candidates are strictly larger than the
// true minimum, so neither the argmin nor any tie is lost.  Rows that pass are
// rescanned with the reference's exact expression (d - u_i) - u_j and tie order.
//
// Execution: ONE persistent cooperative kernel runs all N-2 iterations (one CTA per
// SM, 4 grid barriers per iteration); no host round trips.  Phases per iteration:
//   A  merge update of rows/columns x,y (as updateDisMatrix :161-194), deterministic U,
//      max u-drift, partner/K bookkeeping for the moved row
//   B  fold the new column into K, re-evaluate partners -> per-CTA upper bounds
//   C1 select rows with lb <= ub into a work list
//   C2 all CTAs scan (row, 4096-column chunk) items of the list
//   D  every CTA reduces the per-CTA winners to the same (x, y); CTA 0 records the tree
#include <cstdlib>
#include <vector>
#include "common.cuh"
#include "nj.cuh"
#include "nj_bound.cuh"

namespace dipb {

namespace {

constexpr int PT = 1024;        // threads per CTA
constexpr int EPT = 8;          // columns per thread per scan item
constexpr int CHUNK = EPT * PT; // columns per scan item
constexpr int POOL = 512;       // carried candidate pairs (upper bound of the minimum without a grid reduction)

struct PCand {
    double t;   // candidate value, 1e300 = empty
    int i, j;
    double d, ui, uj;
};

struct PShared {
    unsigned int bar_count;
    unsigned int bar_gen;
    unsigned int sel_count;
    unsigned int pad;
    unsigned long long rows_scanned, iters;
    unsigned long long cyc[8];   // CTA 0 cycle counters per phase: A, bar, B, bar, C1, bar, C2+bar, D
};

// Grid barrier on one monotonically increasing counter: barrier k is complete when the
// counter reaches k * nblocks.  One release-fence + atomic per CTA, relaxed polling by a
// single thread (all-to-all flag polling was measured slower: it floods L2 while other
// CTAs still work).  Requires all CTAs co-resident (cooperative launch).
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int nblocks, unsigned int& gen) {
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

__device__ __forceinline__ double block_min(double v, double* sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, s));
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double r = sh[lane];
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) r = fmin(r, __shfl_xor_sync(0xffffffffu, r, s));
    __syncthreads();
    return r;
}
__device__ __forceinline__ double block_max(double v, double* sh) { return -block_min(-v, sh); }

}  // namespace

__global__ void __launch_bounds__(PT, 1)
nj_pruned_kernel(double* __restrict__ D, size_t ld, double* __restrict__ U, double* __restrict__ u,
                 unsigned long long* __restrict__ K, double* __restrict__ partial_sum,
                 double* __restrict__ partial_max, int* __restrict__ sel_rows,
                 PCand* __restrict__ cta_best, PShared* ps, unsigned int* __restrict__ bar_flags,
                 int2* __restrict__ log_xy, double2* __restrict__ log_bl, int n_total, double dmax) {
    __shared__ double sh[32];
    __shared__ PCand shc[32];
    __shared__ double s_ux, s_C, s_ub;
    __shared__ double s_part[PT];
    __shared__ int pool_i[POOL], pool_j[POOL];
    __shared__ int s_pool_head;
    const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    int n = n_total;
    double C = 0.0;
    int next_id = n_total;
    int x = -1, y = -1;
    double dxy = 0.0;
    bool first = true;
    unsigned long long my_rows = 0;
    unsigned int bar_gen = 0;
    int iter = 0;

    for (int p = tid; p < POOL; p += PT) { pool_i[p] = -1; pool_j[p] = -1; }
    if (tid == 0) s_pool_head = 0;
    __syncthreads();
    long long tmark = clock64();
#define PHASE_MARK(k)                                                   \
    do {                                                                \
        if (cta == 0 && tid == 0) {                                     \
            long long now__ = clock64();                                \
            ps->cyc[k] += (unsigned long long)(now__ - tmark);          \
            tmark = now__;                                              \
        }                                                               \
    } while (0)

    while (n > 2) {
        double ub = 1e300;
        if (!first) {
            PHASE_MARK(7);
            // ------------------------------------------------ Phase A: merge update
            const int last = n - 1;
            const double den_new = (double)(n - 3);
            for (int blk = cta; blk * PT < last; blk += G) {
                const int i = blk * PT + tid;
                double slot = 0.0, dlt = -1e300;
                if (i < last && i != x) {
                    if (i != y) {
                        double a = __ldcg(&D[(size_t)x * ld + i]), b = __ldcg(&D[(size_t)y * ld + i]);
                        double val = (a + b - dxy) * 0.5;
                        double far = __ldcg(&D[(size_t)last * ld + i]);
                        double Ui = __ldcg(&U[i]);
                        Ui += -a - b + val;
                        U[i] = Ui;
                        D[(size_t)x * ld + i] = val;
                        D[(size_t)i * ld + x] = val;
                        D[(size_t)y * ld + i] = far;
                        D[(size_t)i * ld + y] = far;
                        slot = val;
                        if (n > 3) {
                            double un = Ui / den_new;
                            dlt = un - __ldcg(&u[i]);
                            u[i] = un;
                        }
                    } else {
                        double a = __ldcg(&D[(size_t)x * ld + last]), b = __ldcg(&D[(size_t)y * ld + last]);
                        double val = (a + b - dxy) * 0.5;
                        double uy = __ldcg(&U[last]);
                        uy += -a - b + val;
                        U[y] = uy;
                        D[(size_t)x * ld + y] = val;
                        D[(size_t)y * ld + x] = val;
                        slot = val;
                        if (n > 3) {
                            double un = uy / den_new;
                            dlt = un - __ldcg(&u[last]);
                            u[y] = un;
                        }
                        K[y] = __ldcg(&K[last]);
                    }
                }
                // canonical block sum (same order as nj.cu / the oracle) and max drift, one exchange
                double v = warp_tree_sum(slot);
                double mxw = dlt;
#pragma unroll
                for (int s = 16; s >= 1; s >>= 1) mxw = fmax(mxw, __shfl_xor_sync(0xffffffffu, mxw, s));
                if (lane == 0) { sh[w] = v; s_part[w] = mxw; }
                __syncthreads();
                if (w == 0) {
                    double g = warp_tree_sum(sh[lane]);
                    double mx = s_part[lane];
#pragma unroll
                    for (int s = 16; s >= 1; s >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, s));
                    if (lane == 0) { partial_sum[blk] = g; partial_max[blk] = mx; }
                }
                __syncthreads();
            }
            PHASE_MARK(0);
            grid_barrier(bar_flags, G, bar_gen);
            PHASE_MARK(1);

            // ------------------------------------------------ Phase B: finalise U[x], fold column x, upper bound
            n = last;
            if (n <= 2) break;
            {
                const int nb = (last + PT - 1) / PT;   // <= 1024 blocks (n <= 2^20) fit one pass
                double pm = tid < nb ? __ldcg(&partial_max[tid]) : -1e300;
                if (tid < nb) s_part[tid] = __ldcg(&partial_sum[tid]);
                double mx = block_max(pm, sh);
                if (tid == 0) {
                    double acc = 0.0;
                    for (int b = 0; b < nb; b++) acc += s_part[b];   // ascending block order (canonical sum)
                    s_ux = acc / (double)(n - 2);
                    s_C = C + mx;
                    if (cta == 0) { U[x] = acc; u[x] = s_ux; }
                }
            }
            __syncthreads();
            const double ux = s_ux;
            C = s_C;
            // upper bound of the minimum: carried pairs re-evaluated exactly with the post-merge u
            // (no pair touches x or y; every u[] entry was rewritten in phase A)
            {
                double pv = 1e300;
                if (tid < POOL && pool_i[tid] >= 0) {
                    const int pi = pool_i[tid], pj = pool_j[tid];
                    pv = (__ldcg(&D[(size_t)pi * ld + pj]) - __ldcg(&u[pi])) - __ldcg(&u[pj]);
                }
                ub = block_min(pv, sh);
            }
            // fold the new column into K and select the rows of this CTA's segment
            {
                const double margin = 1e-9 * (4.0 * dmax + fabs(C));
                const int seg = (n + G - 1) / G;
                for (int k = tid; k < seg; k += PT) {
                    int i = cta * seg + k;
                    if (i >= n) break;
                    bool take = (i == x);   // the new row is always rescanned
                    if (!take) {
                        double ui = __ldcg(&u[i]);
                        double dix = __ldcg(&D[(size_t)i * ld + x]);
                        unsigned long long kc = enc_f64((dix - ux) + C);
                        unsigned long long ko = __ldcg(&K[i]);
                        if (kc < ko) { ko = kc; K[i] = kc; }
                        double lb = (dec_f64(ko) - C) - ui - margin;
                        take = (ko == 0ull) || !(lb > ub);
                    }
                    if (take) {
                        unsigned int pos = atomicAdd(&ps->sel_count, 1u);
                        sel_rows[pos] = i;
                        K[i] = 0xffffffffffffffffull;   // reset, the scan atomically lowers it
                    }
                }
            }
            PHASE_MARK(2);
        } else {
            // first search: every row
            const int seg = (n + G - 1) / G;
            for (int k = tid; k < seg; k += PT) {
                int i = cta * seg + k;
                if (i >= n) break;
                unsigned int pos = atomicAdd(&ps->sel_count, 1u);
                sel_rows[pos] = i;
                K[i] = 0xffffffffffffffffull;
            }
        }
        PHASE_MARK(4);
        grid_barrier(bar_flags, G, bar_gen);
        PHASE_MARK(5);

        // ---------------------------------------------------- Phase C2: scan (row, chunk) items
        {
            const int nsel = (int)*((volatile unsigned int*)&ps->sel_count);
            const int nchunk = (n + CHUNK - 1) / CHUNK;
            const long long items = (long long)nsel * nchunk;
            double bt = 1e300; int bi = 0, bj = 0; double bd = 0, bui = 0, buj = 0;
            for (long long it = cta; it < items; it += G) {
                const int r = __ldcg(&sel_rows[(int)(it / nchunk)]);
                const int c0 = (int)(it % nchunk) * CHUNK;
                const double ur = __ldcg(&u[r]);
                const double* row = D + (size_t)r * ld;
                double lt = 1e300, lm = 1e300, ld_ = 0, luj = 0; int lj = -1;
                double dv[EPT], uv[EPT];
#pragma unroll
                for (int q = 0; q < EPT; q++) {
                    int j = c0 + q * PT + tid;
                    bool ok = j < n && j != r;
                    dv[q] = ok ? __ldcg(&row[j]) : 0.0;
                    uv[q] = ok ? __ldcg(&u[j]) : 0.0;
                }
#pragma unroll
                for (int q = 0; q < EPT; q++) {
                    int j = c0 + q * PT + tid;
                    if (j < n && j != r) {
                        double d = dv[q];
                        double uj = uv[q];
                        double t = (d - ur) - uj;
                        double m = d - uj;
                        lm = fmin(lm, m);
                        if (t < 10000.0 && (t < lt || (t == lt && (((j & 255) < (lj & 255)) || ((j & 255) == (lj & 255) && j < lj))))) {
                            lt = t; lj = j; ld_ = d; luj = uj;
                        }
                    }
                }
                // block-reduce: min m, and best (t, j) in the reference order within row r
                double rt = lt; int rj = lj; double rd = ld_, ruj = luj, rm = lm;
#pragma unroll
                for (int s = 16; s >= 1; s >>= 1) {
                    double ot = __shfl_xor_sync(0xffffffffu, rt, s);
                    int oj = __shfl_xor_sync(0xffffffffu, rj, s);
                    double od = __shfl_xor_sync(0xffffffffu, rd, s);
                    double ouj = __shfl_xor_sync(0xffffffffu, ruj, s);
                    rm = fmin(rm, __shfl_xor_sync(0xffffffffu, rm, s));
                    bool take = oj >= 0 && (rj < 0 || ot < rt || (ot == rt && (((oj & 255) < (rj & 255)) || ((oj & 255) == (rj & 255) && oj < rj))));
                    if (take) { rt = ot; rj = oj; rd = od; ruj = ouj; }
                }
                if (lane == 0) { shc[w].t = rt; shc[w].j = rj; shc[w].d = rd; shc[w].uj = ruj; shc[w].ui = rm; }
                __syncthreads();
                if (w == 0) {
                    rt = shc[lane].t; rj = shc[lane].j; rd = shc[lane].d; ruj = shc[lane].uj; rm = shc[lane].ui;
#pragma unroll
                    for (int s = 16; s >= 1; s >>= 1) {
                        double ot = __shfl_xor_sync(0xffffffffu, rt, s);
                        int oj = __shfl_xor_sync(0xffffffffu, rj, s);
                        double od = __shfl_xor_sync(0xffffffffu, rd, s);
                        double ouj = __shfl_xor_sync(0xffffffffu, ruj, s);
                        rm = fmin(rm, __shfl_xor_sync(0xffffffffu, rm, s));
                        bool take = oj >= 0 && (rj < 0 || ot < rt || (ot == rt && (((oj & 255) < (rj & 255)) || ((oj & 255) == (rj & 255) && oj < rj))));
                        if (take) { rt = ot; rj = oj; rd = od; ruj = ouj; }
                    }
                    if (lane == 0) {
                        if (rm < 1e299) atomicMin(&K[r], enc_f64(rm + C));
                        if (rj >= 0) {
                            if (p_before(rt, r, rj, bt, bi, bj, n)) { bt = rt; bi = r; bj = rj; bd = rd; bui = ur; buj = ruj; }
                        }
                    }
                }
                __syncthreads();
                if (c0 == 0 && tid == 0) my_rows++;
            }
            if (tid == 0) {
                PCand c; c.t = bt; c.i = bi; c.j = bj; c.d = bd; c.ui = bui; c.uj = buj;
                cta_best[cta] = c;
            }
        }
        grid_barrier(bar_flags, G, bar_gen);
        PHASE_MARK(6);

        // ---------------------------------------------------- Phase D: pick (identical in every CTA)
        {
            double t = 1e300; int ci = 0, cj = 0; double cd = 0, cui = 0, cuj = 0;
            if (tid < G) {
                const PCand* cb = cta_best + tid;
                t = __ldcg(&cb->t); ci = __ldcg(&cb->i); cj = __ldcg(&cb->j);
                cd = __ldcg(&cb->d); cui = __ldcg(&cb->ui); cuj = __ldcg(&cb->uj);
            }
            const double mt = t; const int mi = ci, mj = cj;   // this CTA-winner record goes into the candidate pool
            // reduce over the first ceil(G/32) warps
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) {
                double ot = __shfl_xor_sync(0xffffffffu, t, s);
                int oi = __shfl_xor_sync(0xffffffffu, ci, s), oj = __shfl_xor_sync(0xffffffffu, cj, s);
                double od = __shfl_xor_sync(0xffffffffu, cd, s), oui = __shfl_xor_sync(0xffffffffu, cui, s),
                       ouj = __shfl_xor_sync(0xffffffffu, cuj, s);
                if (p_before(ot, oi, oj, t, ci, cj, n)) { t = ot; ci = oi; cj = oj; cd = od; cui = oui; cuj = ouj; }
            }
            if (lane == 0) { shc[w].t = t; shc[w].i = ci; shc[w].j = cj; shc[w].d = cd; shc[w].ui = cui; shc[w].uj = cuj; }
            __syncthreads();
            if (w == 0) {
                t = shc[lane].t; ci = shc[lane].i; cj = shc[lane].j; cd = shc[lane].d; cui = shc[lane].ui; cuj = shc[lane].uj;
#pragma unroll
                for (int s = 16; s >= 1; s >>= 1) {
                    double ot = __shfl_xor_sync(0xffffffffu, t, s);
                    int oi = __shfl_xor_sync(0xffffffffu, ci, s), oj = __shfl_xor_sync(0xffffffffu, cj, s);
                    double od = __shfl_xor_sync(0xffffffffu, cd, s), oui = __shfl_xor_sync(0xffffffffu, cui, s),
                           ouj = __shfl_xor_sync(0xffffffffu, cuj, s);
                    if (p_before(ot, oi, oj, t, ci, cj, n)) { t = ot; ci = oi; cj = oj; cd = od; cui = oui; cuj = ouj; }
                }
                if (lane == 0) { shc[0].t = t; shc[0].i = ci; shc[0].j = cj; shc[0].d = cd; shc[0].ui = cui; shc[0].uj = cuj; }
            }
            __syncthreads();
            int wi = shc[0].i, wj = shc[0].j;
            double wd = shc[0].d, wui = shc[0].ui, wuj = shc[0].uj;
            __syncthreads();
            double uxo, uyo;
            if (wi < wj) { x = wi; y = wj; uxo = wui; uyo = wuj; } else { x = wj; y = wi; uxo = wuj; uyo = wui; }
            dxy = wd;
            // candidate pool: drop pairs touching x or y, move `last` to y, append this scan's per-CTA winners
            {
                const int last_ = n - 1;
                for (int p = tid; p < POOL; p += PT) {
                    int pi = pool_i[p], pj = pool_j[p];
                    if (pi >= 0) {
                        if (pi == x || pi == y || pj == x || pj == y) pool_i[p] = -1;
                        else {
                            if (pi == last_) pool_i[p] = y;
                            if (pj == last_) pool_j[p] = y;
                        }
                    }
                }
                __syncthreads();
                if (tid < G && mt < 1e299 && mi != x && mi != y && mj != x && mj != y) {
                    const int slot = (s_pool_head + tid) % POOL;
                    pool_i[slot] = mi == last_ ? y : mi;
                    pool_j[slot] = mj == last_ ? y : mj;
                }
                __syncthreads();
                if (tid == 0) s_pool_head = (s_pool_head + G) % POOL;
            }
            if (cta == 0 && tid == 0) {
                // host step of the reference, src/neighborJoining.cu:219-237; realID bookkeeping
                // (:233-237) is replayed on the host from this log after the kernel
                double blX = (dxy + uxo - uyo) * 0.5;
                double blY = dxy - blX;
                if (blX < 0) { blY += blX; blX = 0; }
                if (blY < 0) { blX += blY; blY = 0; }
                log_xy[iter] = make_int2(x, y);
                log_bl[iter] = make_double2(blX, blY);
                ps->sel_count = 0;   // next use is after two more grid barriers
            }
            iter++;
            next_id++;
        }
        first = false;
    }
    if (tid == 0 && my_rows) atomicAdd(&ps->rows_scanned, my_rows);
    if (cta == 0 && tid == 0) ps->iters = (unsigned long long)iter;
}

int nj_pruned_loop(dipb_matrix* m, double* U, double* u, double* partial, NJState* st, int* realID, int32_t* c0,
                   int32_t* c1, double* l0, double* l1) {
    (void)partial; (void)st;
    dipb_ctx* c = m->ctx;
    const int n = m->n;
    int G = c->num_sms;
    int max_blocks = 0;
    DIPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&max_blocks, nj_pruned_kernel, PT, 0));
    if (max_blocks < 1) { set_error("nj_pruned: kernel does not fit an SM"); return DIPB_E_CUDA; }
    unsigned long long* K = nullptr;
    double *psum = nullptr, *pmax = nullptr, *dmax_d = nullptr;
    int* sel = nullptr;
    PCand* cb = nullptr;
    PShared* ps = nullptr;
    const int nblk = (n + PT - 1) / PT + 1;
    DIPB_CUDA(pool_alloc(c, (void**)&K, sizeof(unsigned long long) * n));
    DIPB_CUDA(pool_alloc(c, (void**)&psum, sizeof(double) * nblk));
    DIPB_CUDA(pool_alloc(c, (void**)&pmax, sizeof(double) * nblk));
    DIPB_CUDA(pool_alloc(c, (void**)&sel, sizeof(int) * n));
    DIPB_CUDA(pool_alloc(c, (void**)&cb, sizeof(PCand) * G));
    DIPB_CUDA(pool_alloc(c, (void**)&ps, sizeof(PShared)));
    DIPB_CUDA(pool_alloc(c, (void**)&dmax_d, sizeof(double)));
    DIPB_CUDA(cudaMemsetAsync(ps, 0, sizeof(PShared), c->stream));
    DIPB_CUDA(cudaMemsetAsync(K, 0, sizeof(unsigned long long) * n, c->stream));       // 0 = "rescan me"
    // scale for the safety margin: the initial row sums bound every later |d| and |u|
    double dmax = 0.0;
    {
        // max_i U[i] / (n-2) * 2 is a cheap, safe scale (u values stay within the initial range of row means)
        std::vector<double> hu(n);
        DIPB_CUDA(cudaMemcpyAsync(hu.data(), u, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
        DIPB_CUDA(cudaStreamSynchronize(c->stream));
        for (int i = 0; i < n; i++) { double a = hu[i] < 0 ? -hu[i] : hu[i]; if (a == a && a > dmax && a < 1e300) dmax = a; }
        dmax *= 2.0;
    }
    size_t ld = (size_t)n;
    double* Dp = m->d;
    int n_total = n;
    unsigned int* flags = nullptr;
    int2* log_xy = nullptr;
    double2* log_bl = nullptr;
    DIPB_CUDA(pool_alloc(c, (void**)&flags, sizeof(unsigned int) * 1024));
    DIPB_CUDA(cudaMemsetAsync(flags, 0, sizeof(unsigned int) * 1024, c->stream));
    DIPB_CUDA(pool_alloc(c, (void**)&log_xy, sizeof(int2) * n));
    DIPB_CUDA(pool_alloc(c, (void**)&log_bl, sizeof(double2) * n));
    void* args[] = {&Dp, &ld, &U, &u, &K, &psum, &pmax, &sel, &cb, &ps, &flags, &log_xy, &log_bl, &n_total, &dmax};
    cudaError_t e = cudaLaunchCooperativeKernel((void*)nj_pruned_kernel, dim3(G), dim3(PT), args, 0, c->stream);
    if (e != cudaSuccess) { set_error("nj_pruned: cooperative launch failed: %s", cudaGetErrorString(e)); return DIPB_E_CUDA; }
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
    pool_free(c, flags); pool_free(c, log_xy); pool_free(c, log_bl);
    PShared hs;
    DIPB_CUDA(cudaMemcpy(&hs, ps, sizeof(hs), cudaMemcpyDeviceToHost));
    c->nj_rows_scanned = hs.rows_scanned;
    c->nj_iterations = hs.iters;
    c->nj_bytes_scanned = 0;
    if (getenv("DIPB_NJ_PROFILE")) {
        const char* nm[8] = {"A update", "barrier1", "B fold+ub+select", "-", "-", "barrier3", "C2 scan+barrier4", "D pick+pool"};
        double tot = 0;
        for (int k = 0; k < 8; k++) tot += (double)hs.cyc[k];
        fprintf(stderr, "[nj_pruned] n=%d iters=%llu rows_scanned=%llu (%.1f/iter)\n", n, hs.iters, hs.rows_scanned,
                hs.iters ? (double)hs.rows_scanned / hs.iters : 0.0);
        for (int k = 0; k < 8; k++)
            fprintf(stderr, "[nj_pruned]   %-18s %10.0f cyc/iter  %5.1f%%\n", nm[k], hs.iters ? hs.cyc[k] / (double)hs.iters : 0.0,
                    tot > 0 ? 100.0 * hs.cyc[k] / tot : 0.0);
    }
    pool_free(c, K); pool_free(c, psum); pool_free(c, pmax); pool_free(c, sel); pool_free(c, cb); pool_free(c, ps); pool_free(c, dmax_d);
    return 0;
}

}  // namespace dipb
