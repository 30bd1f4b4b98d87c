// Divide-and-conquer driver (-m 3) and its kernels (sm_100a).
//
// Replaces KPlacementDeviceArraysDC::{findBackboneTreeDC, findClustersDC, findClusterTreeDC}
// (reference DC/placement_close_k.cu:731-1535) and the D&C distance variants
// (DC/msa.cu:219-504, DC/mash.cu:453-640).  Same three stages and the same tree arrays /
// numbering as the reference (slots and internal nodes are numbered by prefix sums over
// clusters, equal to the reference's running counters), but:
//  * stage 1 (backbone) is the persistent placement kernel of placement.cu;
//  * stage 2 (cluster assignment) computes query x backbone distance BLOCKS with the tiled
//    kernels and scores one query per CTA -- the reference launches 3 kernels + a D2H per query;
//  * stage 3 places every cluster in its own CTA, all clusters concurrently (they touch
//    disjoint slots; the reference's own CPU twin runs them in a tbb::parallel_for,
//    DC/placement_close_k.cpp:752-760); everything stays device-resident, no per-cluster
//    host gather + H2D.
#include <algorithm>
#include <vector>
#include "common.cuh"
#include "mash.cuh"
#include "msa_pair.cuh"
#include "placement_dev.cuh"

namespace dipb {

constexpr int DC_THREADS = 128;

// ---- stage 2: one query per CTA, all 4B-4 backbone slots -----------------------------
__global__ void __launch_bounds__(256)
dc_assign_kernel(const int* __restrict__ e, const int* __restrict__ belong, const double* __restrict__ len,
                 const int* __restrict__ cid, const double* __restrict__ cdis, const int* __restrict__ rev, int nslots,
                 const double* __restrict__ rows, size_t ld, int q0, int nq, int* __restrict__ cluster) {
    __shared__ PlCand sb[8];
    for (int qi = blockIdx.x; qi < nq; qi += gridDim.x) {
        const double* dis = rows + (size_t)qi * ld;
        double badd = 2.0, bfrac = 0.0;
        int bslot = 0;
        for (int q = threadIdx.x; q < nslots; q += blockDim.x) {
            if (belong[q] > e[q]) {
                double f, a;
                score_slot(dis, cid, cdis, len, rev, q, f, a);
                if (a < badd || (a == badd && q < bslot)) { badd = a; bfrac = f; bslot = q; }
            }
        }
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) {
            double oa = __shfl_xor_sync(0xffffffffu, badd, s);
            int os = __shfl_xor_sync(0xffffffffu, bslot, s);
            if (oa < badd || (oa == badd && os < bslot)) { badd = oa; bslot = os; }
        }
        if (lane == 0) { sb[w].add = badd; sb[w].slot = bslot; }
        __syncthreads();
        if (threadIdx.x == 0) {
            PlCand b = sb[0];
            for (int k = 1; k < (int)(blockDim.x / 32); k++)
                if (sb[k].add < b.add || (sb[k].add == b.add && sb[k].slot < b.slot)) b = sb[k];
            cluster[q0 + qi] = (b.add < 2.0) ? b.slot : 0;   // the (0,0,2) tuple at position 0 wins otherwise
        }
        __syncthreads();
        (void)bfrac;
    }
}

// ---- stage 3: one cluster per CTA -------------------------------------------------------
struct DcSource {
    // exactly one of the three
    const uint32_t* planes; const int* nv; int nkc; int dist_type;   // aligned
    const uint64_t* sketches; int s, k;                              // mash
    const double* matrix; size_t mld;                                // matrix
};

struct DcArgs {
    int *head, *e, *nxt, *belong, *cid, *rev;
    double *len, *cdis;
    int n, B;
    int num_clusters;
    const int* cl_slot;     // [num_clusters] backbone slot of each cluster, ascending
    const int* cl_off;      // [num_clusters+1] prefix of cluster sizes
    const int* cl_tips;     // tips sorted by (cluster, tip)
    int* leaf_mask;         // 10 * num_clusters + total tips
    double* distm;          // same shape
    int* edge_mask;         // 2 * num_clusters + 4 * total tips
    int* q_node; int* q_from; double* q_dis;   // 4 + 4 * total tips + 2 per cluster ... sized like edge_mask + tips
    int* pos_of;            // [n] position of a cluster tip in its leaf mask
    int* owner;             // [8n] cluster owning a slot, -1 otherwise
    unsigned int* next_cluster;
};

__device__ __forceinline__ double dc_lookup(const DcArgs& a, const int* lm, const double* dm, int id) {
    if (id < a.B) {
#pragma unroll
        for (int k = 0; k < 10; k++)
            if (lm[k] == id) return dm[k];
        return 0.0;   // unreachable: closest lists of masked slots only hold mask leaves
    }
    return dm[a.pos_of[id]];
}

__global__ void __launch_bounds__(DC_THREADS)
dc_cluster_kernel(DcArgs a, DcSource src) {
    __shared__ int s_cluster;
    __shared__ PlCand s_best[DC_THREADS / 32];
    __shared__ int s_pnc[DC_THREADS / 32];
    __shared__ unsigned int s_lo, s_hi, s_tail;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    while (true) {
        if (tid == 0) s_cluster = (int)atomicAdd(a.next_cluster, 1u);
        __syncthreads();
        const int c = s_cluster;
        __syncthreads();
        if (c >= a.num_clusters) break;
        const int cj = a.cl_slot[c];
        const int t0 = a.cl_off[c], t1 = a.cl_off[c + 1];
        int* lm = a.leaf_mask + (size_t)10 * c + t0;
        double* dm = a.distm + (size_t)10 * c + t0;
        int* em = a.edge_mask + (size_t)2 * c + 4 * (size_t)t0;
        int* qn = a.q_node + (size_t)4 * c + 4 * (size_t)t0;
        int* qf = a.q_from + (size_t)4 * c + 4 * (size_t)t0;
        double* qd = a.q_dis + (size_t)4 * c + 4 * (size_t)t0;
        // initializeClusterDC (DC/placement_close_k.cu:604-628)
        const int oth = a.rev[cj];
        if (tid < 5) lm[tid] = a.cid[cj * KC5 + tid];
        else if (tid < 10) lm[tid] = a.cid[oth * KC5 + (tid - 5)];
        if (tid == 0) { em[0] = cj; em[1] = oth; a.owner[cj] = c; a.owner[oth] = c; }
        __syncthreads();
        int leafCount = 10, edgeCount = 2;
        for (int t = t0; t < t1; t++) {
            const int leaf = a.cl_tips[t];
            const int idx = 4 * a.B - 4 + 4 * t;        // running slot counter of the reference
            const int placeCount = a.B + t;             // insertLeafCount
            // ---- distances tip -> mask leaves
            if (src.planes) {
                for (int p = w; p < leafCount; p += DC_THREADS / 32) {
                    const int id = lm[p];
                    if (id != -1) {
                        double d = msa_pair_warp(src.planes, src.nv, src.nkc, src.dist_type, leaf, id);
                        if (lane == 0) dm[p] = d;
                    }
                }
            } else {
                for (int p = tid; p < leafCount; p += DC_THREADS) {
                    const int id = lm[p];
                    if (id != -1) {
                        if (src.sketches) dm[p] = mash_pair_thread(src.sketches + (size_t)id * src.s, src.sketches + (size_t)leaf * src.s, src.s, src.k);
                        else dm[p] = src.matrix[(size_t)leaf * src.mld + id];
                    }
                }
            }
            __syncthreads();
            // ---- score masked edges; first minimum by mask POSITION (:180-233)
            double badd = 1e300, bfrac = 0.0;
            int bpos = 0x7fffffff, bslot = 0, pnc = 0x7fffffff;
            for (int p = tid; p < edgeCount; p += DC_THREADS) {
                const int q = em[p];
                if (a.belong[q] < a.e[q]) { if (p < pnc) pnc = p; continue; }
                const int r = a.rev[q];
                double d1 = 0, d2 = 0;
                for (int k = 0; k < KC5; k++) {
                    int id = a.cid[q * KC5 + k];
                    if (id != -1) { double v = dc_lookup(a, lm, dm, id) - a.cdis[q * KC5 + k]; if (v > d1) d1 = v; }
                }
                for (int k = 0; k < KC5; k++) {
                    int id = a.cid[r * KC5 + k];
                    if (id != -1) { double v = dc_lookup(a, lm, dm, id) - a.cdis[r * KC5 + k]; if (v > d2) d2 = v; }
                }
                const double L = a.len[q];
                double ad = (d1 + d2 - L) / 2;
                if (ad < 0) ad = 0;
                d1 -= ad; d2 -= ad;
                if (d1 < 0) d1 = 0;
                if (d2 < 0) d2 = 0;
                if (d1 > L) { ad += d1 - L; d1 = L; }
                if (d2 > L) { ad += d2 - L; d2 = L; }
                const double rest = L - d1 - d2;
                d1 += rest / 2;
                if (ad < badd || (ad == badd && p < bpos)) { badd = ad; bfrac = d1; bpos = p; bslot = q; }
            }
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) {
                double oa = __shfl_xor_sync(0xffffffffu, badd, s), of = __shfl_xor_sync(0xffffffffu, bfrac, s);
                int op = __shfl_xor_sync(0xffffffffu, bpos, s), os = __shfl_xor_sync(0xffffffffu, bslot, s);
                int on = __shfl_xor_sync(0xffffffffu, pnc, s);
                if (oa < badd || (oa == badd && op < bpos)) { badd = oa; bfrac = of; bpos = op; bslot = os; }
                pnc = min(pnc, on);
            }
            if (lane == 0) { s_best[w].add = badd; s_best[w].frac = bfrac; s_best[w].slot = bslot; s_best[w].pad = bpos; s_pnc[w] = pnc; }
            __syncthreads();
            if (tid == 0) {
                PlCand b = s_best[0];
                int pn = s_pnc[0];
                for (int k = 1; k < DC_THREADS / 32; k++) {
                    if (s_best[k].add < b.add || (s_best[k].add == b.add && s_best[k].pad < b.pad)) b = s_best[k];
                    pn = min(pn, s_pnc[k]);
                }
                // default tuple (0,0,2) sits at every non-candidate position
                if (!(b.add < 2.0 || (b.add == 2.0 && b.pad < pn))) { b.slot = 0; b.frac = 0.0; b.add = 2.0; }
                // updateTreeStructureInClusterDC (:442-525): middle = placeCount + n - 1
                split_edge(a.head, a.nxt, a.e, a.len, a.cdis, a.cid, a.belong, a.rev, b.slot, b.frac, b.add, leaf, idx,
                           a.n + placeCount - leaf);
                // updateClusterInfoDC (:553-572)
                lm[leafCount] = leaf;
                a.pos_of[leaf] = leafCount;
                for (int k = 1; k <= 4; k++) { em[edgeCount + k - 1] = idx + 4 - k; a.owner[idx + 4 - k] = c; }
                // BFS seed (updateClosestNodesInClusterDC :312-356)
                qn[0] = leaf; qf[0] = -1; qd[0] = 0.0;
                s_lo = 0; s_hi = 1; s_tail = 1;
            }
            leafCount++; edgeCount += 4;
            __syncthreads();
            const int ed1 = a.e[cj], ed2 = a.belong[cj];
            while (true) {
                const unsigned int l = s_lo, h = s_hi;
                if (l >= h) break;
                for (unsigned int u = l + tid; u < h; u += DC_THREADS) {
                    const int node = qn[u], fb = qf[u];
                    const double d = qd[u];
                    if (node == ed1 || node == ed2) continue;
                    for (int s = a.head[node]; s != -1; s = a.nxt[s]) {
                        if (a.owner[s] != c) continue;
                        if (a.e[s] == fb) continue;
                        if (list_insert(a.cdis, a.cid, s, d, leaf)) {
                            unsigned int pos = atomicAdd(&s_tail, 1u);
                            qn[pos] = a.e[s]; qf[pos] = node; qd[pos] = d + a.len[s];
                        }
                    }
                }
                __syncthreads();
                if (tid == 0) { s_lo = h; s_hi = s_tail; }
                __syncthreads();
            }
            __syncthreads();
        }
    }
}

__global__ void fill_int_kernel(int* p, long long n, int v) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace dipb

using namespace dipb;

static std::vector<int32_t> g_last_clusters;   // test hook storage (dipb_dc_cluster_ids)

extern "C" {

int dipb_dc(dipb_ctx* c, const dipb_dist_source* src, int n, int backbone, dipb_tree** out) {
    if (!c || !src || !out || n < 4) { set_error("dipb_dc: bad argument"); return DIPB_E_ARG; }
    const int B = backbone;
    if (B < 2 || B >= n) { set_error("dipb_dc: backbone size %d must be in [2, n)", B); return DIPB_E_ARG; }
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
    // ---- stage 1: backbone tree over tips [0, B), internal ids offset by n
    if (!rc) rc = place_from_scratch(c, src, n, B, t, &sc);
    place_scratch_free(&sc);
    if (rc) { dipb_tree_free(t); return rc; }

    // ---- stage 2: cluster of every tip >= B
    const int nslots = 4 * B - 4;
    int* d_cluster = nullptr;
    DIPB_CUDA(cudaMalloc(&d_cluster, sizeof(int) * n));
    DIPB_CUDA(cudaMemsetAsync(d_cluster, 0xff, sizeof(int) * n, c->stream));
    {
        int qb = 1024;
        const size_t ld = (size_t)((B + 127) / 128 * 128);
        while ((size_t)qb * ld * sizeof(double) > (1ull << 30) && qb > 128) qb /= 2;
        double* buf = nullptr;
        if (!src->matrix) DIPB_CUDA(cudaMalloc(&buf, (size_t)qb * ld * sizeof(double)));
        for (int q0 = B; q0 < n && !rc; q0 += qb) {
            int q1 = q0 + qb < n ? q0 + qb : n;
            const double* rows; size_t ldr;
            if (src->matrix) { rows = src->matrix->d + (size_t)q0 * src->matrix->n; ldr = (size_t)src->matrix->n; }
            else {
                rows = buf; ldr = ld;
                rc = src->msa ? msa_block(src->msa, src->dist_type, q0, q1, B, buf, ld) : dipb_mash_dist_block(src->mash, q0, q1, B, buf, ld);
                if (rc) break;
            }
            int grid = q1 - q0 < c->num_sms * 8 ? q1 - q0 : c->num_sms * 8;
            dc_assign_kernel<<<grid, 256, 0, c->stream>>>(t->e, t->belong, t->len, t->cid, t->cdis, t->rev, nslots, rows, ldr, q0, q1 - q0, d_cluster);
            c->launches++;
        }
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (buf) cudaFree(buf);
        if (!rc && e != cudaSuccess) { set_error("dipb_dc stage 2: %s", cudaGetErrorString(e)); rc = DIPB_E_CUDA; }
        if (rc) { cudaFree(d_cluster); dipb_tree_free(t); return rc; }
    }
    std::vector<int32_t> cl(n);
    DIPB_CUDA(cudaMemcpy(cl.data(), d_cluster, sizeof(int) * n, cudaMemcpyDeviceToHost));
    cudaFree(d_cluster);
    g_last_clusters = cl;

    // ---- cluster lists: ascending slot, tips ascending (contains[], :1283-1285)
    const int ntips = n - B;
    std::vector<int> order(ntips);
    for (int i = 0; i < ntips; i++) order[i] = B + i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cl[x] < cl[y]; });
    std::vector<int> cl_slot, cl_off;
    for (int i = 0; i < ntips; i++) {
        if (i == 0 || cl[order[i]] != cl[order[i - 1]]) { cl_slot.push_back(cl[order[i]]); cl_off.push_back(i); }
    }
    cl_off.push_back(ntips);
    const int nc = (int)cl_slot.size();

    // ---- stage 3
    DcArgs a{};
    a.head = t->head; a.e = t->e; a.nxt = t->nxt; a.belong = t->belong; a.cid = t->cid; a.rev = t->rev; a.len = t->len; a.cdis = t->cdis;
    a.n = n; a.B = B; a.num_clusters = nc;
    int *d_slot = nullptr, *d_off = nullptr, *d_tips = nullptr;
    const size_t lm_sz = (size_t)10 * nc + ntips, em_sz = (size_t)4 * nc + 4 * (size_t)ntips + 8;
    DIPB_CUDA(cudaMalloc(&d_slot, sizeof(int) * (nc + 1)));
    DIPB_CUDA(cudaMalloc(&d_off, sizeof(int) * (nc + 1)));
    DIPB_CUDA(cudaMalloc(&d_tips, sizeof(int) * (ntips + 1)));
    DIPB_CUDA(cudaMalloc(&a.leaf_mask, sizeof(int) * lm_sz));
    DIPB_CUDA(cudaMalloc(&a.distm, sizeof(double) * lm_sz));
    DIPB_CUDA(cudaMalloc(&a.edge_mask, sizeof(int) * em_sz));
    DIPB_CUDA(cudaMalloc(&a.q_node, sizeof(int) * em_sz));
    DIPB_CUDA(cudaMalloc(&a.q_from, sizeof(int) * em_sz));
    DIPB_CUDA(cudaMalloc(&a.q_dis, sizeof(double) * em_sz));
    DIPB_CUDA(cudaMalloc(&a.pos_of, sizeof(int) * n));
    DIPB_CUDA(cudaMalloc(&a.owner, sizeof(int) * 8 * (size_t)n));
    DIPB_CUDA(cudaMalloc(&a.next_cluster, sizeof(unsigned int)));
    DIPB_CUDA(cudaMemsetAsync(a.next_cluster, 0, sizeof(unsigned int), c->stream));
    DIPB_CUDA(cudaMemcpyAsync(d_slot, cl_slot.data(), sizeof(int) * nc, cudaMemcpyHostToDevice, c->stream));
    DIPB_CUDA(cudaMemcpyAsync(d_off, cl_off.data(), sizeof(int) * (nc + 1), cudaMemcpyHostToDevice, c->stream));
    DIPB_CUDA(cudaMemcpyAsync(d_tips, order.data(), sizeof(int) * ntips, cudaMemcpyHostToDevice, c->stream));
    a.cl_slot = d_slot; a.cl_off = d_off; a.cl_tips = d_tips;
    {
        long long tot = 8LL * n;
        fill_int_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, c->stream>>>(a.owner, tot, -1);
        c->launches++;
    }
    DcSource ds{};
    if (src->msa) { ds.planes = src->msa->planes; ds.nv = src->msa->nv; ds.nkc = src->msa->nkc; ds.dist_type = src->dist_type; }
    else if (src->mash) { ds.sketches = src->mash->sketches; ds.s = src->mash->s; ds.k = src->mash->k; }
    else { ds.matrix = src->matrix->d; ds.mld = (size_t)src->matrix->n; }
    {
        int grid = nc < c->num_sms * 8 ? nc : c->num_sms * 8;
        if (grid < 1) grid = 1;
        dc_cluster_kernel<<<grid, DC_THREADS, 0, c->stream>>>(a, ds);
        c->launches++;
    }
    cudaError_t e = cudaStreamSynchronize(c->stream);
    cudaFree(d_slot); cudaFree(d_off); cudaFree(d_tips); cudaFree(a.leaf_mask); cudaFree(a.distm); cudaFree(a.edge_mask);
    cudaFree(a.q_node); cudaFree(a.q_from); cudaFree(a.q_dis); cudaFree(a.pos_of); cudaFree(a.owner); cudaFree(a.next_cluster);
    if (e != cudaSuccess) { set_error("dipb_dc stage 3: %s", cudaGetErrorString(e)); dipb_tree_free(t); return DIPB_E_CUDA; }
    rc = timer_end(c, DIPB_T_PLACE);
    if (rc) return rc;
    *out = t;
    return 0;
}

int dipb_dc_cluster_ids(dipb_ctx* c, int32_t* h_out, int n) {
    if (!c || !h_out || (int)g_last_clusters.size() != n) { set_error("dipb_dc_cluster_ids: no matching dipb_dc run"); return DIPB_E_STATE; }
    memcpy(h_out, g_last_clusters.data(), sizeof(int32_t) * n);
    return 0;
}

}  // extern "C"
