// Per-pair building blocks of the aligned distance: the boolean functors whose popcounts
// give the reference's pair statistics, and the distance formulas.  Shared by the tile
// kernel (msa_dist.cu) and the divide-and-conquer cluster kernel (dc.cu).
#pragma once
#include "common.cuh"
#include "msa.cuh"

namespace dipb {

// ---------------------------------------------------------------------------
// per-word boolean functors: two 32-bit masks whose popcounts are accumulated
// ---------------------------------------------------------------------------
template <int FN>
__device__ __forceinline__ void pair_masks(uint32_t a0, uint32_t a1, uint32_t av, uint32_t b0, uint32_t b1,
                                           uint32_t bv, uint32_t& m1, uint32_t& m2) {
    uint32_t vv = av & bv;
    uint32_t x0 = a0 ^ b0, x1 = a1 ^ b1;
    if (FN == 0) {  // match, both-valid                       src/MSA.cu:96-97
        m1 = vv & ~(x0 | x1);
        m2 = vv;
    } else if (FN == 1) {  // transitions, transversions         DC/msa.cu:161-162
        m1 = vv & ~x0 & x1;
        m2 = vv & x0;
    } else if (FN == 2) {  // GC at mismatching sites: row seq, column seq   DC/msa.cu:195-196
        uint32_t mm = vv & (x0 | x1);
        m1 = mm & (a0 ^ a1);
        m2 = mm & (b0 ^ b1);
    } else if (FN >= 3 && FN <= 6) {  // occurrences of code FN-3 over both-valid sites  DC/msa.cu:115
        const int c = FN - 3;
        uint32_t ea = ((c & 1) ? a0 : ~a0) & ((c & 2) ? a1 : ~a1);
        uint32_t eb = ((c & 1) ? b0 : ~b0) & ((c & 2) ? b1 : ~b1);
        m1 = vv & ea;
        m2 = vv & eb;
    } else if (FN == 7) {  // unordered pairs {A,G}, {A,T}        DC/msa.cu:121-122
        m1 = vv & x1 & ~a0 & ~b0;
        m2 = vv & x0 & x1 & ~(a0 ^ a1);
    } else {  // FN == 8: {C,G}, {C,T}                            DC/msa.cu:123-124
        m1 = vv & x0 & x1 & (a0 ^ a1);
        m2 = vv & x1 & a0 & b0;
    }
}

__device__ __forceinline__ double dist_p_jc(int match, int useful, int dist_type) {
    // src/MSA.cu:233-235, same expression order
    double uncor = 1 - double(match) / useful;
    if (dist_type == DIPB_DIST_UNCORRECTED) return uncor;
    return -0.75 * log(1.0 - uncor / 0.75);
}

__device__ __forceinline__ double dist_from_counts(int type, int match, int both, int nvi, int nvj, int ts, int tv, int gcr, int gcc,
                                   const int* frac, const int* pr) {
    if (type == DIPB_DIST_UNCORRECTED || type == DIPB_DIST_JC) return dist_p_jc(match, nvi + nvj - both, type);
    int tot = both;
    if (type == DIPB_DIST_TAJIMANEI) {  // DC/msa.cu:239-250
        double fr[4];
        for (int i = 0; i < 4; i++) fr[i] = double(frac[i]) / tot / 2.0;
        double h = 0;
        h += 0.5 * pr[0] * fr[0] * fr[2];
        h += 0.5 * pr[1] * fr[0] * fr[3];
        h += 0.5 * pr[2] * fr[1] * fr[2];
        h += 0.5 * pr[3] * fr[1] * fr[3];
        double D = double(tot - match) / tot;
        double b = 0.5 * (1.0 - fr[0] * fr[0] - fr[2] * fr[2] + D * D / h);
        return -b * log(1.0 - D / b);
    }
    if (type == DIPB_DIST_K2P || type == DIPB_DIST_JINNEI) {  // DC/msa.cu:252-257
        double pp = double(ts) / tot, qq = double(tv) / tot;
        if (type == DIPB_DIST_K2P) return -0.5 * log((1 - 2 * pp - qq) * sqrt(1 - 2 * qq));
        return 0.5 * (1.0 / (1 - 2 * pp - qq) + 0.5 / (1 - qq * 2) - 1.5);
    }
    if (type == DIPB_DIST_TAMURA) {  // DC/msa.cu:259-263
        double pp = double(ts) / tot, qq = double(tv) / tot,
               c = double(gcr) / tot + double(gcc) / tot - 2 * double(gcr) * double(gcc) / tot / tot;
        return -c * log(1 - pp / c - qq) - 0.5 * (1 - c) * log(1 - 2 * qq);
    }
    return 0.0;
}


// One warp computes d(i, j) straight from the blocked planes (used where pairs are
// scattered: in-cluster placement).  All lanes return the distance.
__device__ __forceinline__ double msa_pair_warp(const uint32_t* __restrict__ planes, const int* __restrict__ nv, int nkc,
                                                int type, int i, int j) {
    const int lane = threadIdx.x & 31;
    const size_t bi = (size_t)(i / MSA_TS) * nkc, bj = (size_t)(j / MSA_TS) * nkc;
    const int li = i % MSA_TS, lj = j % MSA_TS;
    int cnt[18];
#pragma unroll
    for (int q = 0; q < 18; q++) cnt[q] = 0;
    const int words = nkc * MSA_KC;
    for (int w = lane; w < words; w += 32) {
        const int kc = w / MSA_KC, kk = w % MSA_KC;
        const uint32_t* A = planes + (bi + kc) * MSA_SLAB_WORDS + kk * MSA_TS + li;
        const uint32_t* B = planes + (bj + kc) * MSA_SLAB_WORDS + kk * MSA_TS + lj;
        const uint32_t a0 = A[0], a1 = A[MSA_KC * MSA_TS], av = A[2 * MSA_KC * MSA_TS];
        const uint32_t b0 = B[0], b1 = B[MSA_KC * MSA_TS], bv = B[2 * MSA_KC * MSA_TS];
        uint32_t m1, m2;
        pair_masks<0>(a0, a1, av, b0, b1, bv, m1, m2); cnt[0] += __popc(m1); cnt[1] += __popc(m2);
        if (type >= DIPB_DIST_TAJIMANEI) {
            pair_masks<1>(a0, a1, av, b0, b1, bv, m1, m2); cnt[2] += __popc(m1); cnt[3] += __popc(m2);
            pair_masks<2>(a0, a1, av, b0, b1, bv, m1, m2); cnt[4] += __popc(m1); cnt[5] += __popc(m2);
            pair_masks<3>(a0, a1, av, b0, b1, bv, m1, m2); cnt[6] += __popc(m1); cnt[7] += __popc(m2);
            pair_masks<4>(a0, a1, av, b0, b1, bv, m1, m2); cnt[8] += __popc(m1); cnt[9] += __popc(m2);
            pair_masks<5>(a0, a1, av, b0, b1, bv, m1, m2); cnt[10] += __popc(m1); cnt[11] += __popc(m2);
            pair_masks<6>(a0, a1, av, b0, b1, bv, m1, m2); cnt[12] += __popc(m1); cnt[13] += __popc(m2);
            pair_masks<7>(a0, a1, av, b0, b1, bv, m1, m2); cnt[14] += __popc(m1); cnt[15] += __popc(m2);
            pair_masks<8>(a0, a1, av, b0, b1, bv, m1, m2); cnt[16] += __popc(m1); cnt[17] += __popc(m2);
        }
    }
#pragma unroll
    for (int q = 0; q < 18; q++)
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) cnt[q] += __shfl_xor_sync(0xffffffffu, cnt[q], s);
    int frac[4] = {cnt[6] + cnt[7], cnt[8] + cnt[9], cnt[10] + cnt[11], cnt[12] + cnt[13]};
    int pr[4] = {cnt[14], cnt[15], cnt[16], cnt[17]};
    // row = i (the tip being placed), column = j
    return dist_from_counts(type, cnt[0], cnt[1], nv[i], nv[j], cnt[2], cnt[3], cnt[4], cnt[5], frac, pr);
}

}  // namespace dipb
