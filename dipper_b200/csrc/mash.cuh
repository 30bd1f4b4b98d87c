// Mash device object (see mash.cu).
#pragma once
#include "common.cuh"

struct dipb_mash {
    dipb_ctx* ctx = nullptr;
    int n = 0, k = 15, s = 1000;
    uint64_t* seqs = nullptr;      // flat 2-bit words (+1 pad word)
    uint64_t* word_off = nullptr;  // [n]
    uint64_t* lens = nullptr;      // [n] bases
    uint64_t* sketches = nullptr;  // [n][s]
    uint32_t* ranks = nullptr;     // [n][s]: dense rank of every hash among all n*s hashes (same order, same equalities), built lazily
    bool sketched = false;
};


namespace dipb {
// thread-sequential merge of two sorted sketches (src/mash.cu:437-454): A = column, B = row
__device__ __forceinline__ double mash_pair_thread(const uint64_t* __restrict__ A, const uint64_t* __restrict__ B, int s, int k) {
    int a = 0, b = 0, uni = 0, inter = 0;
    uint64_t av = A[0], bv = B[0];
    while (uni < s) {
        const bool bvalid = b < s;
        const bool takeB = bvalid && bv <= av;
        const bool eq = takeB && bv == av;
        if (takeB) { b++; bv = B[b < s ? b : s - 1]; }
        else { a++; av = A[a < s ? a : s - 1]; }
        uni += eq ? 0 : 1;
        inter += eq ? 1 : 0;
    }
    double jac = fmax(double(inter), 1.0) / uni;
    return fmin(1.0, fabs(log(2.0 * jac / (1.0 + jac)) / k));
}
}  // namespace dipb
