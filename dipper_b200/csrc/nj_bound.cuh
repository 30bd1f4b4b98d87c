// Helpers shared by the bound-pruned NJ kernels (nj_pruned.cu: whole-grid version, nj_cluster.cu: one
// thread-block cluster): order-preserving float encodings and the reference's tie order
// (src/neighborJoining.cu:117-148 + thrust::min_element over the per-block results, :214).
#pragma once
#include "common.cuh"

namespace dipb {

__device__ __forceinline__ unsigned long long enc_f64(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dec_f64(unsigned long long e) {
    unsigned long long b = (e >> 63) ? (e & 0x7fffffffffffffffull) : ~e;
    return __longlong_as_double((long long)b);
}
__device__ __forceinline__ unsigned int enc_f32(float v) {
    unsigned int b = __float_as_uint(v);
    return (b >> 31) ? ~b : (b | 0x80000000u);
}

__device__ __forceinline__ float dec_f32(unsigned int e) {
    unsigned int b = (e >> 31) ? (e & 0x7fffffffu) : ~e;
    return __uint_as_float(b);
}

__device__ __forceinline__ int p_rowblock_of(int i, int n) {
    const int sz = n / 256, rem = n % 256;
    const int split = (sz + 1) * rem;   // <= n
    if (i < split) return i / (sz + 1);
    return rem + (i - split) / sz;
}
__device__ __forceinline__ unsigned long long p_tie_key(int i, int j, int n) {
    return ((unsigned long long)p_rowblock_of(i, n) << 56) | ((unsigned long long)(j & 255) << 48) |
           ((unsigned long long)j << 24) | (unsigned long long)i;
}
__device__ __forceinline__ bool p_before(double ta, int ia, int ja, double tb, int ib, int jb, int n) {
    if (ta < tb) return true;
    if (ta > tb) return false;
    if (ta >= 10000.0) return false;
    return p_tie_key(ia, ja, n) < p_tie_key(ib, jb, n);
}

}  // namespace dipb
