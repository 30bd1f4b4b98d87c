// K2/K3: MinHash sketches of unaligned sequences and sketch-vs-sketch Mash distance.
//
// Replaces MashDeviceArrays::{allocateDeviceArrays, sketchConstructionOnGpu,
// distConstructionOnGpu} (reference src/mash.cu:14-122,260-471) and the D&C twins
// (DC/mash.cu).  Not a port:
//  * sketching: one CTA per sequence hashes every canonical k-mer (MurmurHash3_x64_128,
//    seed 42, low 64 bits; canonical form chosen by comparing the 2-bit forward and
//    reverse-complement words instead of byte strings) and keeps the bottom-s MULTISET
//    with a threshold-filtered shared-memory buffer + bitonic sort (the reference re-sorts
//    1536 keys per 512 k-mers); sketches stay on the device, row-major [n][s].
//  * distance: 16 x 8 sketch tiles are staged in shared memory with 1-D TMA bulk copies
//    and every thread walks one pair with the reference's merge rule, branch-free per
//    step; p-value free Mash distance in fp64 with the reference's expression.
#include <vector>
#include "common.cuh"

#include "mash.cuh"
#include <cub/cub.cuh>

namespace dipb {

constexpr int SK_THREADS = 1024;
constexpr int SK_BUF = 4096;

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ uint64_t fmix64(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}

// MurmurHash3_x64_128 (low 64 bits) of the k ASCII bytes of a k-mer given as 2-bit codes,
// first base in the lowest bits of `codes` (k <= 32).  src/mash.cu:159-236.
__device__ __forceinline__ uint64_t murmur_kmer(uint64_t codes, int k, uint32_t seed) {
    // expand to ASCII, 8 bases per 64-bit word, byte t = base t (little endian like the char array)
    uint64_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int t = 0; t < 32; t++) {
        if (t < k) {
            uint32_t c = (uint32_t)(codes >> (2 * t)) & 3u;
            // A 0x41, C 0x43, G 0x47, T 0x54
            uint64_t ch = c == 0 ? 0x41 : (c == 1 ? 0x43 : (c == 2 ? 0x47 : 0x54));
            w[t >> 3] |= ch << (8 * (t & 7));
        }
    }
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = seed, h2 = seed;
    const int nblocks = k / 16;
#define DIPB_MURMUR_BLOCK(K1, K2)                                              \
    do {                                                                       \
        uint64_t k1 = (K1), k2 = (K2);                                         \
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;                     \
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;               \
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;                     \
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;               \
    } while (0)
    if (nblocks >= 1) DIPB_MURMUR_BLOCK(w[0], w[1]);
    if (nblocks >= 2) DIPB_MURMUR_BLOCK(w[2], w[3]);
#undef DIPB_MURMUR_BLOCK
    const int rem = k & 15;
    if (rem) {
        // bytes beyond k are zero in w[], so the switch fall-through of the reference is implicit
        uint64_t k1 = nblocks == 0 ? w[0] : w[2], k2 = nblocks == 0 ? w[1] : w[3];
        if (rem > 8) { k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    }
    h1 ^= (uint64_t)k; h2 ^= (uint64_t)k;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2;
    return h1;
}

// canonical k-mer (src/mash.cu:239-258,320-321): the lexicographically smaller of the
// forward string and its reverse complement (A<C<G<T == code order), forward on ties.
__device__ __forceinline__ uint64_t canonical_codes(uint64_t fwd, int k) {
    // rc: complement (3 - c == ~c & 3) and reverse the order of the k 2-bit groups
    uint64_t x = ~fwd;
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    x = __byte_perm((uint32_t)(x >> 32), 0, 0x0123) | ((uint64_t)__byte_perm((uint32_t)x, 0, 0x0123) << 32);
    uint64_t rc = k == 32 ? x : (x >> (64 - 2 * k));
    uint64_t mask = k == 32 ? ~0ULL : ((1ULL << (2 * k)) - 1);
    fwd &= mask;
    // lexicographic compare == integer compare with the FIRST base most significant
    auto msb_first = [&](uint64_t v) {
        uint64_t y = v;
        y = ((y >> 2) & 0x3333333333333333ULL) | ((y & 0x3333333333333333ULL) << 2);
        y = ((y >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((y & 0x0F0F0F0F0F0F0F0FULL) << 4);
        y = __byte_perm((uint32_t)(y >> 32), 0, 0x0123) | ((uint64_t)__byte_perm((uint32_t)y, 0, 0x0123) << 32);
        return k == 32 ? y : (y >> (64 - 2 * k));
    };
    return msb_first(fwd) <= msb_first(rc) ? fwd : rc;
}

__device__ __forceinline__ void bitonic_sort_smem(uint64_t* buf, int n_pow2) {
    for (int size = 2; size <= n_pow2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < n_pow2 / 2; t += blockDim.x) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool up = (lo & size) == 0;
                uint64_t a = buf[lo], b = buf[hi];
                if ((a > b) == up) { buf[lo] = b; buf[hi] = a; }
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(SK_THREADS, 1)
sketch_kernel(const uint64_t* __restrict__ seqs, const uint64_t* __restrict__ word_off, const uint64_t* __restrict__ lens,
              int n, int k, int s, uint64_t* __restrict__ out) {
    __shared__ uint64_t buf[SK_BUF];
    __shared__ int cnt;
    __shared__ uint64_t thr;
    const int tid = threadIdx.x;
    for (int seq = blockIdx.x; seq < n; seq += gridDim.x) {
        const uint64_t* w = seqs + word_off[seq];
        const uint64_t len = lens[seq];
        for (int t = tid; t < SK_BUF; t += SK_THREADS) buf[t] = ~0ULL;
        if (tid == 0) { cnt = 0; thr = ~0ULL; }
        __syncthreads();
        const uint64_t nk = len >= (uint64_t)k ? len - k + 1 : 0;
        for (uint64_t base = 0; base < nk; base += SK_THREADS) {
            // make room: after a sort only the bottom s survive
            if (cnt > SK_BUF - SK_THREADS) {
                bitonic_sort_smem(buf, SK_BUF);
                for (int t = s + tid; t < SK_BUF; t += SK_THREADS) buf[t] = ~0ULL;
                __syncthreads();
                if (tid == 0) { cnt = s; thr = buf[s - 1]; }
                __syncthreads();
            }
            const uint64_t j = base + tid;
            if (j < nk) {
                const uint64_t idx = j >> 5;
                const int sh = 2 * (int)(j & 31);
                uint64_t kmer = w[idx] >> sh;
                if (sh) kmer |= w[idx + 1] << (64 - sh);   // flat buffer carries one pad word (App. B18)
                uint64_t h = murmur_kmer(canonical_codes(kmer, k), k, 42u);
                if (h < thr) {
                    int pos = atomicAdd(&cnt, 1);
                    buf[pos] = h;
                }
            }
            __syncthreads();
        }
        bitonic_sort_smem(buf, SK_BUF);
        for (int t = tid; t < s; t += SK_THREADS) out[(size_t)seq * s + t] = buf[t];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// Mash distance tiles
// ---------------------------------------------------------------------------
constexpr int MD_TA = 16;   // column sketches per tile (the "A" list of src/mash.cu:437)
constexpr int MD_TB = 8;    // row sketches per tile (the "B" list, rowId)
constexpr int MD_THREADS = MD_TA * MD_TB;

struct MashTileParams {
    const uint64_t* sk;
    int n, s, k;
    int tri;           // 1: lower triangle with mirror into an n x n matrix
    int r0, r1, ncols; // rectangle: rows [r0,r1) x cols [0,ncols)
    double* out;
    size_t ld;
    int row_off;
};

__global__ void __launch_bounds__(MD_THREADS, 1) mash_tile_kernel(MashTileParams p, long long num_tiles, int tiles_x) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* sA = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* sB = sA + (size_t)MD_TA * p.s;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + (size_t)MD_TB * p.s);
    const int tid = threadIdx.x;
    const int ia = tid % MD_TA, ib = tid / MD_TA;
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncthreads();
    uint32_t phase = 0;
    const uint32_t sk_bytes = (uint32_t)p.s * 8u;
    for (long long t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int tb, ta;
        if (p.tri) {
            // tile rows of MD_TB, tile cols of MD_TA over the lower triangle (col0 <= row_max)
            tb = (int)(t / tiles_x); ta = (int)(t % tiles_x);
        } else {
            tb = (int)(t / tiles_x); ta = (int)(t % tiles_x);
        }
        const int row0 = p.r0 + tb * MD_TB, col0 = ta * MD_TA;
        const int row_max = min(row0 + MD_TB, p.r1) - 1;
        if (p.tri && col0 > row_max) continue;   // strictly above the diagonal: nothing to do (uniform per CTA)
        if (tid == 0) {
            int na = min(MD_TA, p.n - col0), nb = min(MD_TB, p.r1 - row0);
            mbar_arrive_expect_tx(bar, (uint32_t)(na + nb) * sk_bytes);
            for (int q = 0; q < na; q++) tma_bulk_g2s(sA + (size_t)q * p.s, p.sk + (size_t)(col0 + q) * p.s, sk_bytes, bar);
            for (int q = 0; q < nb; q++) tma_bulk_g2s(sB + (size_t)q * p.s, p.sk + (size_t)(row0 + q) * p.s, sk_bytes, bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        const int i = row0 + ib, j = col0 + ia;
        const int jlim = p.tri ? i : p.ncols;
        if (i < p.r1 && j < jlim && j < p.n) {
            const uint64_t* A = sA + (size_t)ia * p.s;
            const uint64_t* B = sB + (size_t)ib * p.s;
            const int s = p.s;
            int a = 0, b = 0, uni = 0, inter = 0;
            uint64_t av = A[0], bv = B[0];
            // src/mash.cu:439-450, one consumed element per step
            while (uni < s) {
                const bool bvalid = b < s;
                const bool takeB = bvalid && bv <= av;
                const bool eq = takeB && bv == av;
                if (takeB) { b++; bv = B[b < s ? b : s - 1]; }
                else { a++; av = A[a < s ? a : s - 1]; }
                uni += eq ? 0 : 1;
                inter += eq ? 1 : 0;
            }
            // :453-454
            double jac = fmax(double(inter), 1.0) / uni;
            double d = fmin(1.0, fabs(log(2.0 * jac / (1.0 + jac)) / p.k));
            if (p.tri) {
                p.out[(size_t)i * p.ld + j] = d;
                p.out[(size_t)j * p.ld + i] = d;
            } else {
                p.out[(size_t)(i - p.row_off) * p.ld + j] = d;
            }
        }
        __syncthreads();   // tile buffers are reused
    }
}

// Warp-per-pair variant (default for sketch sizes up to 1024): the thread-per-pair merge above is a chain of up to 2s
// dependent shared-memory loads with only 4 warps per SM (the tile fills shared memory).  Here the same 16 x 8 tile is
// worked on by 16 warps and each pair's merge is cut into 32 diagonal chunks (merge path): lane l finds by binary search
// the (a, b) at which the merged order -- B before A on ties, exactly the order of the reference's loop -- reaches
// element l * CH, walks its CH <= 64 elements and records one bit per element (union event / intersection event).
// The reference's stop rule "the union counter reaches s" becomes: find the s-th union bit in lane order and count the
// intersection bits before it.  Same integers as src/mash.cu:437-454, so the distances are bit-identical.
constexpr int MW_WARPS = 16;
constexpr int MW_THREADS = MW_WARPS * 32;

__global__ void __launch_bounds__(MW_THREADS, 1) mash_warp_kernel(MashTileParams p, long long num_tiles, int tiles_x) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* sA = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* sB = sA + (size_t)MD_TA * p.s;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + (size_t)MD_TB * p.s);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncthreads();
    uint32_t phase = 0;
    const int s = p.s, total = 2 * s, CH = (total + 31) >> 5;
    const uint32_t sk_bytes = (uint32_t)s * 8u;
    for (long long t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int tb = (int)(t / tiles_x), ta = (int)(t % tiles_x);
        const int row0 = p.r0 + tb * MD_TB, col0 = ta * MD_TA;
        const int row_max = min(row0 + MD_TB, p.r1) - 1;
        if (p.tri && col0 > row_max) continue;   // strictly above the diagonal (uniform per CTA)
        if (tid == 0) {
            int na = min(MD_TA, p.n - col0), nb = min(MD_TB, p.r1 - row0);
            mbar_arrive_expect_tx(bar, (uint32_t)(na + nb) * sk_bytes);
            for (int q = 0; q < na; q++) tma_bulk_g2s(sA + (size_t)q * s, p.sk + (size_t)(col0 + q) * s, sk_bytes, bar);
            for (int q = 0; q < nb; q++) tma_bulk_g2s(sB + (size_t)q * s, p.sk + (size_t)(row0 + q) * s, sk_bytes, bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        for (int pp = wid; pp < MD_TA * MD_TB; pp += MW_WARPS) {
            const int ia = pp % MD_TA, ib = pp / MD_TA;
            const int i = row0 + ib, j = col0 + ia;
            const int jlim = p.tri ? i : p.ncols;
            if (!(i < p.r1 && j < jlim && j < p.n)) continue;   // warp-uniform
            const uint64_t* A = sA + (size_t)ia * s;
            const uint64_t* B = sB + (size_t)ib * s;
            const int d0 = min(lane * CH, total), d1 = min(d0 + CH, total);
            // merge-path split: a = how many A elements are among the first d0 merged ones
            int lo = max(0, d0 - s), hi = min(d0, s);
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (A[mid] < B[d0 - mid - 1]) lo = mid + 1; else hi = mid;
            }
            int a = lo, b = d0 - lo;
            uint64_t av = A[min(a, s - 1)], bv = B[min(b, s - 1)];
            uint32_t um[2] = {0u, 0u}, im[2] = {0u, 0u};
            const int steps = d1 - d0;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int cnt = min(32, steps - 32 * h);
                uint32_t u = 0, m = 0;
                for (int q = 0; q < cnt; q++) {
                    const bool takeB = (b < s) && (a >= s || bv <= av);
                    const bool eq = takeB && a < s && bv == av;
                    const uint32_t bit = 1u << q;
                    u |= eq ? 0u : bit;
                    m |= eq ? bit : 0u;
                    a += takeB ? 0 : 1;
                    b += takeB ? 1 : 0;
                    const uint64_t v = takeB ? B[min(b, s - 1)] : A[min(a, s - 1)];
                    if (takeB) bv = v; else av = v;
                }
                um[h] = u; im[h] = m;
            }
            const int uc = __popc(um[0]) + __popc(um[1]);
            int P = uc;
#pragma unroll
            for (int sh = 1; sh < 32; sh <<= 1) { const int o = __shfl_up_sync(0xffffffffu, P, sh); if (lane >= sh) P += o; }
            const unsigned int reached = __ballot_sync(0xffffffffu, P >= s);   // never empty: the A list alone holds s union events
            const int Ls = __ffs(reached) - 1;
            int contrib = 0;
            if (lane < Ls) contrib = __popc(im[0]) + __popc(im[1]);
            else if (lane == Ls) {
                const int tth = s - (P - uc);             // the s-th union event overall is this lane's tth (1-based)
                const int c0 = __popc(um[0]);
                const bool low = tth <= c0;
                uint32_t m = low ? um[0] : um[1];
                const int skip = (low ? tth : tth - c0) - 1;
                for (int r = 0; r < skip; r++) m &= m - 1;
                const int pos = __ffs(m) - 1;              // intersection events after it are never consumed by the reference loop
                const uint32_t below = pos ? (0xffffffffu >> (32 - pos)) : 0u;
                contrib = low ? __popc(im[0] & below) : __popc(im[0]) + __popc(im[1] & below);
            }
            const int inter = __reduce_add_sync(0xffffffffu, contrib);
            if (lane == 0) {
                // :453-454 (the union counter is exactly s on exit)
                double jac = fmax(double(inter), 1.0) / s;
                double d = fmin(1.0, fabs(log(2.0 * jac / (1.0 + jac)) / p.k));
                if (p.tri) {
                    p.out[(size_t)i * p.ld + j] = d;
                    p.out[(size_t)j * p.ld + i] = d;
                } else {
                    p.out[(size_t)(i - p.row_off) * p.ld + j] = d;
                }
            }
        }
        __syncthreads();   // tile buffers are reused
    }
}

// Rank-compressed variant (default while n * s < 2^32 - 1 and the tile fits shared memory).  The merge only asks "is
// B[b] <= A[a]" and "is B[b] == A[a]", so every 64-bit hash can be replaced by its dense rank among ALL n * s hashes of the
// data set (one radix sort + scan + scatter, mash_build_ranks): same order, same equalities, hence the same inter / union
// integers -- but 4-byte keys.  That halves the shared-memory bytes per step, lets a tile hold 32 column sketches x 16
// row sketches (512 pairs, 16 warps instead of 4) and, above all, allows a bank-conflict-free layout: the tile is stored
// INTERLEAVED (element t of column sketch q at word t * 32 + q, of row sketch q at t * 16 + q) and lane l of warp r works
// on the pair (column l, row (l + r) mod 16).  Every lane then reads its column list through its own bank whatever its
// position, and at most two lanes meet on a row-list bank -- the thread-per-pair kernel above spends ~5.5 shared-memory
// wavefronts per step on random 8-byte reads.  A few extra rows of 0xFFFFFFFF after each list replace the index clamps.
constexpr int MR_TA = 32;   // column sketches ("A" lists) per tile = lanes
constexpr int MR_PAD = 6;   // 0xFFFFFFFF rows after each list

// MR_TB = row sketches ("B" lists) per tile = warps: 24 when the tile fits shared memory (s <= 1031: 6 warps per scheduler hide
// more of the compare -> select -> load chain than 4), else 16
template <int MR_TB>
__global__ void __launch_bounds__(MR_TB * 32, 1) mash_rank_kernel(MashTileParams p, const uint32_t* __restrict__ rk, long long num_tiles, int tiles_x) {
    constexpr int MR_THREADS = MR_TB * 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int s = p.s;
    uint32_t* sA = reinterpret_cast<uint32_t*>(smem_raw);          // [(s + MR_PAD)][32]
    uint32_t* sB = sA + (size_t)(s + MR_PAD) * MR_TA;              // [(s + MR_PAD)][16]
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid < MR_PAD * MR_TA) sA[(size_t)s * MR_TA + tid] = 0xFFFFFFFFu;   // rows past the end: the look-ahead and the up to 3 discarded steps read them
    if (tid < MR_PAD * MR_TB) sB[(size_t)s * MR_TB + tid] = 0xFFFFFFFFu;
    const int half = s >> 1;                                       // sketch sizes are even: 8-byte global loads
    for (long long t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int tb = (int)(t / tiles_x), ta = (int)(t % tiles_x);
        const int row0 = p.r0 + tb * MR_TB, col0 = ta * MR_TA;
        const int row_max = min(row0 + MR_TB, p.r1) - 1;
        if (p.tri && col0 > row_max) continue;   // strictly above the diagonal (uniform per CTA)
        __syncthreads();                         // the previous tile is done with the buffers
        {
            // columns: lane = sketch, the 16 warps split the element pairs; stores hit bank = lane
            const int q = col0 + lane;
            if (q < p.n) {
                const uint2* src = reinterpret_cast<const uint2*>(rk + (size_t)q * s);
#pragma unroll 8
                for (int h = wid; h < half; h += MR_TB) {
                    const uint2 v = __ldg(src + h);
                    sA[(size_t)(2 * h) * MR_TA + lane] = v.x;
                    sA[(size_t)(2 * h + 1) * MR_TA + lane] = v.y;
                }
            }
            // rows: consecutive threads = consecutive sketches of one element pair (conflict-free stores)
#pragma unroll 4
            for (int e = tid; e < MR_TB * half; e += MR_THREADS) {
                const int qb = e % MR_TB, h = e / MR_TB, r = row0 + qb;
                if (r < p.r1) {
                    const uint2 v = __ldg(reinterpret_cast<const uint2*>(rk + (size_t)r * s) + h);
                    sB[(size_t)(2 * h) * MR_TB + qb] = v.x;
                    sB[(size_t)(2 * h + 1) * MR_TB + qb] = v.y;
                }
            }
        }
        __syncthreads();
        const int ib = (lane + wid) % MR_TB;
        const int i = row0 + ib, j = col0 + lane;
        const int jlim = p.tri ? i : p.ncols;
        if (i < p.r1 && j < jlim && j < p.n) {
            // src/mash.cu:439-450, one consumed element per step.  Current and next element of both lists live in registers, so
            // the shared-memory load of a step (the element after next of the consumed list) is off the compare -> select
            // critical path; an exhausted B list reads 0xFFFFFFFF (> every rank).  Every step consumes one element, so
            // inter = steps - uni.
            uint32_t pa = (uint32_t)__cvta_generic_to_shared(sA + lane), pb = (uint32_t)__cvta_generic_to_shared(sB + ib);
            int uni = 0, steps = 0;
            uint32_t av, an, bv, bn;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(av) : "r"(pa));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(an) : "r"(pa + MR_TA * 4));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(bv) : "r"(pb));
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(bn) : "r"(pb + MR_TB * 4));
            // The element loaded in a step is committed at the top of the NEXT step (warps issue in order: consuming it in
            // the same step would park the warp on the load).  It becomes "next" then and "current" one step later at the
            // earliest, so nothing is compared before it has arrived.  Four steps per loop trip: the steps past the one
            // that brings the union counter to s only read the 0xFFFFFFFF rows and are discarded.
            uint32_t v = bn;
            bool pt = true;
            auto step = [&](int u) -> int {
                if (pt) bn = v; else an = v;
                const bool takeB = bv <= av;
                u += (bv == av) ? 0 : 1;
                const uint32_t addr = takeB ? pb + 2 * MR_TB * 4 : pa + 2 * MR_TA * 4;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
                if (takeB) { bv = bn; pb += MR_TB * 4; }
                else { av = an; pa += MR_TA * 4; }
                pt = takeB;
                return u;
            };
            for (;;) {
                const int u1 = step(uni), u2 = step(u1), u3 = step(u2), u4 = step(u3);
                if (u4 >= s) { steps += u1 >= s ? 1 : (u2 >= s ? 2 : (u3 >= s ? 3 : 4)); uni = s; break; }
                uni = u4; steps += 4;
            }
            const int inter = steps - uni;
            // :453-454
            double jac = fmax(double(inter), 1.0) / uni;
            double d = fmin(1.0, fabs(log(2.0 * jac / (1.0 + jac)) / p.k));
            if (p.tri) {
                p.out[(size_t)i * p.ld + j] = d;
                p.out[(size_t)j * p.ld + i] = d;
            } else {
                p.out[(size_t)(i - p.row_off) * p.ld + j] = d;
            }
        }
    }
}

__global__ void iota_u32_kernel(uint32_t* v, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) v[i] = (uint32_t)i;
}
__global__ void rank_flag_kernel(const uint64_t* __restrict__ sorted, uint32_t* __restrict__ flag, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        flag[i] = (i > 0 && sorted[i] != sorted[i - 1]) ? 1u : 0u;
}
__global__ void rank_scatter_kernel(const uint32_t* __restrict__ rank_sorted, const uint32_t* __restrict__ idx, uint32_t* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[idx[i]] = rank_sorted[i];
}

// dense ranks of all n * s hashes: radix sort (hash, position), mark value changes, prefix sum, scatter back
static int mash_build_ranks(dipb_mash* m) {
    dipb_ctx* c = m->ctx;
    const size_t M = (size_t)m->n * m->s;
    uint64_t* keys_out = nullptr;
    uint32_t *idx = nullptr, *idx_out = nullptr, *flag = nullptr;
    void* tmp = nullptr;
    size_t tmp_sort = 0, tmp_scan = 0;
    DIPB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, M, 0, 64, c->stream));
    DIPB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tmp_scan, (const uint32_t*)nullptr, (uint32_t*)nullptr, M, c->stream));
    const size_t tmp_bytes = tmp_sort > tmp_scan ? tmp_sort : tmp_scan;
    DIPB_CUDA(cudaMalloc(&m->ranks, M * sizeof(uint32_t)));   // (lives as long as the sketches)
    DIPB_CUDA(pool_alloc(c, (void**)&keys_out, M * sizeof(uint64_t)));
    DIPB_CUDA(pool_alloc(c, (void**)&idx, M * sizeof(uint32_t)));
    DIPB_CUDA(pool_alloc(c, (void**)&idx_out, M * sizeof(uint32_t)));
    DIPB_CUDA(pool_alloc(c, (void**)&flag, M * sizeof(uint32_t)));
    DIPB_CUDA(pool_alloc(c, &tmp, tmp_bytes ? tmp_bytes : 16));
    const int grid = c->num_sms * 8;
    iota_u32_kernel<<<grid, 256, 0, c->stream>>>(idx, M);
    DIPB_KERNEL_CHECK(c);
    size_t tb = tmp_bytes;
    DIPB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, (const uint64_t*)m->sketches, keys_out, (const uint32_t*)idx, idx_out, M, 0, 64, c->stream));
    rank_flag_kernel<<<grid, 256, 0, c->stream>>>(keys_out, flag, M);
    DIPB_KERNEL_CHECK(c);
    tb = tmp_bytes;
    DIPB_CUDA(cub::DeviceScan::InclusiveSum(tmp, tb, (const uint32_t*)flag, idx, M, c->stream));   // idx now holds the rank of sorted position i
    rank_scatter_kernel<<<grid, 256, 0, c->stream>>>(idx, idx_out, m->ranks, M);
    DIPB_KERNEL_CHECK(c);
    c->launches += 4;   // the library's sort and scan passes (approximate; not hand-written kernels)
    pool_free(c, keys_out); pool_free(c, idx); pool_free(c, idx_out); pool_free(c, flag); pool_free(c, tmp);
    return 0;
}

__global__ void zero_diag_kernel(double* D, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) D[(size_t)i * n + i] = 0.0;
}

static int mash_launch(dipb_mash* m, MashTileParams p) {
    dipb_ctx* c = m->ctx;
    const int rows = p.r1 - p.r0;
    const int ncols = p.tri ? p.r1 : p.ncols;
    if (rows <= 0 || ncols <= 0) return 0;
    // ---- rank-compressed, interleaved tiles (default)
    const char* er = getenv("DIPB_MASH_RANKS");   // 0: keep the 64-bit hashes (first versions, kept for comparison)
    // 24-row tiles for whole (triangular) matrices when the sketches fit: measured 38.4 vs 42.6 ms at C2; 16-row tiles for the
    // row blocks of placement / D&C, where the 24-row tile measured 14 % SLOWER (120 000 tips: 9.3 vs 8.1 s).  DIPB_MASH_TB=16
    // / =24 forces one
    const char* etb = getenv("DIPB_MASH_TB");
    const bool fits24 = (size_t)(m->s + MR_PAD) * (MR_TA + 24) * sizeof(uint32_t) <= 227 * 1024;
    int TB = fits24 && (etb ? atoi(etb) == 24 : p.tri != 0) ? 24 : 16;
    const size_t rk_smem = (size_t)(m->s + MR_PAD) * (MR_TA + TB) * sizeof(uint32_t);
    if (!(er && atoi(er) == 0) && (size_t)m->n * m->s < 0xFFFFFFFFull && rk_smem <= 227 * 1024) {
        if (!m->ranks) { int rc = mash_build_ranks(m); if (rc) return rc; }
        const int tiles_y = (rows + TB - 1) / TB, tiles_x = (ncols + MR_TA - 1) / MR_TA;
        const long long tiles = (long long)tiles_y * tiles_x;
        const int grid = (int)(tiles < (long long)c->num_sms * 4 ? tiles : (long long)c->num_sms * 4);
        auto go = [&](auto kern, int threads) -> int {
            DIPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rk_smem));   // (per device)
            kern<<<grid, threads, rk_smem, c->stream>>>(p, m->ranks, tiles, tiles_x);
            return 0;
        };
        const int rc = TB == 24 ? go(mash_rank_kernel<24>, 24 * 32) : go(mash_rank_kernel<16>, 16 * 32);
        if (rc) return rc;
        DIPB_KERNEL_CHECK(c);
        return 0;
    }
    // ---- 64-bit hashes
    size_t smem = (size_t)(MD_TA + MD_TB) * m->s * 8 + 64;
    if (smem > 227 * 1024) { set_error("mash distance: sketch size %d does not fit shared memory tiles", m->s); return DIPB_E_ARG; }
    DIPB_CUDA(cudaFuncSetAttribute(mash_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // (per device)
    DIPB_CUDA(cudaFuncSetAttribute(mash_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const char* ew = getenv("DIPB_MASH_WARP");   // 1: merge-path warp-per-pair kernel (measured equal to the thread-per-pair one: both sit on
    const bool warp_merge = m->s <= 1024 && ew && atoi(ew) == 1;   // the shared-memory wavefronts of random 8-byte reads)
    int tiles_y = (rows + MD_TB - 1) / MD_TB, tiles_x = (ncols + MD_TA - 1) / MD_TA;
    long long tiles = (long long)tiles_y * tiles_x;
    int grid = (int)(tiles < (long long)c->num_sms * 8 ? tiles : (long long)c->num_sms * 8);
    if (warp_merge) mash_warp_kernel<<<grid, MW_THREADS, smem, c->stream>>>(p, tiles, tiles_x);
    else mash_tile_kernel<<<grid, MD_THREADS, smem, c->stream>>>(p, tiles, tiles_x);
    DIPB_KERNEL_CHECK(c);
    return 0;
}

}  // namespace dipb

using namespace dipb;

extern "C" {

int dipb_mash_upload_flat(dipb_ctx* c, const uint64_t* flat, const uint64_t* word_off, const uint64_t* len, size_t n,
                          int k, int s, dipb_mash** out) {
    if (!c || !flat || !word_off || !len || !out || n == 0) { set_error("dipb_mash_upload_flat: bad argument"); return DIPB_E_ARG; }
    if (k < 2 || k > 32) { set_error("dipb_mash_upload: k-mer size %d outside 2..32", k); return DIPB_E_ARG; }
    if (s < 2 || s > SK_BUF - SK_THREADS || (s & 1)) { set_error("dipb_mash_upload: sketch size %d must be even and within 2..%d", s, SK_BUF - SK_THREADS); return DIPB_E_ARG; }
    DIPB_CUDA(cudaSetDevice(c->device));
    dipb_mash* m = new dipb_mash();
    m->ctx = c; m->n = (int)n; m->k = k; m->s = s;
    ctx_retain(c);
    size_t words = word_off[n - 1] + (len[n - 1] + 31) / 32;
    auto body = [&]() -> int {
        DIPB_CUDA(cudaMalloc(&m->seqs, (words + 1) * 8));
        DIPB_CUDA(cudaMalloc(&m->word_off, n * 8));
        DIPB_CUDA(cudaMalloc(&m->lens, n * 8));
        DIPB_CUDA(cudaMalloc(&m->sketches, n * (size_t)s * 8));
        DIPB_CUDA(cudaMemsetAsync(m->seqs + words, 0, 8, c->stream));
        DIPB_CUDA(cudaMemcpyAsync(m->seqs, flat, words * 8, cudaMemcpyHostToDevice, c->stream));
        DIPB_CUDA(cudaMemcpyAsync(m->word_off, word_off, n * 8, cudaMemcpyHostToDevice, c->stream));
        DIPB_CUDA(cudaMemcpyAsync(m->lens, len, n * 8, cudaMemcpyHostToDevice, c->stream));
        DIPB_CUDA(cudaStreamSynchronize(c->stream));
        return 0;
    };
    const int rc = body();
    if (rc) { dipb_mash_free(m); return rc; }
    *out = m;
    return 0;
}

int dipb_mash_upload(dipb_ctx* c, const uint64_t* const* seq2, const uint64_t* len, size_t n, int k, int s, dipb_mash** out) {
    if (!c || !seq2 || !len || !out || n == 0) { set_error("dipb_mash_upload: bad argument"); return DIPB_E_ARG; }
    std::vector<uint64_t> off(n);
    size_t tot = 0;
    for (size_t i = 0; i < n; i++) { off[i] = tot; tot += (len[i] + 31) / 32; }
    std::vector<uint64_t> flat(tot ? tot : 1);
    for (size_t i = 0; i < n; i++) memcpy(flat.data() + off[i], seq2[i], ((len[i] + 31) / 32) * 8);
    return dipb_mash_upload_flat(c, flat.data(), off.data(), len, n, k, s, out);
}

int dipb_mash_set_sketches(dipb_ctx* c, const uint64_t* h_sk, size_t n, int k, int s, dipb_mash** out) {
    if (!c || !h_sk || !out || n == 0 || s < 2 || (s & 1)) { set_error("dipb_mash_set_sketches: bad argument (sketch size must be even)"); return DIPB_E_ARG; }
    DIPB_CUDA(cudaSetDevice(c->device));
    dipb_mash* m = new dipb_mash();
    m->ctx = c; m->n = (int)n; m->k = k; m->s = s;
    ctx_retain(c);
    if (cudaMalloc(&m->sketches, n * (size_t)s * 8) != cudaSuccess ||
        cudaMemcpy(m->sketches, h_sk, n * (size_t)s * 8, cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("dipb_mash_set_sketches: device allocation / copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        dipb_mash_free(m);
        return DIPB_E_CUDA;
    }
    m->sketched = true;
    *out = m;
    return 0;
}

void dipb_mash_free(dipb_mash* m) {
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    cudaFree(m->seqs); cudaFree(m->word_off); cudaFree(m->lens); cudaFree(m->sketches); cudaFree(m->ranks);
    ctx_release(m->ctx);
    delete m;
}

int dipb_mash_sketch(dipb_mash* m) {
    if (!m || !m->seqs) { set_error("dipb_mash_sketch: no sequences uploaded"); return DIPB_E_STATE; }
    dipb_ctx* c = m->ctx;
    DIPB_CUDA(cudaSetDevice(c->device));
    int rc = timer_begin(c);
    if (rc) return rc;
    int grid = m->n < c->num_sms * 2 ? m->n : c->num_sms * 2;
    sketch_kernel<<<grid, SK_THREADS, 0, c->stream>>>(m->seqs, m->word_off, m->lens, m->n, m->k, m->s, m->sketches);
    DIPB_KERNEL_CHECK(c);
    rc = timer_end(c, DIPB_T_SKETCH);
    if (rc) return rc;
    if (m->ranks) { cudaFree(m->ranks); m->ranks = nullptr; }   // ranks belong to the previous sketches
    m->sketched = true;
    return 0;
}

int dipb_mash_get_sketches(dipb_mash* m, uint64_t* h_out) {
    if (!m || !h_out) { set_error("dipb_mash_get_sketches: bad argument"); return DIPB_E_ARG; }
    if (!m->sketched) { set_error("dipb_mash_get_sketches: call dipb_mash_sketch first"); return DIPB_E_STATE; }
    DIPB_CUDA(cudaSetDevice(m->ctx->device));
    DIPB_CUDA(cudaStreamSynchronize(m->ctx->stream));
    DIPB_CUDA(cudaMemcpy(h_out, m->sketches, (size_t)m->n * m->s * 8, cudaMemcpyDeviceToHost));
    return 0;
}

int dipb_mash_dist_block(dipb_mash* m, int r0, int r1, int ncols, double* d_out, size_t ld) {
    if (!m || !d_out || r0 < 0 || r1 > m->n || r0 >= r1 || ncols < 0 || ncols > m->n) { set_error("dipb_mash_dist_block: bad argument"); return DIPB_E_ARG; }
    if (!m->sketched) { set_error("dipb_mash_dist_block: sketches not built (call dipb_mash_sketch)"); return DIPB_E_STATE; }
    if (ncols == 0) return 0;
    DIPB_CUDA(cudaSetDevice(m->ctx->device));
    MashTileParams p{};
    p.sk = m->sketches; p.n = m->n; p.s = m->s; p.k = m->k; p.tri = 0; p.r0 = r0; p.r1 = r1; p.ncols = ncols;
    p.out = d_out; p.ld = ld; p.row_off = r0;
    return mash_launch(m, p);
}

int dipb_mash_dist_row(dipb_mash* m, int row, double* d_out) {
    if (!m || !d_out || row < 0 || row >= m->n) { set_error("dipb_mash_dist_row: bad argument"); return DIPB_E_ARG; }
    if (row == 0) return 0;
    int rc = dipb_mash_dist_block(m, row, row + 1, row, d_out, (size_t)m->n);
    if (rc) return rc;
    DIPB_CUDA(cudaStreamSynchronize(m->ctx->stream));
    return 0;
}

int dipb_mash_dist_row_host(dipb_mash* m, int row, double* h_out) {
    if (!m || !h_out || row < 0 || row >= m->n) { set_error("dipb_mash_dist_row_host: bad argument"); return DIPB_E_ARG; }
    if (row == 0) return 0;
    DIPB_CUDA(cudaSetDevice(m->ctx->device));
    double* d = nullptr;
    DIPB_CUDA(cudaMalloc(&d, sizeof(double) * m->n));
    int rc = dipb_mash_dist_row(m, row, d);
    if (!rc && cudaMemcpy(h_out, d, sizeof(double) * row, cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("dipb_mash_dist_row_host: D2H failed"); rc = DIPB_E_CUDA; }
    cudaFree(d);
    return rc;
}

int dipb_mash_dist_matrix(dipb_mash* m, dipb_matrix** out) {
    if (!m || !out) { set_error("dipb_mash_dist_matrix: bad argument"); return DIPB_E_ARG; }
    if (!m->sketched) { set_error("dipb_mash_dist_matrix: sketches not built (call dipb_mash_sketch)"); return DIPB_E_STATE; }
    dipb_ctx* c = m->ctx;
    DIPB_CUDA(cudaSetDevice(c->device));
    dipb_matrix* M = new dipb_matrix();
    M->ctx = c; M->n = m->n;
    ctx_retain(c);
    size_t bytes = (size_t)m->n * m->n * sizeof(double);
    if (pool_alloc(c, (void**)&M->d, bytes) != cudaSuccess) { set_error("dipb_mash_dist_matrix: allocation of %zu bytes failed", bytes); M->d = nullptr; dipb_matrix_free(M); return DIPB_E_NOMEM; }
    int rc = timer_begin(c);
    MashTileParams p{};
    p.sk = m->sketches; p.n = m->n; p.s = m->s; p.k = m->k; p.tri = 1; p.r0 = 0; p.r1 = m->n; p.ncols = m->n;
    p.out = M->d; p.ld = (size_t)m->n; p.row_off = 0;
    if (!rc) rc = mash_launch(m, p);
    if (!rc) {
        zero_diag_kernel<<<(m->n + 255) / 256, 256, 0, c->stream>>>(M->d, m->n);
        c->launches++;
        rc = timer_end(c, DIPB_T_MASH_DIST);
    }
    if (rc) { dipb_matrix_free(M); return rc; }
    *out = M;
    return 0;
}

}  // extern "C"
