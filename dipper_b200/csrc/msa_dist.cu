// K1: aligned-MSA all-pairs distances on bit planes (sm_100a).
//
// Replaces MSADeviceArrays::{allocateDeviceArrays,distConstructionOnGpu}
// (reference src/MSA.cu:14-72,214-282), the D&C twins (DC/msa.cu:219-504) and the
// matrix build of NJDeviceArrays::getDismatrix/fillDismatrix
// (src/neighborJoining.cu:20-85).
//
// Design (not a port: the reference walks nibbles one block per pair):
//  * upload repacks the 4-bit stream once into three 1-bit planes per site
//    (b0 = code bit 0, b1 = code bit 1, v = code < 4 and site < L), stored
//    tile-blocked: [seq block of 128][k chunk of 16 words][plane][word][seq].
//    One (seq block, k chunk) slab is 24 KB contiguous, so a pipeline stage is two
//    1-D TMA bulk copies (cp.async.bulk -> UBLKCP) completing on an mbarrier.
//  * a persistent 512-thread CTA owns a 128 x 128 tile of pairs; each thread keeps
//    4 x 8 pairs in registers with (match, both-valid) packed 16+16 bits, and per
//    32-site word does 4 LOP3 + 2 POPC + 2 adds:
//        u = (a1^b1) | (a0^b0);  vv = va & vb;  match += popc(vv & ~u);  both += popc(vv)
//    useful = nv[i] + nv[j] - both (nv = per-sequence valid-site count).
//  * the epilogue applies p / JC in fp64 with the reference's expression order and
//    writes D[i][j] and the mirror D[j][i] (fillDismatrix fused away).
//  * models 3-6 and alignments longer than 65 024 sites go through the same tile
//    loop with other boolean functors, accumulating int32 counters per pair.
#include <cstdlib>
#include "common.cuh"
#include "msa.cuh"
#include "msa_pair.cuh"

namespace dipb {

// ---------------------------------------------------------------------------
// repack: 4-bit [n][comp64] -> blocked planes
// ---------------------------------------------------------------------------
__global__ void msa_repack_kernel(const uint64_t* __restrict__ in, int n, int seq_len, int comp64, int nkc,
                                  uint32_t* __restrict__ planes, int* __restrict__ nv, int npad) {
    // one thread per (32-site word w, sequence s); consecutive threads -> consecutive s
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int wtot = nkc * MSA_KC;
    if (gid >= (long long)wtot * npad) return;
    int s = (int)(gid % npad);
    int w = (int)(gid / npad);
    uint32_t p0 = 0, p1 = 0, pv = 0;
    if (s < n) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            int w64 = 2 * w + h;
            if (w64 >= comp64) continue;
            uint64_t x = in[(size_t)s * comp64 + w64];
#pragma unroll
            for (int t = 0; t < 16; t++) {
                int site = w64 * 16 + t;
                uint32_t code = (uint32_t)(x >> (4 * t)) & 15u;
                uint32_t ok = (code < 4u && site < seq_len) ? 1u : 0u;
                int bit = h * 16 + t;
                p0 |= (ok & code & 1u) << bit;
                p1 |= (ok & (code >> 1) & 1u) << bit;
                pv |= ok << bit;
            }
        }
        if (pv) atomicAdd(&nv[s], __popc(pv));
    }
    int sb = s / MSA_TS, sl = s % MSA_TS, kc = w / MSA_KC, kk = w % MSA_KC;
    size_t base = ((size_t)sb * nkc + kc) * (3 * MSA_KC * MSA_TS);
    planes[base + (0 * MSA_KC + kk) * MSA_TS + sl] = p0;
    planes[base + (1 * MSA_KC + kk) * MSA_TS + sl] = p1;
    planes[base + (2 * MSA_KC + kk) * MSA_TS + sl] = pv;
}

struct TileParams {
    const uint32_t* planes;
    const int* nv;
    int nkc;        // chunks per sequence
    int kc0, kc1;   // chunk range of this launch
    int tri;        // 1: lower-triangle tiles with mirror, 0: rectangle
    int bi0, nbi;   // row blocks
    int nbj;        // column blocks (rectangle)
    int n;          // number of sequences
    // fast path output
    double* out;
    size_t ld;
    int row_lo, row_hi;  // rows written: row_lo <= i < row_hi
    int col_hi;          // columns written: j < col_hi
    int row_off;         // rectangle: out row = i - row_off
    int dist_type;
    // generic path output (int32 accumulators, rectangle only)
    int* c1;
    int* c2;
    int accumulate;
};

__device__ __forceinline__ void tile_of(const TileParams& p, long long t, int& bi, int& bj) {
    if (p.tri) {
        // tiles of row blocks bi0.. : linear index over rows, row bi has bi+1 tiles
        long long base = (long long)p.bi0 * (p.bi0 + 1) / 2;
        long long g = base + t;
        long long b = (long long)((sqrt(8.0 * (double)g + 1.0) - 1.0) * 0.5);
        while (b * (b + 1) / 2 > g) b--;
        while ((b + 1) * (b + 2) / 2 <= g) b++;
        bi = (int)b;
        bj = (int)(g - b * (b + 1) / 2);
    } else {
        bi = p.bi0 + (int)(t / p.nbj);
        bj = (int)(t % p.nbj);
    }
}

template <int FN, bool FAST>
__global__ void __launch_bounds__(MSA_THREADS, 1) msa_tile_kernel(TileParams p, long long num_tiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t* stage_base = reinterpret_cast<uint32_t*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)MSA_STAGES * MSA_STAGE_BYTES);

    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    if (tid == 0) {
        for (int s = 0; s < MSA_STAGES; s++) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int nchunk = p.kc1 - p.kc0;
    long long my_tiles = 0;
    if ((long long)blockIdx.x < num_tiles) my_tiles = (num_tiles - 1 - blockIdx.x) / gridDim.x + 1;
    const long long total = my_tiles * nchunk;

    auto issue = [&](long long g) {
        long long tseq = g / nchunk;
        int kc = p.kc0 + (int)(g % nchunk);
        int bi, bj;
        tile_of(p, (long long)blockIdx.x + tseq * gridDim.x, bi, bj);
        int st = (int)(g % MSA_STAGES);
        uint32_t* dst = stage_base + (size_t)st * (MSA_STAGE_BYTES / 4);
        const uint32_t* srcA = p.planes + ((size_t)bi * p.nkc + kc) * MSA_SLAB_WORDS;
        const uint32_t* srcB = p.planes + ((size_t)bj * p.nkc + kc) * MSA_SLAB_WORDS;
        mbar_arrive_expect_tx(&full[st], MSA_STAGE_BYTES);
        tma_bulk_g2s(dst, srcA, MSA_SLAB_WORDS * 4, &full[st]);
        tma_bulk_g2s(dst + MSA_SLAB_WORDS, srcB, MSA_SLAB_WORDS * 4, &full[st]);
    };

    if (tid == 0) {
        for (long long g = 0; g < MSA_STAGES - 1 && g < total; g++) issue(g);
    }

    uint32_t acc[4][8];
    long long g = 0;
    for (long long tseq = 0; tseq < my_tiles; tseq++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 8; c++) acc[r][c] = 0;

        for (int ck = 0; ck < nchunk; ck++, g++) {
            if (tid == 0 && g + MSA_STAGES - 1 < total) issue(g + MSA_STAGES - 1);
            const int st = (int)(g % MSA_STAGES);
            mbar_wait(&full[st], (uint32_t)((g / MSA_STAGES) & 1));
            const uint32_t* A = stage_base + (size_t)st * (MSA_STAGE_BYTES / 4);
            const uint32_t* B = A + MSA_SLAB_WORDS;
#pragma unroll 2
            for (int kk = 0; kk < MSA_KC; kk++) {
                const uint4 a0 = *reinterpret_cast<const uint4*>(A + (0 * MSA_KC + kk) * MSA_TS + 4 * ty);
                const uint4 a1 = *reinterpret_cast<const uint4*>(A + (1 * MSA_KC + kk) * MSA_TS + 4 * ty);
                const uint4 av = *reinterpret_cast<const uint4*>(A + (2 * MSA_KC + kk) * MSA_TS + 4 * ty);
                const uint4 b0l = *reinterpret_cast<const uint4*>(B + (0 * MSA_KC + kk) * MSA_TS + 4 * tx);
                const uint4 b0h = *reinterpret_cast<const uint4*>(B + (0 * MSA_KC + kk) * MSA_TS + 64 + 4 * tx);
                const uint4 b1l = *reinterpret_cast<const uint4*>(B + (1 * MSA_KC + kk) * MSA_TS + 4 * tx);
                const uint4 b1h = *reinterpret_cast<const uint4*>(B + (1 * MSA_KC + kk) * MSA_TS + 64 + 4 * tx);
                const uint4 bvl = *reinterpret_cast<const uint4*>(B + (2 * MSA_KC + kk) * MSA_TS + 4 * tx);
                const uint4 bvh = *reinterpret_cast<const uint4*>(B + (2 * MSA_KC + kk) * MSA_TS + 64 + 4 * tx);
                const uint32_t ra0[4] = {a0.x, a0.y, a0.z, a0.w};
                const uint32_t ra1[4] = {a1.x, a1.y, a1.z, a1.w};
                const uint32_t rav[4] = {av.x, av.y, av.z, av.w};
                const uint32_t rb0[8] = {b0l.x, b0l.y, b0l.z, b0l.w, b0h.x, b0h.y, b0h.z, b0h.w};
                const uint32_t rb1[8] = {b1l.x, b1l.y, b1l.z, b1l.w, b1h.x, b1h.y, b1h.z, b1h.w};
                const uint32_t rbv[8] = {bvl.x, bvl.y, bvl.z, bvl.w, bvh.x, bvh.y, bvh.z, bvh.w};
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        uint32_t m1, m2;
                        pair_masks<FN>(ra0[r], ra1[r], rav[r], rb0[c], rb1[c], rbv[c], m1, m2);
                        acc[r][c] += (uint32_t)__popc(m1) + ((uint32_t)__popc(m2) << 16);
                    }
            }
            __syncthreads();  // everyone is done with stage st before it is refilled
        }

        // ---- epilogue for this tile ----
        int bi, bj;
        tile_of(p, (long long)blockIdx.x + tseq * gridDim.x, bi, bj);
        const int i0 = bi * MSA_TS + 4 * ty;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int i = i0 + r;
            if (i >= p.n) continue;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const int j = bj * MSA_TS + (c < 4 ? 4 * tx + c : 64 + 4 * tx + (c - 4));
                if (j >= p.n) continue;
                const int k1 = (int)(acc[r][c] & 0xffffu), k2 = (int)(acc[r][c] >> 16);
                if (FAST) {
                    double d = 0.0;
                    if (i != j) {
                        int useful = p.nv[i] + p.nv[j] - k2;
                        d = dist_p_jc(k1, useful, p.dist_type);
                    }
                    if (p.tri) {
                        if (bi == bj) {
                            p.out[(size_t)i * p.ld + j] = d;
                        } else {
                            if (i >= p.row_lo && i < p.row_hi) p.out[(size_t)i * p.ld + j] = d;
                            if (i >= p.row_lo && i < p.row_hi) p.out[(size_t)j * p.ld + i] = d;
                        }
                    } else {
                        if (i >= p.row_lo && i < p.row_hi && j < p.col_hi) p.out[(size_t)(i - p.row_off) * p.ld + j] = d;
                    }
                } else {
                    if (i >= p.row_lo && i < p.row_hi && j < p.col_hi) {
                        size_t o = (size_t)(i - p.row_off) * p.ld + j;
                        if (p.accumulate) {
                            p.c1[o] += k1;
                            p.c2[o] += k2;
                        } else {
                            p.c1[o] = k1;
                            p.c2[o] = k2;
                        }
                    }
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// generic finalisation: distance from int32 counters (models 1-6)
// ---------------------------------------------------------------------------
struct StatPtrs {
    const int* c[18];  // FN f -> c[2f], c[2f+1]
};

__global__ void msa_finalize_kernel(StatPtrs sp, const int* __restrict__ nv, int type, int r0, int r1, int ncols,
                                    size_t ld_cnt, double* out, size_t ld_out, int row_off, int mirror) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int i = r0 + blockIdx.y;
    if (i >= r1 || j >= ncols) return;
    size_t o = (size_t)(i - r0) * ld_cnt + j;
    int match = sp.c[0][o], both = sp.c[1][o];
    int ts = 0, tv = 0, gcr = 0, gcc = 0, frac[4] = {0, 0, 0, 0}, pr[4] = {0, 0, 0, 0};
    if (type == DIPB_DIST_K2P || type == DIPB_DIST_JINNEI || type == DIPB_DIST_TAMURA) { ts = sp.c[2][o]; tv = sp.c[3][o]; }
    if (type == DIPB_DIST_TAMURA) { gcr = sp.c[4][o]; gcc = sp.c[5][o]; }
    if (type == DIPB_DIST_TAJIMANEI) {
        for (int c = 0; c < 4; c++) frac[c] = sp.c[6 + 2 * c][o] + sp.c[7 + 2 * c][o];
        pr[0] = sp.c[14][o]; pr[1] = sp.c[15][o]; pr[2] = sp.c[16][o]; pr[3] = sp.c[17][o];
    }
    double d = (i == j) ? 0.0 : dist_from_counts(type, match, both, nv[i], nv[j], ts, tv, gcr, gcc, frac, pr);
    if (mirror) {
        if (j > i) return;
        out[(size_t)i * ld_out + j] = d;
        out[(size_t)j * ld_out + i] = d;
    } else {
        out[(size_t)(i - row_off) * ld_out + j] = d;
    }
}

template <int FN>
static int launch_generic(dipb_msa* m, TileParams p, long long tiles) {
    dipb_ctx* c = m->ctx;
    int grid = (int)(tiles < c->num_sms ? tiles : c->num_sms);
    if (grid < 1) return 0;
    // (per device and cheap: set on every launch, several devices / host threads may use the library, csrc/multi.cu)
    DIPB_CUDA(cudaFuncSetAttribute(msa_tile_kernel<FN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MSA_SMEM_BYTES));
    msa_tile_kernel<FN, false><<<grid, MSA_THREADS, MSA_SMEM_BYTES, c->stream>>>(p, tiles);
    DIPB_KERNEL_CHECK(c);
    return 0;
}

static int launch_generic_fn(dipb_msa* m, int fn, const TileParams& p, long long tiles) {
    switch (fn) {
        case 0: return launch_generic<0>(m, p, tiles);
        case 1: return launch_generic<1>(m, p, tiles);
        case 2: return launch_generic<2>(m, p, tiles);
        case 3: return launch_generic<3>(m, p, tiles);
        case 4: return launch_generic<4>(m, p, tiles);
        case 5: return launch_generic<5>(m, p, tiles);
        case 6: return launch_generic<6>(m, p, tiles);
        case 7: return launch_generic<7>(m, p, tiles);
        default: return launch_generic<8>(m, p, tiles);
    }
}

static int launch_fast(dipb_msa* m, TileParams p, long long tiles) {
    dipb_ctx* c = m->ctx;
    int grid = (int)(tiles < c->num_sms ? tiles : c->num_sms);
    if (grid < 1) return 0;
    DIPB_CUDA(cudaFuncSetAttribute(msa_tile_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MSA_SMEM_BYTES));
    msa_tile_kernel<0, true><<<grid, MSA_THREADS, MSA_SMEM_BYTES, c->stream>>>(p, tiles);
    DIPB_KERNEL_CHECK(c);
    return 0;
}

// which functors a model needs
static int model_fns(int type, int* fns) {
    int k = 0;
    fns[k++] = 0;
    if (type == DIPB_DIST_K2P || type == DIPB_DIST_JINNEI || type == DIPB_DIST_TAMURA) fns[k++] = 1;
    if (type == DIPB_DIST_TAMURA) fns[k++] = 2;
    if (type == DIPB_DIST_TAJIMANEI) { for (int f = 3; f <= 8; f++) fns[k++] = f; }
    return k;
}

// rows [r0,r1) x cols [0,ncols) -> counters (rectangle), all k segments
static int generic_counts(dipb_msa* m, const int* fns, int nfn, int r0, int r1, int ncols, int** cnt, size_t ld) {
    TileParams p{};
    p.planes = m->planes; p.nv = m->nv; p.nkc = m->nkc; p.tri = 0; p.n = m->n;
    p.bi0 = r0 / MSA_TS; p.nbi = (r1 - 1) / MSA_TS - p.bi0 + 1; p.nbj = (ncols + MSA_TS - 1) / MSA_TS;
    p.row_lo = r0; p.row_hi = r1; p.col_hi = ncols; p.row_off = r0; p.ld = ld;
    long long tiles = (long long)p.nbi * p.nbj;
    for (int f = 0; f < nfn; f++) {
        p.c1 = cnt[2 * fns[f]]; p.c2 = cnt[2 * fns[f] + 1];
        for (int k0 = 0; k0 < m->nkc; k0 += MSA_MAX_CHUNKS) {
            p.kc0 = k0; p.kc1 = k0 + MSA_MAX_CHUNKS < m->nkc ? k0 + MSA_MAX_CHUNKS : m->nkc;
            p.accumulate = k0 > 0;
            int rc = launch_generic_fn(m, fns[f], p, tiles);
            if (rc) return rc;
        }
    }
    return 0;
}

// generic rectangle -> distances
static int generic_block(dipb_msa* m, int type, int r0, int r1, int ncols, double* out, size_t ld_out, int row_off,
                         int mirror) {
    dipb_ctx* c = m->ctx;
    int fns[9];
    int nfn = model_fns(type, fns);
    const int panel = 1024;
    size_t ld = ((size_t)(mirror ? r1 : ncols) + 127) / 128 * 128;
    int* buf = nullptr;
    DIPB_CUDA(cudaMalloc(&buf, (size_t)panel * ld * sizeof(int) * 2 * nfn));
    int* cnt[18] = {nullptr};
    StatPtrs sp{};
    for (int f = 0; f < nfn; f++) {
        cnt[2 * fns[f]] = buf + (size_t)(2 * f) * panel * ld;
        cnt[2 * fns[f] + 1] = buf + (size_t)(2 * f + 1) * panel * ld;
    }
    for (int q = 0; q < 18; q++) sp.c[q] = cnt[q];
    int rc = 0;
    for (int p0 = r0; p0 < r1 && !rc; p0 += panel) {
        int p1 = p0 + panel < r1 ? p0 + panel : r1;
        int nc = mirror ? p1 : ncols;
        rc = generic_counts(m, fns, nfn, p0, p1, nc, cnt, ld);
        if (rc) break;
        dim3 grid((nc + 255) / 256, p1 - p0);
        msa_finalize_kernel<<<grid, 256, 0, c->stream>>>(sp, m->nv, type, p0, p1, nc, ld, out, ld_out, row_off, mirror);
        c->launches++;
        if (cudaGetLastError() != cudaSuccess) { set_error("msa_finalize_kernel launch failed"); rc = DIPB_E_CUDA; }
    }
    cudaStreamSynchronize(c->stream);
    cudaFree(buf);
    return rc;
}

static bool fast_ok(const dipb_msa* m, int type) {
    return (type == DIPB_DIST_UNCORRECTED || type == DIPB_DIST_JC) && m->nkc <= MSA_MAX_CHUNKS;
}

int msa_block(dipb_msa* m, int type, int r0, int r1, int ncols, double* d_out, size_t ld) {
    if (r0 < 0 || r1 > m->n || r0 >= r1 || ncols < 0 || ncols > m->n) { set_error("msa_block: bad range"); return DIPB_E_ARG; }
    if (ncols == 0) return 0;
    if (msa_tc_supported(m, type)) {
        // Tensor-core kernel when there are enough rows to fill 128-row tiles and enough pairs to amortise the
        // launch set-up (tensor maps, tile list) and, on first use, the int8 operand expansion: placement row
        // blocks, D&C query batches against large backbones.  Small blocks stay on the popcount kernel (measured:
        // D&C of 30 000 tips with a 1 500-tip backbone 158 ms vs 340 ms).  DIPB_MSA_TC=0 never, =2 always (tests).
        const char* e = getenv("DIPB_MSA_TC");
        const bool never = e && e[0] == '0', always = e && e[0] == '2';
        if (!never && (always || (r1 - r0 >= 64 && (long long)(r1 - r0) * ncols >= (1ll << 22))))
            return msa_tc_block(m, type, r0, r1, ncols, d_out, ld);
    }
    if (!fast_ok(m, type)) return generic_block(m, type, r0, r1, ncols, d_out, ld, r0, 0);
    TileParams p{};
    p.planes = m->planes; p.nv = m->nv; p.nkc = m->nkc; p.kc0 = 0; p.kc1 = m->nkc; p.tri = 0; p.n = m->n;
    p.bi0 = r0 / MSA_TS; p.nbi = (r1 - 1) / MSA_TS - p.bi0 + 1; p.nbj = (ncols + MSA_TS - 1) / MSA_TS;
    p.out = d_out; p.ld = ld; p.row_lo = r0; p.row_hi = r1; p.col_hi = ncols; p.row_off = r0; p.dist_type = type;
    return launch_fast(m, p, (long long)p.nbi * p.nbj);
}

int msa_matrix(dipb_msa* m, int type, int row_begin, int row_end, double* d_out) {
    // lower-triangle tiles whose row block intersects [row_begin,row_end); mirrored
    if (row_begin < 0 || row_end > m->n || row_begin >= row_end) { set_error("msa_matrix: bad rows"); return DIPB_E_ARG; }
    if (msa_tc_supported(m, type)) {
        // tensor-core path (msa_tc.cu): 3.8x the popcount kernel at 30 000 x 30 000 (53 ms vs 201 ms), bit-identical
        // output; below ~4 M pairs its set-up costs more than it saves (2 000 x 10 000: 1.0 ms vs 0.4 ms).
        // DIPB_MSA_TC=0 never, =2 always.
        const char* e = getenv("DIPB_MSA_TC");
        const bool never = e && e[0] == '0', always = e && e[0] == '2';
        const long long pairs = ((long long)row_end * row_end - (long long)row_begin * row_begin) / 2;
        if (!never && (always || pairs >= (1ll << 22))) return msa_tc_matrix(m, type, row_begin, row_end, d_out);
    }
    if (!fast_ok(m, type)) return generic_block(m, type, row_begin, row_end, 0, d_out, (size_t)m->n, 0, 1);
    TileParams p{};
    p.planes = m->planes; p.nv = m->nv; p.nkc = m->nkc; p.kc0 = 0; p.kc1 = m->nkc; p.tri = 1; p.n = m->n;
    p.bi0 = row_begin / MSA_TS;
    int bi1 = (row_end - 1) / MSA_TS;
    p.nbi = bi1 - p.bi0 + 1;
    p.out = d_out; p.ld = (size_t)m->n; p.row_lo = row_begin; p.row_hi = row_end; p.col_hi = m->n; p.dist_type = type;
    long long tiles = (long long)(bi1 + 1) * (bi1 + 2) / 2 - (long long)p.bi0 * (p.bi0 + 1) / 2;
    return launch_fast(m, p, tiles);
}

int msa_counts_dev(dipb_msa* m, int i0, int i1, int j1, int* d_match, int* d_both, size_t ld) {
    int fns[1] = {0};
    int* cnt[18] = {nullptr};
    cnt[0] = d_match; cnt[1] = d_both;
    return generic_counts(m, fns, 1, i0, i1, j1, cnt, ld);
}

int msa_repack(dipb_msa* m, const uint64_t* d_in, int comp64) {
    dipb_ctx* c = m->ctx;
    long long total = (long long)m->nkc * MSA_KC * m->npad;
    int threads = 256;
    long long blocks = (total + threads - 1) / threads;
    msa_repack_kernel<<<(unsigned)blocks, threads, 0, c->stream>>>(d_in, m->n, m->seq_len, comp64, m->nkc, m->planes, m->nv, m->npad);
    DIPB_KERNEL_CHECK(c);
    return 0;
}

}  // namespace dipb
