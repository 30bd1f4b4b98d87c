// K1-TC: aligned-MSA all-pairs distances on the 5th-gen tensor cores (tcgen05, sm_100a).
//
// Same outputs as msa_tile_kernel<0,true> (bit-identical counts, same fp64 epilogue) for the p / JC
// models; replaces the POPC-bound SIMT counting (XU pipe 92 % busy, profiles/r1_ncu_summary.json)
// with two int8 GEMMs accumulated in TMEM:
//   every site of a sequence becomes a vector s in {-1,0,1}^3 (simplex corners: <s_a,s_b> = 3 if the
//   bases are equal, -1 if they differ, 0 if either is not ACGT) and a validity flag v in {0,1}:
//       D1(i,j) = sum_sites <s_i, s_j> = 4 * match - both        (K = 3L int8)
//       D2(i,j) = sum_sites  v_i * v_j  = both                   (K =  L int8)
//   => match = (D1 + D2) / 4,  useful = nv[i] + nv[j] - D2        (exact in s32)
// Structure (one CTA per SM, persistent over 128 x 256 tiles of the lower triangle):
//   warp 0   TMA producer: cp.async.bulk.tensor.2d (SWIZZLE_128B) of a 128 x 128 B A slab and a
//            256 x 128 B B slab per stage, 4-stage mbarrier ring
//   warp 1   MMA issuer: one lane issues tcgen05.mma.cta_group::1.kind::i8 (M128 N256 K32), four per
//            stage, D1 in TMEM columns [0,256), D2 in [256,512); tcgen05.commit frees the stage
//   warps 2-5 epilogue: tcgen05.ld 32x32b, fp64 p / JC in the reference's expression order, D[i][j] + mirror
// Operand formats (template parameter FMT, fixed per alignment at its first expansion):
//   2 (default)  e2m1 nibbles, two elements per byte in HBM / L2; the tensor maps are CU_TENSOR_MAP_DATA_TYPE_16U4_ALIGN16B, so
//                the TMA unpacks 16 nibbles into a 16-byte group of 8 data + 8 padding bytes (same shared-memory footprint,
//                swizzle, descriptors and K = 32 step); tcgen05.mma.kind::f8f6f4 with f32 accumulators -- exact: every product
//                is -1, 0 or 1 and every partial sum an integer < 2^24.  expect_tx counts the PACKED bytes.
//   0            int8, kind::i8, s32 accumulators (DIPB_TC_FMT=0; the multicast cluster experiments)
// See DESIGN.md 4.1 for the measurements (neither L2->SM nor DRAM bytes bound the kernel; the epilogue is not overlapped).
#include <cuda.h>
#include <algorithm>
#include <vector>
#include "common.cuh"
#include "msa.cuh"
#include "msa_pair.cuh"

namespace dipb {

constexpr int TC_M = 128, TC_N = 256, TC_KB = 128;          // tile rows, tile cols, K bytes per stage
constexpr int TC_STAGES = 4;
constexpr int TC_A_BYTES = TC_M * TC_KB, TC_B_BYTES = TC_N * TC_KB;
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;     // 48 KB
constexpr int TC_SMEM = TC_STAGES * TC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TC_THREADS = 192;

// ---- operand expansion: planes -> simplex int8 (S) and validity int8 (V), K-major rows ----------
// rows [row0, row0 + nrows) of the alignment into rows 0.. of S / V
__global__ void msa_tc_expand_kernel(const uint32_t* __restrict__ planes, int row0, int nrows, int nkc, int w32, int8_t* __restrict__ S,
                                     size_t ks, int8_t* __restrict__ V, size_t kv, int fmt) {
    const uint32_t pos = 0x01u, neg = 0xFFu;   // two's complement int8 (fmt 0)
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)nrows * w32) return;
    const int sd = (int)(gid / w32), w = (int)(gid % w32);
    const int s = row0 + sd;
    const int sb = s / MSA_TS, sl = s % MSA_TS, kc = w / MSA_KC, kk = w % MSA_KC;
    const size_t base = ((size_t)sb * nkc + kc) * MSA_SLAB_WORDS + (size_t)kk * MSA_TS + sl;
    const uint32_t b0 = planes[base], b1 = planes[base + MSA_KC * MSA_TS], v = planes[base + 2 * MSA_KC * MSA_TS];
    if (fmt == 2) {
        // e2m1 nibbles (+1 = 0x2, -1 = 0xA, 0 = 0x0), two elements per byte, element k in the low nibble of byte k / 2 for even k:
        // 48 bytes of S (three 16-byte runs of 32 elements) and 16 bytes of V per 32-site word
        uint8_t* so4 = reinterpret_cast<uint8_t*>(S) + (size_t)sd * ks + (size_t)w * 48;
        uint8_t* vo4 = reinterpret_cast<uint8_t*>(V) + (size_t)sd * kv + (size_t)w * 16;
        uint32_t e0[4], e1[4], e2[4], vv[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            e0[q] = e1[q] = e2[q] = vv[q] = 0u;
#pragma unroll
            for (int t = 0; t < 8; t++) {
                const int bit = 8 * q + t;
                const int ok = (v >> bit) & 1, x0 = (b0 >> bit) & 1, x1 = (b1 >> bit) & 1;
                const uint32_t n0 = ok ? (x1 ? 0xAu : 0x2u) : 0u, n1 = ok ? (x0 ? 0xAu : 0x2u) : 0u, n2 = ok ? ((x0 ^ x1) ? 0xAu : 0x2u) : 0u;
                e0[q] |= n0 << (4 * t); e1[q] |= n1 << (4 * t); e2[q] |= n2 << (4 * t);
                vv[q] |= (ok ? 0x2u : 0u) << (4 * t);
            }
        }
        // (16-byte stores: a thread's 48 + 16 bytes leave in 4 instructions instead of 16)
        reinterpret_cast<uint4*>(so4)[0] = make_uint4(e0[0], e0[1], e0[2], e0[3]);
        reinterpret_cast<uint4*>(so4)[1] = make_uint4(e1[0], e1[1], e1[2], e1[3]);
        reinterpret_cast<uint4*>(so4)[2] = make_uint4(e2[0], e2[1], e2[2], e2[3]);
        reinterpret_cast<uint4*>(vo4)[0] = make_uint4(vv[0], vv[1], vv[2], vv[3]);
        return;
    }
    int8_t* so = S + (size_t)sd * ks + (size_t)w * 96;
    int8_t* vo = V + (size_t)sd * kv + (size_t)w * 32;
#pragma unroll 4
    for (int q = 0; q < 8; q++) {
        uint32_t e0 = 0, e1 = 0, e2 = 0, vv = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int bit = 4 * q + t;
            const int ok = (v >> bit) & 1, x0 = (b0 >> bit) & 1, x1 = (b1 >> bit) & 1;
            const int c0 = ok ? 1 - 2 * x1 : 0, c1 = ok ? 1 - 2 * x0 : 0, c2 = ok ? 1 - 2 * (x0 ^ x1) : 0;
            e0 |= (c0 == 0 ? 0u : (c0 > 0 ? pos : neg)) << (8 * t);
            e1 |= (c1 == 0 ? 0u : (c1 > 0 ? pos : neg)) << (8 * t);
            e2 |= (c2 == 0 ? 0u : (c2 > 0 ? pos : neg)) << (8 * t);
            vv |= (ok ? pos : 0u) << (8 * t);
        }
        reinterpret_cast<uint32_t*>(so)[q] = e0;
        reinterpret_cast<uint32_t*>(so + 32)[q] = e1;
        reinterpret_cast<uint32_t*>(so + 64)[q] = e2;
        reinterpret_cast<uint32_t*>(vo)[q] = vv;
    }
}

// ---- PTX helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
    // K-major, SWIZZLE_128B: 8-row groups are 1024 B apart (SBO), LBO unused, descriptor version 1 (sm_100)
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// multicast forms: the TMA box lands at the same shared-memory offset in every CTA of `mask` and completes bytes on
// the barrier at the same offset there; the commit arrives on that barrier in every CTA of `mask`
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}

struct TcParams {
    const int2* tiles;      // (row block of 128, col block of 256)
    int num_tiles;
    int ns_chunks, nv_chunks;   // 128-byte K chunks of the S and V operands
    const int* nv;
    int n;
    double* out;
    size_t ld;
    int dist_type;
    int rect;                   // 0: lower triangle + mirror into an n x n matrix; 1: rows [r0,r1) x cols [0,ncols)
    int r0, r1, ncols;
    int a_row0;                 // the A operand buffers start at this alignment row (0: persistent buffers, else scratch)
};

// Epilogue of one 128 x 256 tile: thread (q, lane) owns row 32 q + lane of the tile; D1 in TMEM columns [0,256), D2 in
// [256,512).  match = (D1 + D2) / 4, useful = nv_i + nv_j - D2, then p / JC in fp64 in the reference's expression order.
template <int FMT = 0>
__device__ __forceinline__ void tc_epilogue_tile(const TcParams& p, uint32_t tmem_base, int tile_row, int tile_col, int q, int lane) {
    // FMT 0: s32 accumulators; else f32 accumulators holding exact integers
    auto acc = [](uint32_t r) -> int { return FMT == 0 ? (int)r : __float2int_rn(__uint_as_float(r)); };
    const int2 tl = make_int2(tile_row, tile_col);
    const int i = tl.x * TC_M + 32 * q + lane;
    const int nvi = i < p.n ? p.nv[i] : 0;
    for (int cb = 0; cb < TC_N / 32; cb++) {
        uint32_t r1[32], r2[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(cb * 32);
        tmem_ld32(taddr, r1);
        tmem_ld32(taddr + 256u, r2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int j0 = tl.y * TC_N + cb * 32;
        if (p.rect) {
            if (i >= p.r0 && i < p.r1 && j0 < p.ncols) {
#pragma unroll
                for (int c = 0; c < 32; c++) {
                    const int j = j0 + c;
                    if (j < p.ncols) {
                        const int both = acc(r2[c]);
                        const int match = (acc(r1[c]) + both) >> 2;
                        p.out[(size_t)(i - p.r0) * p.ld + j] = dist_p_jc(match, nvi + p.nv[j] - both, p.dist_type);
                    }
                }
            }
        } else if (i >= p.r0 && i < p.r1 && j0 <= i) {
#pragma unroll
            for (int c = 0; c < 32; c++) {
                const int j = j0 + c;
                if (j < i) {
                    const int both = acc(r2[c]);
                    const int match = (acc(r1[c]) + both) >> 2;
                    const double d = dist_p_jc(match, nvi + p.nv[j] - both, p.dist_type);
                    p.out[(size_t)i * p.ld + j] = d;
                    p.out[(size_t)j * p.ld + i] = d;
                } else if (j == i) {
                    p.out[(size_t)i * p.ld + i] = 0.0;
                }
            }
        }
    }
}

// CM x CN CTAs form a cluster that computes a (CM * 128) x (CN * 256) tile: the A slab of a tile row is needed by the CN
// CTAs of that row and the B slab of a tile column by the CM CTAs of that column, so every CTA loads 1/CN of its A
// slab and 1/CM of its B slab and TMA-multicasts them to the peers: (1/CN + 2/CM) / 3 of the L2 -> SM traffic of
// independent CTAs, and sharing no longer depends on L2 residency (ncu, 1 x 1: 248 GB of DRAM reads for 3.6 GB of operands).
template <int CM, int CN, int FMT = 0>
__global__ void __launch_bounds__(TC_THREADS, 1)
msa_tc_kernel(const __grid_constant__ CUtensorMap mapSA, const __grid_constant__ CUtensorMap mapSB,
              const __grid_constant__ CUtensorMap mapVA, const __grid_constant__ CUtensorMap mapVB, TcParams p) {
    constexpr int CSZ = CM * CN;
    uint32_t crank = 0;
    if (CSZ > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const int cr = (int)crank / CN, cc = (int)crank % CN;
    const int cluster_id = (int)blockIdx.x / CSZ, num_clusters = (int)gridDim.x / CSZ;
    const uint16_t mask_row = (uint16_t)(((1u << CN) - 1u) << (cr * CN));      // CTAs that need my part of the A slab
    uint16_t mask_col = 0;                                                      // CTAs that need my part of the B slab
#pragma unroll
    for (int r = 0; r < CM; r++) mask_col |= (uint16_t)(1u << (r * CN + cc));
    constexpr int A_PART = TC_A_BYTES / CN, B_PART = TC_B_BYTES / CM;
    extern __shared__ unsigned char smem_raw[];
    // SWIZZLE_128B operands need 1024-byte alignment
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES);
    uint64_t* full = bars;                    // [TC_STAGES]
    uint64_t* empty = bars + TC_STAGES;       // [TC_STAGES]
    uint64_t* tmem_full = bars + 2 * TC_STAGES;
    uint64_t* tmem_empty = bars + 2 * TC_STAGES + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        // a stage is free when every CTA that reads what this CTA writes into it has consumed it
        for (int s = 0; s < TC_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], CM + CN - 1); }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 4);   // one arrival per epilogue warp
        mbar_fence_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CSZ > 1) {   // barriers of every CTA initialised before any peer multicasts into them
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int nchunks = p.ns_chunks + p.nv_chunks;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int t = cluster_id; t < p.num_tiles; t += num_clusters) {
                const int2 tl = p.tiles[t];
                const int arow = (tl.x * CM + cr) * TC_M + cc * (TC_M / CN) - p.a_row0;   // my part of my tile row's A slab
                const int brow = (tl.y * CN + cc) * TC_N + cr * (TC_N / CM);   // my part of my tile column's B slab
                for (int c = 0; c < nchunks; c++) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char* a = smem + (size_t)stage * TC_STAGE_BYTES + cc * A_PART;
                    unsigned char* b = smem + (size_t)stage * TC_STAGE_BYTES + TC_A_BYTES + cr * B_PART;
                    mbar_arrive_expect_tx(&full[stage], FMT == 2 ? TC_STAGE_BYTES / 2 : TC_STAGE_BYTES);   // own parts + the peers' multicasts (e2m1: packed bytes)
                    const bool sv = c >= p.ns_chunks;
                    const int kc = (sv ? c - p.ns_chunks : c) * TC_KB;
                    if (CSZ == 1) {
                        tma_load_2d(a, sv ? &mapVA : &mapSA, kc, arow, &full[stage]);
                        tma_load_2d(b, sv ? &mapVB : &mapSB, kc, brow, &full[stage]);
                    } else {
                        tma_load_2d_mc(a, sv ? &mapVA : &mapSA, kc, arow, &full[stage], mask_row);
                        tma_load_2d_mc(b, sv ? &mapVB : &mapSB, kc, brow, &full[stage], mask_col);
                    }
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            // instruction descriptor: D = s32, A = B = signed 8-bit, both K-major, N = 256, M = 128
            const uint32_t idesc = (FMT == 0 ? ((2u << 4) | (1u << 7) | (1u << 10)) : ((1u << 4) | (5u << 7) | (5u << 10))) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
            uint32_t stage = 0, phase = 0, tphase = 0;
            for (int t = cluster_id; t < p.num_tiles; t += num_clusters) {
                mbar_wait(tmem_empty, tphase ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int c = 0; c < nchunks; c++) {
                    mbar_wait(&full[stage], phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = smem_u32(smem + (size_t)stage * TC_STAGE_BYTES);
                    const uint64_t adesc = make_sw128_desc(a_addr), bdesc = make_sw128_desc(a_addr + TC_A_BYTES);
                    const bool second = c >= p.ns_chunks;
                    const uint32_t d = tmem_base + (second ? 256u : 0u);
                    const bool first_of_acc = (c == 0) || (c == p.ns_chunks);
#pragma unroll
                    for (int k = 0; k < TC_KB / 32; k++)   // UMMA_K = 32 int8 = 32 B: advance the start address by 2 (x16 B)
                        if (FMT == 0) umma_i8(d, adesc + 2 * k, bdesc + 2 * k, idesc, (first_of_acc && k == 0) ? 0u : 1u);
                        else umma_f8(d, adesc + 2 * k, bdesc + 2 * k, idesc, (first_of_acc && k == 0) ? 0u : 1u);
                    // the stage is free when these MMAs have read it: tell every CTA that writes into it
                    if (CSZ == 1) umma_commit(&empty[stage]);
                    else umma_commit_mc(&empty[stage], (uint16_t)(mask_row | mask_col));
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(tmem_full);                    // both accumulators complete
                tphase ^= 1;
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;                            // TMEM lane quarter this warp may read
        uint32_t tphase = 0;
        for (int t = cluster_id; t < p.num_tiles; t += num_clusters) {
            const int2 tl = make_int2(p.tiles[t].x * CM + cr, p.tiles[t].y * CN + cc);   // this CTA's 128 x 256 tile
            mbar_wait(tmem_full, tphase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            tc_epilogue_tile<FMT>(p, tmem_base, tl.x, tl.y, q, lane);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty);
            tphase ^= 1;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CSZ > 1) {   // no CTA exits while a peer may still multicast into it or arrive on its barriers
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// =====================================================================================================
// 2-CTA variant (tcgen05 cta_group::2).  ncu on the 1-CTA kernel: the tensor pipe is ~50 % active because every
// SM has to take in 48 KB of operands per 128-byte K step (~1060 cycles) against 541 cycles of MMA.  A CTA pair
// computes a 256 x 256 tile with ONE MMA stream: each CTA stages only its own 128 A rows and its own half of the
// 256 B rows (32 KB per step, 7 stages instead of 4); the leader's MMA reads both halves from both shared memories
// and writes each CTA's 128 accumulator rows into that CTA's TMEM.  Both CTAs run a TMA producer (completing bytes on
// the LEADER's full barrier) and an epilogue; only the leader issues MMAs and multicasts its commits to both CTAs.
// =====================================================================================================
constexpr int T2_STAGES = 7;
constexpr int T2_HALF_BYTES = 128 * TC_KB;                  // 128 rows x 128 B
constexpr int T2_STAGE_BYTES = 2 * T2_HALF_BYTES;           // A rows + half of B: 32 KB
constexpr int T2_SMEM = T2_STAGES * T2_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t cluster_addr(const void* p, uint32_t rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(p)), "r"(rank));
    return ra;
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t bar_cluster_addr, uint32_t bytes) {
    // relaxed: the bytes are ordered by the TMA completion itself; a release at cluster scope is a MEMBAR.GPU (~1000 cycles,
    // once per stage in the producer: measured 1970 instead of ~700 cycles per stage)
    asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(bar_cluster_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
// TMA load into this CTA's shared memory that completes its bytes on a barrier of the pair's leader
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster_addr) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void umma2_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma2_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {   // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

template <int FMT>
__global__ void __launch_bounds__(TC_THREADS, 1)
msa_tc2_kernel(const __grid_constant__ CUtensorMap mapSA, const __grid_constant__ CUtensorMap mapSB,
               const __grid_constant__ CUtensorMap mapVA, const __grid_constant__ CUtensorMap mapVB, TcParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T2_STAGES * T2_STAGE_BYTES);
    uint64_t* full = bars;                      // [T2_STAGES]  used in the leader: both producers arm it and complete bytes on it
    uint64_t* empty = bars + T2_STAGES;         // [T2_STAGES]  in each CTA: the leader's commit frees the stage in both
    uint64_t* tmem_full = bars + 2 * T2_STAGES;       // in each CTA: accumulators of the tile complete
    uint64_t* tmem_empty = bars + 2 * T2_STAGES + 1;  // in the leader: 4 + 4 epilogue warps of the pair have drained TMEM
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * T2_STAGES + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t crank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const int pair_id = (int)blockIdx.x >> 1, num_pairs = (int)gridDim.x >> 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < T2_STAGES; s++) { mbar_init(&full[s], 2); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 8);
        mbar_fence_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int nchunks = p.ns_chunks + p.nv_chunks;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int t = pair_id; t < p.num_tiles; t += num_pairs) {
                const int2 tl = p.tiles[t];                                  // 256-row block, 256-column block
                const int arow = tl.x * 256 + (int)crank * 128 - p.a_row0;  // my 128 A rows
                const int brow = tl.y * 256 + (int)crank * 128;              // my half of the B rows
                for (int c = 0; c < nchunks; c++) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    unsigned char* a = smem + (size_t)stage * T2_STAGE_BYTES;
                    const uint32_t lfull = cluster_addr(&full[stage], 0);
                    // (e2m1: the transaction counts the packed bytes that leave global memory, half of what lands in shared memory)
                    mbar_arrive_expect_tx_cluster(lfull, FMT == 2 ? T2_STAGE_BYTES / 2 : T2_STAGE_BYTES);
                    const bool sv = c >= p.ns_chunks;
                    const int kc = (sv ? c - p.ns_chunks : c) * TC_KB;
                    tma_load_2d_pair(a, sv ? &mapVA : &mapSA, kc, arow, lfull);
                    tma_load_2d_pair(a + T2_HALF_BYTES, sv ? &mapVB : &mapSB, kc, brow, lfull);
                    if (++stage == T2_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader only) =====================
        if (lane == 0 && crank == 0) {
            // D = s32, A = B = signed 8-bit, both K-major, N = 256, M = 256 (128 rows in each CTA's TMEM)
            // (FMT 2: kind::f8f6f4, D = f32 (c_format 1), A = B = e2m1 (a/b_format 5): 4-bit in global memory, unpacked by the TMA
            //  into 16-byte groups of 8 data + 8 padding bytes -- the same shared-memory footprint and K step as int8)
            const uint32_t idesc = (FMT == 0 ? ((2u << 4) | (1u << 7) | (1u << 10)) : ((1u << 4) | (5u << 7) | (5u << 10))) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            uint32_t stage = 0, phase = 0, tphase = 0;
            for (int t = pair_id; t < p.num_tiles; t += num_pairs) {
                mbar_wait(tmem_empty, tphase ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int c = 0; c < nchunks; c++) {
                    mbar_wait(&full[stage], phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = smem_u32(smem + (size_t)stage * T2_STAGE_BYTES);
                    const uint64_t adesc = make_sw128_desc(a_addr), bdesc = make_sw128_desc(a_addr + T2_HALF_BYTES);
                    const bool second = c >= p.ns_chunks;
                    const uint32_t d = tmem_base + (second ? 256u : 0u);
                    const bool first_of_acc = (c == 0) || (c == p.ns_chunks);
#pragma unroll
                    for (int k = 0; k < TC_KB / 32; k++)
                        if (FMT == 0) umma2_i8(d, adesc + 2 * k, bdesc + 2 * k, idesc, (first_of_acc && k == 0) ? 0u : 1u);
                        else umma2_f8(d, adesc + 2 * k, bdesc + 2 * k, idesc, (first_of_acc && k == 0) ? 0u : 1u);
                    umma2_commit(&empty[stage]);
                    if (++stage == T2_STAGES) { stage = 0; phase ^= 1; }
                }
                umma2_commit(tmem_full);
                tphase ^= 1;
            }
        }
    } else {
        // ===================== epilogue (warps 2..5, both CTAs) =====================
        const int q = warp & 3;
        uint32_t tphase = 0;
        const uint32_t l_tmem_empty = cluster_addr(tmem_empty, 0);
        for (int t = pair_id; t < p.num_tiles; t += num_pairs) {
            const int2 tl = p.tiles[t];
            mbar_wait(tmem_full, tphase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            tc_epilogue_tile<FMT>(p, tmem_base, tl.x * 2 + (int)crank, tl.y, q, lane);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(l_tmem_empty);
            tphase ^= 1;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(EncodeTiledFn fn, CUtensorMap* m, void* base, size_t kbytes, size_t rows, int box_rows, int fmt = 0) {
    // fmt 2: rows of packed 4-bit elements (kbytes * 2 of them); a box is still 128 elements = one 128-byte swizzle row of
    // shared memory after the TMA's unpack (CU_TENSOR_MAP_DATA_TYPE_16U4_ALIGN16B)
    cuuint64_t dims[2] = {(cuuint64_t)(fmt == 2 ? kbytes * 2 : kbytes), (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kbytes};
    cuuint32_t box[2] = {(cuuint32_t)TC_KB, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(m, fmt == 2 ? CU_TENSOR_MAP_DATA_TYPE_16U4_ALIGN16B : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return DIPB_E_CUDA; }
    return 0;
}

bool msa_tc_supported(const dipb_msa* m, int type) {
    return (type == DIPB_DIST_UNCORRECTED || type == DIPB_DIST_JC) && m->n >= 2;
}

void msa_tc_reserve(dipb_msa* m, int rows) {
    if (!m->tc_S && rows > 0) m->tc_reserve = (size_t)rows;
}

static int tc_expand(dipb_msa* m, size_t row0, size_t nrows, int8_t* S, int8_t* V) {
    if (!nrows) return 0;
    dipb_ctx* c = m->ctx;
    const int w32 = (m->seq_len + 31) / 32;
    const long long total = (long long)nrows * w32;
    msa_tc_expand_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(m->planes, (int)row0, (int)nrows, m->nkc, w32, S, m->tc_ks, V, m->tc_kv, m->tc_fmt);
    DIPB_KERNEL_CHECK(c);
    return 0;
}

// persistent buffers hold at least the rows [0, rows)
static int tc_ensure_prefix(dipb_msa* m, size_t rows) {
    dipb_ctx* c = m->ctx;
    if (rows > (size_t)m->n) rows = (size_t)m->n;
    if (!m->tc_ks) {
        // operand format: e2m1 (4 bits per element in HBM / L2, kind::f8f6f4, f32 accumulators: exact for these +-1 / 0 products,
        // sums < 2^24) unless DIPB_TC_FMT=0 asks for int8 or one of the int8-only multicast cluster experiments is selected
        const char* ef = getenv("DIPB_TC_FMT");
        m->tc_fmt = (ef ? atoi(ef) == 2 : true) && !getenv("DIPB_MSA_TC_CLUSTER") ? 2 : 0;
        const int w32 = (m->seq_len + 31) / 32;
        m->tc_ks = ((size_t)w32 * 96 + TC_KB - 1) / TC_KB * TC_KB;     // K elements, padded to whole 128-element chunks ...
        m->tc_kv = ((size_t)w32 * 32 + TC_KB - 1) / TC_KB * TC_KB;
        if (m->tc_fmt == 2) { m->tc_ks /= 2; m->tc_kv /= 2; }            // ... = bytes per row, two elements per byte for e2m1
    }
    const size_t full = ((size_t)m->n + 255) / 256 * 256;
    if (rows > m->tc_rows) {
        // (re)allocate: the reserved size when it is enough, else everything
        size_t want = m->tc_reserve ? (m->tc_reserve + 255) / 256 * 256 : full;
        if (want < rows || want > full) want = full;
        int8_t *S = nullptr, *V = nullptr;
        DIPB_CUDA(pool_alloc(c, (void**)&S, want * m->tc_ks));
        DIPB_CUDA(pool_alloc(c, (void**)&V, want * m->tc_kv));
        DIPB_CUDA(cudaMemsetAsync(S, 0, want * m->tc_ks, c->stream));   // K padding and rows never expanded read as 0
        DIPB_CUDA(cudaMemsetAsync(V, 0, want * m->tc_kv, c->stream));
        if (m->tc_have) {
            DIPB_CUDA(cudaMemcpyAsync(S, m->tc_S, m->tc_have * m->tc_ks, cudaMemcpyDeviceToDevice, c->stream));
            DIPB_CUDA(cudaMemcpyAsync(V, m->tc_V, m->tc_have * m->tc_kv, cudaMemcpyDeviceToDevice, c->stream));
        }
        pool_free(c, m->tc_S); pool_free(c, m->tc_V);
        m->tc_S = S; m->tc_V = V; m->tc_rows = want;
    }
    if (rows > m->tc_have) {
        int rc = tc_expand(m, m->tc_have, rows - m->tc_have, m->tc_S + m->tc_have * m->tc_ks, m->tc_V + m->tc_have * m->tc_kv);
        if (rc) return rc;
        m->tc_have = rows;
    }
    return 0;
}

// operands of one launch: B rows (columns of the result) always from the persistent prefix; A rows from it too when
// they fit its capacity, else from a scratch pair that holds just the 256-aligned row range of this call
struct TcOperands {
    int8_t *SA, *VA, *SB, *VB;
    size_t rowsA, rowsB;
    int a_row0;
};
static int tc_operands(dipb_msa* m, int r0, int r1, int ncols, TcOperands* o) {
    dipb_ctx* c = m->ctx;
    int rc = tc_ensure_prefix(m, (size_t)ncols);
    if (rc) return rc;
    const size_t cap_target = m->tc_reserve ? std::min<size_t>(((size_t)m->n + 255) / 256 * 256, (m->tc_reserve + 255) / 256 * 256) : ((size_t)m->n + 255) / 256 * 256;
    if ((size_t)r1 <= m->tc_rows || (size_t)r1 <= cap_target) {
        if ((rc = tc_ensure_prefix(m, (size_t)r1))) return rc;
        o->SA = m->tc_S; o->VA = m->tc_V; o->rowsA = m->tc_rows; o->a_row0 = 0;
    } else {
        const size_t a0 = (size_t)r0 / 256 * 256, a1 = std::min<size_t>(((size_t)r1 + 255) / 256 * 256, (size_t)m->n);
        const size_t need = ((size_t)r1 + 255) / 256 * 256 - a0;
        if (need > m->tc_xrows) {
            pool_free(c, m->tc_Sx); pool_free(c, m->tc_Vx);
            m->tc_Sx = m->tc_Vx = nullptr;
            DIPB_CUDA(pool_alloc(c, (void**)&m->tc_Sx, need * m->tc_ks));
            DIPB_CUDA(pool_alloc(c, (void**)&m->tc_Vx, need * m->tc_kv));
            DIPB_CUDA(cudaMemsetAsync(m->tc_Sx, 0, need * m->tc_ks, c->stream));
            DIPB_CUDA(cudaMemsetAsync(m->tc_Vx, 0, need * m->tc_kv, c->stream));
            m->tc_xrows = need;
        }
        if ((rc = tc_expand(m, a0, a1 - a0, m->tc_Sx, m->tc_Vx))) return rc;
        o->SA = m->tc_Sx; o->VA = m->tc_Vx; o->rowsA = m->tc_xrows; o->a_row0 = (int)a0;
    }
    o->SB = m->tc_S; o->VB = m->tc_V; o->rowsB = m->tc_rows;
    return 0;
}

typedef std::vector<int2> (*TileListFn)(int tile_m, int tile_n, int sm_rows, int sm_cols, const void* arg);

template <int CM, int CN>
static int tc_launch_c(dipb_msa* m, const TcOperands& op, TileListFn make_tiles, const void* arg, TcParams p, bool* launched) {
    dipb_ctx* c = m->ctx;
    *launched = false;
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        DIPB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled not available"); return DIPB_E_CUDA; }
        encode = (EncodeTiledFn)fn;
    }
    // (e2m1 operands: the plain 1 x 1 kernel only; the multicast cluster experiments keep int8, see tc_ensure_prefix)
    const void* kern = (CM == 1 && CN == 1 && m->tc_fmt == 2) ? (const void*)msa_tc_kernel<1, 1, 2> : (const void*)msa_tc_kernel<CM, CN, 0>;
    const int fmt = m->tc_fmt;
    constexpr int CSZ = CM * CN;
    DIPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    if (CSZ > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CSZ); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = TC_SMEM; cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CSZ; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = c->num_sms;
    if (CSZ > 1 && (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) != cudaSuccess || nclusters < 1)) { cudaGetLastError(); return 0; }
    // the clusters that run concurrently form one super-tile (rows x cols of cluster tiles) that shares operand slabs in L2
    int sm_cols = 1;
    while ((sm_cols + 1) * (sm_cols + 1) * CM * 2 <= nclusters * CN) sm_cols++;   // keep the super-tile about square in elements
    int sm_rows = nclusters / sm_cols;
    if (sm_rows < 1) sm_rows = 1;
    std::vector<int2> tiles = make_tiles(CM * TC_M, CN * TC_N, sm_rows, sm_cols, arg);
    if (tiles.empty()) { *launched = true; return 0; }
    int rc;
    CUtensorMap mSA, mSB, mVA, mVB;
    if ((rc = make_map(encode, &mSA, op.SA, m->tc_ks, op.rowsA, TC_M / CN, fmt)) || (rc = make_map(encode, &mSB, op.SB, m->tc_ks, op.rowsB, TC_N / CM, fmt)) ||
        (rc = make_map(encode, &mVA, op.VA, m->tc_kv, op.rowsA, TC_M / CN, fmt)) || (rc = make_map(encode, &mVB, op.VB, m->tc_kv, op.rowsB, TC_N / CM, fmt)))
        return rc;
    int2* d_tiles = nullptr;
    DIPB_CUDA(pool_alloc(c, (void**)&d_tiles, sizeof(int2) * tiles.size()));
    DIPB_CUDA(cudaMemcpyAsync(d_tiles, tiles.data(), sizeof(int2) * tiles.size(), cudaMemcpyHostToDevice, c->stream));
    p.tiles = d_tiles; p.num_tiles = (int)tiles.size();
    p.ns_chunks = (int)(m->tc_ks * (fmt == 2 ? 2 : 1) / TC_KB); p.nv_chunks = (int)(m->tc_kv * (fmt == 2 ? 2 : 1) / TC_KB);
    p.nv = m->nv; p.n = m->n;
    const int use = p.num_tiles < nclusters ? p.num_tiles : nclusters;
    cfg.gridDim = dim3(use * CSZ);
    void* args[] = {&mSA, &mSB, &mVA, &mVB, &p};
    cudaError_t e = cudaLaunchKernelExC(&cfg, kern, args);
    if (e != cudaSuccess) { pool_free(c, d_tiles); set_error("msa_tc: launch failed: %s", cudaGetErrorString(e)); return DIPB_E_CUDA; }
    c->launches++;
    DIPB_CUDA(cudaStreamSynchronize(c->stream));
    pool_free(c, d_tiles);
    *launched = true;
    return 0;
}

// 2-CTA kernel: 256 x 256 tiles, one CTA pair (cluster of 2) each
static int tc_launch_pair(dipb_msa* m, const TcOperands& op, TileListFn make_tiles, const void* arg, TcParams p, bool force, bool* launched) {
    dipb_ctx* c = m->ctx;
    *launched = false;
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        DIPB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled not available"); return DIPB_E_CUDA; }
        encode = (EncodeTiledFn)fn;
    }
    const void* kern2 = m->tc_fmt == 2 ? (const void*)msa_tc2_kernel<2> : (const void*)msa_tc2_kernel<0>;
    if (cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = T2_SMEM; cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int npairs = 0;
    if (cudaOccupancyMaxActiveClusters(&npairs, kern2, &cfg) != cudaSuccess || npairs < 1) { cudaGetLastError(); return 0; }
    int sm_cols = 1;
    while ((sm_cols + 1) * (sm_cols + 1) <= npairs) sm_cols++;      // square tiles: square super-tile
    const int sm_rows = npairs / sm_cols;
    std::vector<int2> tiles = make_tiles(256, 256, sm_rows, sm_cols, arg);
    if (tiles.empty()) { *launched = true; return 0; }
    if (!force && (int)tiles.size() < 4 * npairs) return 0;      // too few 256 x 256 tiles to balance: 128 x 256 tiles of the 1-CTA kernel
    int rc;
    CUtensorMap mSA, mSB, mVA, mVB;
    const int fmt = m->tc_fmt;
    if ((rc = make_map(encode, &mSA, op.SA, m->tc_ks, op.rowsA, 128, fmt)) || (rc = make_map(encode, &mSB, op.SB, m->tc_ks, op.rowsB, 128, fmt)) ||
        (rc = make_map(encode, &mVA, op.VA, m->tc_kv, op.rowsA, 128, fmt)) || (rc = make_map(encode, &mVB, op.VB, m->tc_kv, op.rowsB, 128, fmt)))
        return rc;
    int2* d_tiles = nullptr;
    DIPB_CUDA(pool_alloc(c, (void**)&d_tiles, sizeof(int2) * tiles.size()));
    DIPB_CUDA(cudaMemcpyAsync(d_tiles, tiles.data(), sizeof(int2) * tiles.size(), cudaMemcpyHostToDevice, c->stream));
    p.tiles = d_tiles; p.num_tiles = (int)tiles.size();
    p.ns_chunks = (int)(m->tc_ks * (fmt == 2 ? 2 : 1) / TC_KB); p.nv_chunks = (int)(m->tc_kv * (fmt == 2 ? 2 : 1) / TC_KB);
    p.nv = m->nv; p.n = m->n;
    const int use = p.num_tiles < npairs ? p.num_tiles : npairs;
    cfg.gridDim = dim3(use * 2);
    void* args[] = {&mSA, &mSB, &mVA, &mVB, &p};
    cudaError_t e = cudaLaunchKernelExC(&cfg, kern2, args);
    if (e != cudaSuccess) { pool_free(c, d_tiles); set_error("msa_tc2: launch failed: %s", cudaGetErrorString(e)); return DIPB_E_CUDA; }
    c->launches++;
    DIPB_CUDA(cudaStreamSynchronize(c->stream));
    pool_free(c, d_tiles);
    *launched = true;
    return 0;
}

static int tc_launch(dipb_msa* m, int r0, int r1, int ncols, TileListFn make_tiles, const void* arg, TcParams p) {
    TcOperands op;
    int rc = tc_operands(m, r0, r1, ncols, &op);
    if (rc) return rc;
    p.a_row0 = op.a_row0;
    {
        // 2-CTA kernel (39.9 ms vs 56.7 ms at C3) whenever there are enough 256 x 256 tiles; DIPB_MSA_TC2=0 never, =1 always
        const char* e2 = getenv("DIPB_MSA_TC2");
        if (!(e2 && e2[0] == '0')) {
            bool ok2 = false;
            rc = tc_launch_pair(m, op, make_tiles, arg, p, e2 && e2[0] == '1', &ok2);
            if (rc || ok2) return rc;
        }
    }
    // cluster shape: DIPB_MSA_TC_CLUSTER=RxC (1x1, 2x2, 2x4, 4x2, 4x4).  Multicast clusters cut the DRAM traffic (4x4: 87 GB
    // instead of 248 GB at C3) but not the run time -- the kernel is bound by each SM's 48 KB per stage of operand
    // ingest, not by L2 or DRAM (profiles/r1_ncu_tc_multicast.json) -- so independent CTAs on all 148 SMs stay the default.
    const char* e = getenv("DIPB_MSA_TC_CLUSTER");
    int cm = 1, cn = 1;
    if (m->tc_fmt == 0 && e && e[0] >= '1' && e[0] <= '4' && e[1] == 'x' && e[2] >= '1' && e[2] <= '4') { cm = e[0] - '0'; cn = e[2] - '0'; }   // (int8 operands only)
    bool ok = false;
    if (cm == 4 && cn == 4) rc = tc_launch_c<4, 4>(m, op, make_tiles, arg, p, &ok);
    else if (cm == 2 && cn == 4) rc = tc_launch_c<2, 4>(m, op, make_tiles, arg, p, &ok);
    else if (cm == 4 && cn == 2) rc = tc_launch_c<4, 2>(m, op, make_tiles, arg, p, &ok);
    else if (cm == 2 && cn == 2) rc = tc_launch_c<2, 2>(m, op, make_tiles, arg, p, &ok);
    if (rc || ok) return rc;
    return tc_launch_c<1, 1>(m, op, make_tiles, arg, p, &ok);
}

struct TriArg { int row_begin, row_end; };
// lower-triangle tiles of rows [row_begin, row_end), super-tile by super-tile from the bottom (longest rows first)
static std::vector<int2> tri_tiles(int tm, int tn, int sm_rows, int sm_cols, const void* arg) {
    const TriArg* a = static_cast<const TriArg*>(arg);
    const int mb0 = a->row_begin / tm, mb = (a->row_end + tm - 1) / tm;
    std::vector<int2> tiles;
    for (int sm = (mb + sm_rows - 1) / sm_rows - 1; sm >= mb0 / sm_rows; sm--) {
        const int mi_hi = std::min(mb, (sm + 1) * sm_rows) - 1;
        const int nj_max = (mi_hi * tm + tm - 1) / tn;
        for (int sn = 0; sn * sm_cols <= nj_max; sn++)
            for (int mi = std::max(mb0, sm * sm_rows); mi <= mi_hi; mi++)
                for (int nj = sn * sm_cols; nj < (sn + 1) * sm_cols; nj++)
                    if ((long long)nj * tn <= (long long)mi * tm + tm - 1) tiles.push_back(make_int2(mi, nj));   // touches the triangle
    }
    return tiles;
}

struct RectArg { int r0, r1, ncols; };
static std::vector<int2> rect_tiles(int tm, int tn, int sm_rows, int sm_cols, const void* arg) {
    const RectArg* a = static_cast<const RectArg*>(arg);
    const int mi0 = a->r0 / tm, mi1 = (a->r1 - 1) / tm, nb = (a->ncols + tn - 1) / tn;
    std::vector<int2> tiles;
    for (int sm = mi0; sm <= mi1; sm += sm_rows)
        for (int sn = 0; sn < nb; sn += sm_cols)
            for (int mi = sm; mi <= std::min(mi1, sm + sm_rows - 1); mi++)
                for (int nj = sn; nj < std::min(nb, sn + sm_cols); nj++) tiles.push_back(make_int2(mi, nj));
    return tiles;
}

// symmetric matrix entries (i, j) and (j, i) for i in [row_begin, row_end), j < i, plus the diagonal; p or JC
int msa_tc_matrix(dipb_msa* m, int type, int row_begin, int row_end, double* d_out) {
    TriArg a{row_begin, row_end};
    TcParams p{};
    p.out = d_out; p.ld = (size_t)m->n; p.dist_type = type; p.rect = 0; p.r0 = row_begin; p.r1 = row_end;
    return tc_launch(m, row_begin, row_end, row_end, tri_tiles, &a, p);
}

// rows [r0, r1) x columns [0, ncols) into out[(i - r0) * ld + j]  (placement row blocks, D&C stage 2)
int msa_tc_block(dipb_msa* m, int type, int r0, int r1, int ncols, double* d_out, size_t ld) {
    RectArg a{r0, r1, ncols};
    TcParams p{};
    p.out = d_out; p.ld = ld; p.dist_type = type; p.rect = 1; p.r0 = r0; p.r1 = r1; p.ncols = ncols;
    return tc_launch(m, r0, r1, ncols, rect_tiles, &a, p);
}

}  // namespace dipb
