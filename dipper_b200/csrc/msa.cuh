// Aligned-MSA device object and tile constants (see msa_dist.cu).
#pragma once
#include "common.cuh"

namespace dipb {
constexpr int MSA_TS = 128;                                  // sequences per block / tile edge
constexpr int MSA_KC = 16;                                   // 32-site words per k chunk
constexpr int MSA_THREADS = 512;
constexpr int MSA_STAGES = 3;
constexpr int MSA_SLAB_WORDS = 3 * MSA_KC * MSA_TS;          // one (seq block, k chunk): 6144 words = 24 KB
constexpr int MSA_STAGE_BYTES = 2 * MSA_SLAB_WORDS * 4;      // A slab + B slab
constexpr int MSA_SMEM_BYTES = MSA_STAGES * MSA_STAGE_BYTES + 64;
constexpr int MSA_MAX_CHUNKS = 127;                          // 127 * 512 = 65 024 sites fit the 16-bit packed counters
}  // namespace dipb

struct dipb_msa {
    dipb_ctx* ctx = nullptr;
    int n = 0, npad = 0, seq_len = 0, nkc = 0;
    uint32_t* planes = nullptr;  // [npad/128][nkc][3][16][128]
    int* nv = nullptr;           // valid sites per sequence [npad]
    // tensor-core operands (msa_tc.cu), built on first use: simplex int8 [tc_rows][tc_ks], validity int8 [tc_rows][tc_kv]
    // Persistent buffers hold the expanded rows [0, tc_have) out of a capacity of tc_rows (tc_reserve rows are requested
    // by the caller that knows how many it needs: all for a full matrix / placement, the backbone for D&C); row
    // blocks beyond the capacity (D&C query batches) are expanded into the scratch pair just before they are used.
    int8_t* tc_S = nullptr;
    int8_t* tc_V = nullptr;
    size_t tc_ks = 0, tc_kv = 0, tc_rows = 0, tc_have = 0, tc_reserve = 0;
    int tc_fmt = 0;              // operand encoding of the buffers below: 0 int8, 2 e2m1 (two elements per byte; tc_ks / tc_kv are BYTES per row), fixed at the first expansion
    int8_t* tc_Sx = nullptr;
    int8_t* tc_Vx = nullptr;
    size_t tc_xrows = 0;
};

namespace dipb {
int msa_repack(dipb_msa* m, const uint64_t* d_in, int comp64);
int msa_block(dipb_msa* m, int type, int r0, int r1, int ncols, double* d_out, size_t ld);
int msa_matrix(dipb_msa* m, int type, int row_begin, int row_end, double* d_out);
int msa_counts_dev(dipb_msa* m, int i0, int i1, int j1, int* d_match, int* d_both, size_t ld);
bool msa_tc_supported(const dipb_msa* m, int type);
int msa_tc_matrix(dipb_msa* m, int type, int row_begin, int row_end, double* d_out);
int msa_tc_block(dipb_msa* m, int type, int r0, int r1, int ncols, double* d_out, size_t ld);
void msa_tc_reserve(dipb_msa* m, int rows);   // how many leading rows the persistent operand buffers should hold
}  // namespace dipb
