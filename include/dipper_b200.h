/*
 * dipper_b200.h -- C ABI of the B200-native DIPPER hot path (libdipper_b200.so).
 *
 * DIPPER has no FFI: the seam this library sits behind is the C++ struct API of
 * src/mash_placement.cuh in the reference (paths below are relative to the reference
 * root).  Every entry point cites the reference interface it replaces.  All functions
 * return 0 on success or a negative DIPB_E_* code; dipb_last_error() returns the
 * message of the last failure on the calling thread.  Handles are opaque.  The caller
 * keeps ownership of every host buffer it passes in; the library owns device memory.
 * One caller thread per context.  There is no CPU fallback: without a CUDA device
 * dipb_init fails.
 */
#ifndef DIPPER_B200_H
#define DIPPER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIPB_OK 0
#define DIPB_E_CUDA -1      /* a CUDA runtime call failed (reference: fprintf + exit(1)) */
#define DIPB_E_ARG -2       /* invalid argument */
#define DIPB_E_NOMEM -3     /* host or device allocation failed */
#define DIPB_E_STATE -4     /* call order violated (e.g. distances before sketching) */
#define DIPB_E_UNSUPPORTED -5 /* the requested kernel variant does not fit this device / input size */

/* distance models, same numbering as -d / Param::distanceType (src/MSA.cu:81-86) */
#define DIPB_DIST_UNCORRECTED 1
#define DIPB_DIST_JC 2
#define DIPB_DIST_TAJIMANEI 3
#define DIPB_DIST_K2P 4
#define DIPB_DIST_TAMURA 5
#define DIPB_DIST_JINNEI 6

/* NJ search strategies; every one returns the reference's argmin (same tie-break) */
#define DIPB_NJ_AUTO 0
#define DIPB_NJ_FULLSCAN 1  /* scans all n^2 candidates each iteration (reference cost model) */
#define DIPB_NJ_PRUNED 2    /* exact lower-bound pruning, rescans only rows that can hold the minimum (whole-grid kernel) */
#define DIPB_NJ_CLUSTER 3   /* the same pruned search run by one 16-CTA thread-block cluster (default when it fits) */

typedef struct dipb_ctx dipb_ctx;
typedef struct dipb_msa dipb_msa;
typedef struct dipb_mash dipb_mash;
typedef struct dipb_matrix dipb_matrix;
typedef struct dipb_tree dipb_tree;

/* ---- context ------------------------------------------------------------ */
/* replaces cudaSetDevice(1) in src/tree_generation.cu:240-245 (device is a parameter) */
int dipb_init(int device, dipb_ctx **out);
/* Drops the creator's reference.  Every child handle (msa, mash, matrix, tree, D&C state) holds its own
 * reference on the context, so children may be freed AFTER dipb_destroy, in any order: the stream and the
 * context go away with the last reference (the reference's structs are global statics that are never torn
 * down, src/mash_placement.cuh:287-300).  A second dipb_destroy on the same context is ignored. */
void dipb_destroy(dipb_ctx *ctx);
/* live references on the context (1 = only the creator's); for tests */
int dipb_ctx_refs(const dipb_ctx *ctx);
const char *dipb_last_error(void);
const char *dipb_version(void);
/* device-side duration (CUDA events on the context's stream) of the most recent call
 * of the given kind on this context, in milliseconds; <0 if it has not run. */
#define DIPB_T_MSA_UPLOAD 0
#define DIPB_T_MSA_DIST 1
#define DIPB_T_NJ 2
#define DIPB_T_SKETCH 3
#define DIPB_T_MASH_DIST 4
#define DIPB_T_PLACE 5
#define DIPB_T_NJ_SCAN 6
#define DIPB_T_COUNT 8
double dipb_elapsed_ms(dipb_ctx *ctx, int what);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t dipb_kernel_launches(dipb_ctx *ctx);
int dipb_sync(dipb_ctx *ctx);

/* ---- aligned MSA distances ---------------------------------------------- */
/* MSADeviceArrays::allocateDeviceArrays(uint64_t**, uint64_t*, size_t, Param&)
 * src/mash_placement.cuh:95, src/MSA.cu:14-72.  seq4[i] = 4-bit packed sequence i
 * (fourBitCompressor layout), len[i] in sites; seqLen = len[0] as in the reference. */
int dipb_msa_upload(dipb_ctx *ctx, const uint64_t *const *seq4, const uint64_t *len, size_t n, dipb_msa **out);
/* same, sequences contiguous: flat[i * ceil(seq_len/16) + w] */
int dipb_msa_upload_flat(dipb_ctx *ctx, const uint64_t *flat, size_t n, uint64_t seq_len, dipb_msa **out);
void dipb_msa_free(dipb_msa *msa);
/* Releases the tensor-core operand buffers (e2m1 / int8 expansion of the packed sequences, 2 / 4 bytes per site) that the first
 * p / JC matrix or row block builds and keeps for later calls; the next such call rebuilds them.  No reference
 * counterpart (the reference keeps only the packed sequences, src/MSA.cu:14-72); used to return 3.6 GB at 30 000 x
 * 30 000 and by bench.py so that every timed step pays for the expansion. */
int dipb_msa_drop_operands(dipb_msa *msa);
/* MSADeviceArrays::distConstructionOnGpu(Param&, int rowId, double* d_out) const
 * src/mash_placement.cuh:97, src/MSA.cu:271-282: d_out[j] = d(row, j), j < row,
 * d_out is a DEVICE pointer with room for >= row doubles. */
int dipb_msa_dist_row(dipb_msa *msa, int dist_type, int row, double *d_out);
/* host-output twin used by the CLI and the tests */
int dipb_msa_dist_row_host(dipb_msa *msa, int dist_type, int row, double *h_out);
/* rows [r0,r1) x columns [0,ncols) into a device buffer with leading dimension ld:
 * the batched form placement / divide-and-conquer consume (DC/msa.cu:269-372). */
int dipb_msa_dist_block(dipb_msa *msa, int dist_type, int r0, int r1, int ncols, double *d_out, size_t ld);
/* test hook: bit-exact (match, useful) of src/MSA.cu:89-100 for rows [i0,i1) x cols [j0,j1), host int32 [i][j] */
int dipb_msa_counts(dipb_msa *msa, int i0, int i1, int j0, int j1, int32_t *h_match, int32_t *h_useful);
/* NJDeviceArrays::getDismatrix + fillDismatrix, src/neighborJoining.cu:20-85: dense
 * symmetric fp64 n x n, zero diagonal.  row_begin/row_end restrict the computed rows
 * (row-block sharding across GPUs); pass 0, n for everything. */
int dipb_msa_dist_matrix(dipb_msa *msa, int dist_type, dipb_matrix **out);
int dipb_msa_dist_matrix_rows(dipb_msa *msa, int dist_type, int row_begin, int row_end, dipb_matrix **out);

/* ---- unaligned: MinHash sketches + Mash distance ------------------------ */
/* MashDeviceArrays::allocateDeviceArrays src/mash_placement.cuh:44, src/mash.cu:14-122.
 * seq2[i] = 2-bit packed (twoBitCompressor layout), len[i] in bases. */
int dipb_mash_upload(dipb_ctx *ctx, const uint64_t *const *seq2, const uint64_t *len, size_t n, int k, int s,
                     dipb_mash **out);
int dipb_mash_upload_flat(dipb_ctx *ctx, const uint64_t *flat, const uint64_t *word_off, const uint64_t *len,
                          size_t n, int k, int s, dipb_mash **out);
void dipb_mash_free(dipb_mash *m);
/* MashDeviceArrays::sketchConstructionOnGpu(Param&) src/mash_placement.cuh:46, src/mash.cu:386-424 */
int dipb_mash_sketch(dipb_mash *m);
/* bit-exact check hook: h_out[i * s + t], ascending, ~0-padded (src/mash.cu:284-360) */
int dipb_mash_get_sketches(dipb_mash *m, uint64_t *h_out);
/* load externally built sketches ([n][s] row-major) instead of sketching */
int dipb_mash_set_sketches(dipb_ctx *ctx, const uint64_t *h_sk, size_t n, int k, int s, dipb_mash **out);
/* MashDeviceArrays::distConstructionOnGpu src/mash_placement.cuh:48, src/mash.cu:457-471 */
int dipb_mash_dist_row(dipb_mash *m, int row, double *d_out);
int dipb_mash_dist_row_host(dipb_mash *m, int row, double *h_out);
int dipb_mash_dist_block(dipb_mash *m, int r0, int r1, int ncols, double *d_out, size_t ld);
int dipb_mash_dist_matrix(dipb_mash *m, dipb_matrix **out);

/* ---- distance matrices --------------------------------------------------- */
/* -i d: MatrixReader rows (src/matrix_reader.cu:15-44) already parsed on the host.
 * full != 0: h holds n*n doubles; else h holds the packed lower triangle
 * (row i contributes i entries).  The matrix is mirrored with a zero diagonal. */
int dipb_matrix_from_host(dipb_ctx *ctx, const double *h, int n, int full, dipb_matrix **out);
int dipb_matrix_n(const dipb_matrix *m);
int dipb_matrix_to_host(dipb_matrix *m, double *h_out /* n*n */);
double *dipb_matrix_device_ptr(dipb_matrix *m);
/* multi-GPU gather of a row-sharded matrix: after the rows [r0, r1) of another rank's shard have been copied into this
 * matrix (they are contiguous), fill the mirrored entries D[j][i] = D[i][j], j < i, r0 <= i < r1 (fillDismatrix :20-32) */
int dipb_matrix_mirror_rows(dipb_matrix *m, int r0, int r1);
void dipb_matrix_free(dipb_matrix *m);

/* ---- neighbor joining ---------------------------------------------------- */
/* NJDeviceArrays::findNeighbourJoiningTree src/mash_placement.cuh:212,
 * src/neighborJoining.cu:197-271.  Consumes (overwrites) the matrix like the
 * reference.  Internal node n+k (k = 0..n-2) gets children child0[k], child1[k] with
 * branch lengths len0[k], len1[k], in the reference's push order; node 2n-2 is the root. */
int dipb_nj(dipb_matrix *m, int algo, int32_t *child0, int32_t *child1, double *len0, double *len1);
/* counters of the last dipb_nj on this matrix' context: rows rescanned and matrix
 * bytes read by the search (for the roofline's traffic figure) */
int dipb_nj_stats(dipb_ctx *ctx, uint64_t *rows_scanned, uint64_t *bytes_scanned, uint64_t *iterations);

/* ---- k-closest placement -------------------------------------------------- */
/* distance source for placement / add-tips: exactly one of msa / mash / matrix */
typedef struct {
    dipb_msa *msa;
    int dist_type;
    dipb_mash *mash;
    dipb_matrix *matrix; /* full n x n, row i, j < i is used */
} dipb_dist_source;

/* KPlacementDeviceArrays::allocateDeviceArrays + findPlacementTree
 * src/mash_placement.cuh:180-188, src/placement_close_k.cu:15-68,646-854 */
int dipb_place_kclosest(dipb_ctx *ctx, const dipb_dist_source *src, int n, dipb_tree **out);
/* PlacementDeviceArrays::allocateDeviceArrays + findPlacementTree (exact placement mode, -p 0)
 * src/mash_placement.cuh:137-165, src/placement.cu:28-116,505-789.  Same slot arrays as above
 * (print from node n, src/placement.cu:500).  Up to dipb_place_exact_max_tips() tips the tree
 * lives in the shared memory of one thread-block cluster; larger inputs run the same data flow
 * through global memory on the whole grid.  DIPB_E_UNSUPPORTED when a tip has no candidate edge
 * with pendant length < 2 (see placement_exact.cu). */
int dipb_place_exact(dipb_ctx *ctx, const dipb_dist_source *src, int n, dipb_tree **out);
int dipb_place_exact_max_tips(void);
/* initializeDeviceArrays(Tree*) + addQuery, src/placement_close_k.cu:126-264,858-990.
 * The backbone is given as the adjacency arrays the reference builds from its Tree
 * (host pointers; see dipb_backbone_from_newick in dipper_host.h): 4B-4 slots. */
int dipb_place_add(dipb_ctx *ctx, const dipb_dist_source *src, int n, int backbone, const int32_t *h_head,
                   const int32_t *h_e, const int32_t *h_nxt, const int32_t *h_belong, const double *h_len,
                   dipb_tree **out);
/* divide and conquer (-m 3): KPlacementDeviceArraysDC::{findBackboneTreeDC, findClustersDC,
 * findClusterTreeDC}, DC/placement_close_k.cu:731-1535.  Tips 0..backbone-1 form the
 * backbone (placed as above, internal node ids offset by n), every other tip is assigned
 * to its best backbone edge and the per-edge clusters are placed independently.
 * The reference uses backbone = n / 20 (src/tree_generation.cu:425,545). */
int dipb_dc(dipb_ctx *ctx, const dipb_dist_source *src, int n, int backbone, dipb_tree **out);
/* The same three stages as separate calls, so that one process per GPU can shard stage 2
 * (queries are independent) and stage 3 (clusters touch disjoint slots) and merge on rank 0
 * (SURVEY.md §8e).  Every rank: dipb_dc_begin (backbone; deterministic, so identical on all
 * ranks) -> dipb_dc_assign(own query range) -> all-gather the cluster ids ->
 * dipb_dc_set_clusters -> dipb_dc_run_clusters(own cluster range) -> dipb_dc_export_slice;
 * rank 0 imports the other ranks' slices and calls dipb_dc_finish.  Slot and node numbers are
 * global prefix sums, identical to the single-GPU run. */
typedef struct dipb_dc_state dipb_dc_state;
int dipb_dc_begin(dipb_ctx *ctx, const dipb_dist_source *src, int n, int backbone, dipb_dc_state **out);
int dipb_dc_assign(dipb_dc_state *st, int q0, int q1, int32_t *h_cluster /* q1-q0 */);
int dipb_dc_set_clusters(dipb_dc_state *st, const int32_t *h_cluster_all /* n */, int *num_clusters);
int dipb_dc_cluster_sizes(dipb_dc_state *st, int32_t *h_sizes /* num_clusters */);
int dipb_dc_run_clusters(dipb_dc_state *st, int c0, int c1);
/* h_buf == NULL: only report the size in *bytes */
int dipb_dc_export_slice(dipb_dc_state *st, int c0, int c1, void *h_buf, size_t cap, size_t *bytes);
int dipb_dc_import_slice(dipb_dc_state *st, const void *h_buf, size_t bytes);
int dipb_dc_finish(dipb_dc_state *st, dipb_tree **out /* may be NULL */);
/* test hook: cluster (backbone slot) of every tip after the last dipb_dc on this context;
 * h_out[n], entries of backbone tips are -1 */
int dipb_dc_cluster_ids(dipb_ctx *ctx, int32_t *h_out, int n);

/* printTree's D2H half (src/placement_close_k.cu:595-618): head[2n], e/nxt/belong[8n], len[8n] */
int dipb_tree_export(dipb_tree *t, int32_t *head, int32_t *e, int32_t *nxt, int32_t *belong, double *len);
/* closest lists, for parity tests: cid[40n], cdis[40n] */
int dipb_tree_export_closest(dipb_tree *t, int32_t *cid, double *cdis);
/* The device arrays themselves -- the reference exposes them as public fields of its structs (d_head, d_e, d_nxt,
 * d_belong, d_len, d_closest_id, d_closest_dis; src/mash_placement.cuh:167-197): head[2n], e/nxt/belong[8n], len[8n],
 * closest lists 5 per slot [40n].  Owned by the tree; valid until dipb_tree_free. */
int dipb_tree_device_arrays(dipb_tree *t, int32_t **head, int32_t **e, int32_t **nxt, int32_t **belong, double **len,
                            int32_t **closest_id, double **closest_dis);
int dipb_tree_n(const dipb_tree *t);
void dipb_tree_free(dipb_tree *t);

/* ---- several GPUs of one box, one process -------------------------------------------------------------------
 * The reference is single GPU (device 1 hard-coded, src/tree_generation.cu:240-245).  Here one dipb_ctx per device is
 * driven by one host thread each; shards are exchanged with peer copies over NVLink:
 *   dipb_multi_msa_dist_matrix  row blocks balanced by triangle area, gathered (lower trapezoids only) and mirrored
 *                               on devices[0]: the matrix NJDeviceArrays::getDismatrix builds (src/neighborJoining.cu:35-85)
 *   dipb_multi_dc               -m 3 with the queries of the cluster-assignment stage split over the devices
 *                               (src/divide_and_conquer/placement_close_k.cu:937-1113); the tree ends on devices[0]
 * dipb_multi_elapsed_ms(what): host wall clock of the last call: 0 sharded compute, 1 gather, 2 the rest, 3 total. */
typedef struct dipb_multi dipb_multi;
int dipb_multi_init(const int *devices, int n_devices, dipb_multi **out);
void dipb_multi_destroy(dipb_multi *m);
int dipb_multi_devices(const dipb_multi *m);
dipb_ctx *dipb_multi_ctx(dipb_multi *m, int d);
double dipb_multi_elapsed_ms(const dipb_multi *m, int what);
int dipb_multi_msa_upload_flat(dipb_multi *m, const uint64_t *flat, size_t n, uint64_t seq_len);
int dipb_multi_msa_dist_matrix(dipb_multi *m, int dist_type, dipb_matrix **out);
int dipb_multi_dc(dipb_multi *m, int dist_type, int backbone, dipb_tree **out);
int dipb_multi_dc_cluster_ids(const dipb_multi *m, int32_t *h_out, int n);

#ifdef __cplusplus
}
#endif
#endif
