/*
 * dipper_host.h -- host-side C ABI of dipper_b200: the encoders, readers and Newick
 * code that stay on the CPU in the reference as well (L1/L5 in SURVEY.md).  Exported
 * from the same libdipper_b200.so so that the CLI, the tests and foreign bindings
 * share one implementation.  Paths cited are relative to the reference root.
 */
#ifndef DIPPER_HOST_H
#define DIPPER_HOST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* fourBitCompressor(std::string, size_t, uint64_t*)  src/fourBitCompressor.cpp:5-41 */
void dipb_pack4(const char *seq, size_t len, uint64_t *out /* ceil(len/16) words */);
/* twoBitCompressor(std::string, size_t, uint64_t*)   src/twoBitCompressor.cpp:5-41 */
void dipb_pack2(const char *seq, size_t len, uint64_t *out /* ceil(len/32) words */);

/* FASTA ingest: readSequences (src/tree_generation.cu:132-154, one kseq loop building std::string per record) and
 * the packing loops (:350-362 four-bit, :478-490 two-bit) in one pass over a memory-mapped file with all host threads:
 * record starts are found in parallel, every record is measured and then packed straight into one flat word array.
 * kseq semantics: a record starts at '>' in column 0, its name is the header up to the first white space, its
 * sequence is every graphic character up to the next record; text before the first '>' is ignored.
 * bits = 4: fourBitCompressor codes, 16 per word; bits = 2: twoBitCompressor codes, 32 per word.
 * Plain (uncompressed) files only: returns DIPB_E_UNSUPPORTED for gzip input (the CLI then uses its zlib reader). */
typedef struct dipb_fasta dipb_fasta;
int dipb_fasta_open(const char *path, int bits, int threads /* 0 = all hardware threads */, dipb_fasta **out);
size_t dipb_fasta_count(const dipb_fasta *f);
const char *dipb_fasta_name(const dipb_fasta *f, size_t i);
const uint64_t *dipb_fasta_lengths(const dipb_fasta *f);      /* [count] characters per sequence */
const uint64_t *dipb_fasta_word_offsets(const dipb_fasta *f); /* [count + 1] first word of each sequence */
const uint64_t *dipb_fasta_words(const dipb_fasta *f);        /* packed sequences, back to back */
void dipb_fasta_close(dipb_fasta *f);

/* Newick of an NJ result (src/neighborJoining.cu:252-270: children in push order,
 * lengths as ostream<<double, trailing ";\n").  Returns a malloc'd string (dipb_free_str). */
char *dipb_nj_newick(int n, const int32_t *child0, const int32_t *child1, const double *len0, const double *len1,
                     const char *const *names);
/* Newick of a placement tree (src/placement_close_k.cu:568-643): root_node = n for
 * dipper's own trees, children in adjacency-list order, leaf = single adjacency. */
char *dipb_tree_newick(int n_nodes, int root_node, const int32_t *head, const int32_t *e, const int32_t *nxt,
                       const double *len, const char *const *names);
void dipb_free_str(char *s);

/* The branch-length formatter of the Newick writers: "%g" (ostream << double, src/tree.cpp printers), exact and ~4x
 * faster than snprintf for 1e-15 <= |v| < 1e6, snprintf otherwise.  Writes at most 31 characters + NUL, returns the length. */
int dipb_format_g(double v, char *out32);

/* Tree(std::string newick, size_t totalLeaves) + KPlacementDeviceArrays::initializeDeviceArrays
 * host half (src/tree.cpp:216-361, src/placement_close_k.cu:144-183): parses a rooted
 * Newick whose every non-root node carries a branch length; leaves are numbered in order
 * of appearance, internal nodes total_leaves, total_leaves+1, ... in order of '('.
 * Fills the adjacency arrays (sizes: head 2*total_leaves, others 8*total_leaves) and
 * returns the number of backbone leaves, or a negative DIPB_E_* code.
 * leaf_names_out receives a malloc'd, '\n'-separated list of leaf names in index order. */
int dipb_backbone_from_newick(const char *newick, int total_leaves, int32_t *head, int32_t *e, int32_t *nxt,
                              int32_t *belong, double *len, char **leaf_names_out);

/* -o d: distance matrix in PHYLIP format.  The reference documents the option ("coming soon", docs/index.md:114) but
 * does not implement it; the text written here is what its own reader accepts (MatrixReader, src/matrix_reader.cu:15-44:
 * first line n, then one line per taxon: name, then the distances to the taxa before it (lower = 1) or to all taxa
 * (lower = 0), parsed there with stof).  Values are written with 9 significant digits (round-trips through float32).
 * D is the dense n x n matrix, row-major.  Returns 0 or a negative DIPB_E_* code. */
int dipb_phylip_write(const char *path, int n, const double *D, const char *const *names, int lower);

#ifdef __cplusplus
}
#endif
#endif
