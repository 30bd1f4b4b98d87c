// Neighbor joining entry points (nj.cu: exhaustive search; nj_pruned.cu: bound-pruned search).
#pragma once
#include "common.cuh"

namespace dipb {
struct NJState;
int nj_run(dipb_matrix* m, int algo, int32_t* child0, int32_t* child1, double* len0, double* len1);
int nj_pruned_loop(dipb_matrix* m, double* U, double* u, double* partial, NJState* st, int* realID, int32_t* c0,
                   int32_t* c1, double* l0, double* l1);
bool nj_cluster_fits(int n);
int nj_cluster_loop(dipb_matrix* m, double* U, double* u, int* realID, int32_t* c0, int32_t* c1, double* l0, double* l1);
}  // namespace dipb
