// Entry points declared in include/dipper_b200.h whose kernels are not built yet.
// They fail loudly (no CPU fallback).
#include "common.cuh"
#include "nj.cuh"
using namespace dipb;
#define NOTYET(name) do { set_error(name ": not implemented in this build"); return DIPB_E_STATE; } while (0)
extern "C" {
int dipb_place_kclosest(dipb_ctx*, const dipb_dist_source*, int, dipb_tree**) { NOTYET("dipb_place_kclosest"); }
int dipb_place_add(dipb_ctx*, const dipb_dist_source*, int, int, const int32_t*, const int32_t*, const int32_t*, const int32_t*, const double*, dipb_tree**) { NOTYET("dipb_place_add"); }
int dipb_tree_export(dipb_tree*, int32_t*, int32_t*, int32_t*, int32_t*, double*) { NOTYET("dipb_tree_export"); }
int dipb_tree_export_closest(dipb_tree*, int32_t*, double*) { NOTYET("dipb_tree_export_closest"); }
int dipb_tree_n(const dipb_tree*) { return 0; }
void dipb_tree_free(dipb_tree*) {}
}
