// Entry points declared in include/dipper_b200.h whose kernels are not built yet.
// They fail loudly (no CPU fallback).
#include "common.cuh"
#include "nj.cuh"
using namespace dipb;
#define NOTYET(name) do { set_error(name ": not implemented in this build"); return DIPB_E_STATE; } while (0)
extern "C" {
int dipb_dc(dipb_ctx*, const dipb_dist_source*, int, int, dipb_tree**) { NOTYET("dipb_dc"); }
int dipb_dc_cluster_ids(dipb_ctx*, int32_t*, int) { NOTYET("dipb_dc_cluster_ids"); }
}
