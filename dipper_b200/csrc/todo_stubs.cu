// Entry points declared in include/dipper_b200.h whose kernels are not built yet.
// They fail loudly (no CPU fallback).
#include "common.cuh"
#include "nj.cuh"
using namespace dipb;
#define NOTYET(name) do { set_error(name ": not implemented in this build"); return DIPB_E_STATE; } while (0)
extern "C" {
}
