"""ctypes binding of libdipper_b200.so (include/dipper_b200.h, include/dipper_host.h).

There is no fallback: if the shared library is missing, or no CUDA device is visible
when a context is created, this raises.  Nothing under oracle/ is imported here.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdipper_b200.so")

u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
vp = C.c_void_p
vpp = C.POINTER(C.c_void_p)


class DistSource(C.Structure):
    _fields_ = [("msa", vp), ("dist_type", C.c_int), ("mash", vp), ("matrix", vp)]


class DipperError(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); every symbol declared in include/*.h is listed here and
# tests/test_abi.py checks the shared object exports all of them.
SIGNATURES = {
    "dipb_init": (C.c_int, [C.c_int, vpp]),
    "dipb_destroy": (None, [vp]),
    "dipb_ctx_refs": (C.c_int, [vp]),
    "dipb_last_error": (C.c_char_p, []),
    "dipb_version": (C.c_char_p, []),
    "dipb_elapsed_ms": (C.c_double, [vp, C.c_int]),
    "dipb_kernel_launches": (C.c_uint64, [vp]),
    "dipb_sync": (C.c_int, [vp]),
    "dipb_msa_upload": (C.c_int, [vp, C.POINTER(C.c_void_p), u64p, C.c_size_t, vpp]),
    "dipb_msa_upload_flat": (C.c_int, [vp, u64p, C.c_size_t, C.c_uint64, vpp]),
    "dipb_msa_free": (None, [vp]),
    "dipb_msa_drop_operands": (C.c_int, [vp]),
    "dipb_msa_dist_row": (C.c_int, [vp, C.c_int, C.c_int, vp]),
    "dipb_msa_dist_row_host": (C.c_int, [vp, C.c_int, C.c_int, f64p]),
    "dipb_msa_dist_block": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_size_t]),
    "dipb_msa_counts": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, i32p, i32p]),
    "dipb_msa_dist_matrix": (C.c_int, [vp, C.c_int, vpp]),
    "dipb_msa_dist_matrix_rows": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vpp]),
    "dipb_mash_upload": (C.c_int, [vp, C.POINTER(C.c_void_p), u64p, C.c_size_t, C.c_int, C.c_int, vpp]),
    "dipb_mash_upload_flat": (C.c_int, [vp, u64p, u64p, u64p, C.c_size_t, C.c_int, C.c_int, vpp]),
    "dipb_mash_free": (None, [vp]),
    "dipb_mash_sketch": (C.c_int, [vp]),
    "dipb_mash_get_sketches": (C.c_int, [vp, u64p]),
    "dipb_mash_set_sketches": (C.c_int, [vp, u64p, C.c_size_t, C.c_int, C.c_int, vpp]),
    "dipb_mash_dist_row": (C.c_int, [vp, C.c_int, vp]),
    "dipb_mash_dist_row_host": (C.c_int, [vp, C.c_int, f64p]),
    "dipb_mash_dist_block": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, C.c_size_t]),
    "dipb_mash_dist_matrix": (C.c_int, [vp, vpp]),
    "dipb_matrix_from_host": (C.c_int, [vp, f64p, C.c_int, C.c_int, vpp]),
    "dipb_matrix_n": (C.c_int, [vp]),
    "dipb_matrix_to_host": (C.c_int, [vp, f64p]),
    "dipb_matrix_device_ptr": (vp, [vp]),
    "dipb_matrix_mirror_rows": (C.c_int, [vp, C.c_int, C.c_int]),
    "dipb_matrix_free": (None, [vp]),
    "dipb_nj": (C.c_int, [vp, C.c_int, i32p, i32p, f64p, f64p]),
    "dipb_nj_stats": (C.c_int, [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "dipb_place_kclosest": (C.c_int, [vp, C.POINTER(DistSource), C.c_int, vpp]),
    "dipb_place_exact": (C.c_int, [vp, C.POINTER(DistSource), C.c_int, vpp]),
    "dipb_place_exact_max_tips": (C.c_int, []),
    "dipb_place_add": (C.c_int, [vp, C.POINTER(DistSource), C.c_int, C.c_int, i32p, i32p, i32p, i32p, f64p, vpp]),
    "dipb_dc": (C.c_int, [vp, C.POINTER(DistSource), C.c_int, C.c_int, vpp]),
    "dipb_dc_cluster_ids": (C.c_int, [vp, i32p, C.c_int]),
    "dipb_dc_begin": (C.c_int, [vp, C.POINTER(DistSource), C.c_int, C.c_int, vpp]),
    "dipb_dc_assign": (C.c_int, [vp, C.c_int, C.c_int, i32p]),
    "dipb_dc_set_clusters": (C.c_int, [vp, i32p, C.POINTER(C.c_int)]),
    "dipb_dc_cluster_sizes": (C.c_int, [vp, i32p]),
    "dipb_dc_run_clusters": (C.c_int, [vp, C.c_int, C.c_int]),
    "dipb_dc_export_slice": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_size_t, C.POINTER(C.c_size_t)]),
    "dipb_dc_import_slice": (C.c_int, [vp, vp, C.c_size_t]),
    "dipb_dc_finish": (C.c_int, [vp, vpp]),
    "dipb_tree_export": (C.c_int, [vp, i32p, i32p, i32p, i32p, f64p]),
    "dipb_tree_export_closest": (C.c_int, [vp, i32p, f64p]),
    "dipb_tree_device_arrays": (C.c_int, [vp, vpp, vpp, vpp, vpp, vpp, vpp, vpp]),
    "dipb_tree_n": (C.c_int, [vp]),
    "dipb_tree_free": (None, [vp]),
    "dipb_multi_init": (C.c_int, [i32p, C.c_int, vpp]),
    "dipb_multi_destroy": (None, [vp]),
    "dipb_multi_devices": (C.c_int, [vp]),
    "dipb_multi_ctx": (vp, [vp, C.c_int]),
    "dipb_multi_elapsed_ms": (C.c_double, [vp, C.c_int]),
    "dipb_multi_msa_upload_flat": (C.c_int, [vp, u64p, C.c_size_t, C.c_uint64]),
    "dipb_multi_msa_dist_matrix": (C.c_int, [vp, C.c_int, vpp]),
    "dipb_multi_dc": (C.c_int, [vp, C.c_int, C.c_int, vpp]),
    "dipb_multi_dc_cluster_ids": (C.c_int, [vp, i32p, C.c_int]),
    # dipper_host.h
    "dipb_pack4": (None, [C.c_char_p, C.c_size_t, u64p]),
    "dipb_pack2": (None, [C.c_char_p, C.c_size_t, u64p]),
    "dipb_nj_newick": (vp, [C.c_int, i32p, i32p, f64p, f64p, C.POINTER(C.c_char_p)]),
    "dipb_tree_newick": (vp, [C.c_int, C.c_int, i32p, i32p, i32p, f64p, C.POINTER(C.c_char_p)]),
    "dipb_free_str": (None, [vp]),
    "dipb_format_g": (C.c_int, [C.c_double, C.c_char_p]),
    "dipb_fasta_open": (C.c_int, [C.c_char_p, C.c_int, C.c_int, vpp]),
    "dipb_fasta_count": (C.c_size_t, [vp]),
    "dipb_fasta_name": (C.c_char_p, [vp, C.c_size_t]),
    "dipb_fasta_lengths": (C.POINTER(C.c_uint64), [vp]),
    "dipb_fasta_word_offsets": (C.POINTER(C.c_uint64), [vp]),
    "dipb_fasta_words": (C.POINTER(C.c_uint64), [vp]),
    "dipb_fasta_close": (None, [vp]),
    "dipb_phylip_write": (C.c_int, [C.c_char_p, C.c_int, f64p, C.POINTER(C.c_char_p), C.c_int]),
    "dipb_backbone_from_newick": (C.c_int, [C.c_char_p, C.c_int, i32p, i32p, i32p, i32p, f64p, vpp]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DipperError(
                "libdipper_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C dipper_b200/csrc`. There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise DipperError("dipper_b200 error %d: %s" % (rc, lib().dipb_last_error().decode()))


def take_str(ptr):
    if not ptr:
        raise DipperError("dipper_b200: string allocation failed")
    s = C.string_at(ptr).decode()
    lib().dipb_free_str(ptr)
    return s


def names_array(names):
    """char** of the tip names.  A caller that builds many trees over the same tips converts once and passes the array
    itself (the reference holds its names in C++ strings: no per-tree conversion there either)."""
    if isinstance(names, C.Array):
        return names
    arr = (C.c_char_p * len(names))()
    arr[:] = [x.encode() if isinstance(x, str) else x for x in names]
    return arr
