"""ctypes front-end of the CPU oracle (oracle/dipper_oracle.c).  TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package dipper_b200 never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


class PairStats(C.Structure):
    _fields_ = [("useful", C.c_int), ("match", C.c_int), ("tot", C.c_int), ("ts", C.c_int), ("tv", C.c_int),
                ("frac", C.c_int * 4), ("pr", C.c_int * 4), ("gc_row", C.c_int), ("gc_col", C.c_int)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_pack4.argtypes = [C.c_char_p, C.c_uint64, u64p]
        L.orc_pack2.argtypes = [C.c_char_p, C.c_uint64, u64p]
        L.orc_pair_stats_compute.argtypes = [u64p, u64p, C.c_int, C.POINTER(PairStats)]
        L.orc_dist_from_stats.argtypes = [C.POINTER(PairStats), C.c_int]
        L.orc_dist_from_stats.restype = C.c_double
        L.orc_msa_counts.argtypes = [u64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p, i32p]
        L.orc_msa_dist_row.argtypes = [u64p, C.c_int, C.c_int, C.c_int, f64p]
        L.orc_msa_dist_matrix.argtypes = [u64p, C.c_int, C.c_int, C.c_int, f64p]
        L.orc_fast_counts.argtypes = [u64p, u64p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_murmur3_x64_128.argtypes = [C.c_char_p, C.c_int, C.c_uint32, u64p]
        L.orc_kmer_hash.argtypes = [u64p, C.c_uint64, C.c_int]
        L.orc_kmer_hash.restype = C.c_uint64
        L.orc_sketch.argtypes = [u64p, C.c_uint64, C.c_int, C.c_int, u64p]
        L.orc_sketch_all.argtypes = [u64p, u64p, u64p, C.c_int, C.c_int, C.c_int, u64p]
        L.orc_mash_dist.argtypes = [u64p, u64p, C.c_int, C.c_int]
        L.orc_mash_dist.restype = C.c_double
        L.orc_mash_inter_uni.argtypes = [u64p, u64p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_mash_dist_row.argtypes = [u64p, C.c_int, C.c_int, C.c_int, f64p]
        L.orc_mash_dist_matrix.argtypes = [u64p, C.c_int, C.c_int, C.c_int, f64p]
        L.orc_canon_sum.argtypes = [f64p, C.c_int]
        L.orc_canon_sum.restype = C.c_double
        L.orc_nj.argtypes = [f64p, C.c_int, i32p, i32p, f64p, f64p]
        L.orc_nj.restype = C.c_int
        L.orc_nj_newick.argtypes = [C.c_int, i32p, i32p, f64p, f64p, C.POINTER(C.c_char_p)]
        L.orc_nj_newick.restype = C.c_void_p
        L.orc_ptree_new.argtypes = [C.c_int]
        L.orc_ptree_new.restype = C.c_void_p
        L.orc_ptree_free.argtypes = [C.c_void_p]
        L.orc_place_all_matrix.argtypes = [f64p, C.c_int]
        L.orc_place_all_matrix.restype = C.c_void_p
        L.orc_place_exact_matrix.argtypes = [f64p, C.c_int]
        L.orc_place_exact_matrix.restype = C.c_void_p
        L.orc_place_add_matrix.argtypes = [C.c_void_p, f64p, C.c_int, C.c_int]
        L.orc_ptree_load_backbone.argtypes = [C.c_void_p, C.c_int, i32p, i32p, i32p, f64p, C.c_int]
        L.orc_dc_matrix.argtypes = [f64p, C.c_int, C.c_int, i32p]
        L.orc_dc_matrix.restype = C.c_void_p
        L.orc_dc_matrix_as_shipped.argtypes = [f64p, C.c_int, C.c_int, C.c_double, i32p]
        L.orc_dc_matrix_as_shipped.restype = C.c_void_p
        L.orc_dc_set_stage3_seed.argtypes = [C.c_double, C.c_int]
        L.orc_dc_set_stage3_seed.restype = None
        L.orc_ptree_newick.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_char_p)]
        L.orc_ptree_newick.restype = C.c_void_p
        for nm, ty in (("head", C.c_int), ("e", C.c_int), ("nxt", C.c_int), ("belong", C.c_int),
                       ("len", C.c_double), ("cid", C.c_int), ("cdis", C.c_double)):
            f = getattr(L, "orc_ptree_" + nm)
            f.argtypes = [C.c_void_p]
            f.restype = C.POINTER(ty)
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_num_threads.restype = C.c_int
        _LIB = L
    return _LIB


def num_threads():
    return int(lib().orc_num_threads())


def pack4(seq):
    b = seq if isinstance(seq, bytes) else seq.encode()
    out = np.zeros((len(b) + 15) // 16, np.uint64)
    lib().orc_pack4(b, len(b), out)
    return out


def pack2(seq):
    b = seq if isinstance(seq, bytes) else seq.encode()
    out = np.zeros((len(b) + 31) // 32, np.uint64)
    lib().orc_pack2(b, len(b), out)
    return out


def pair_stats(row, col, seq_len):
    st = PairStats()
    lib().orc_pair_stats_compute(np.ascontiguousarray(row), np.ascontiguousarray(col), seq_len, C.byref(st))
    return st


def dist_from_stats(st, dist_type):
    return float(lib().orc_dist_from_stats(C.byref(st), dist_type))


def msa_counts(seqs, seq_len, i0, i1, j0, j1):
    seqs = np.ascontiguousarray(seqs, np.uint64)
    m = np.zeros((i1 - i0, j1 - j0), np.int32)
    u = np.zeros((i1 - i0, j1 - j0), np.int32)
    lib().orc_msa_counts(seqs, seqs.shape[0], seq_len, i0, i1, j0, j1, m, u)
    return m, u


def msa_dist_row(seqs, seq_len, row, dist_type):
    out = np.zeros(max(row, 1), np.float64)
    lib().orc_msa_dist_row(np.ascontiguousarray(seqs, np.uint64), seq_len, row, dist_type, out)
    return out[:row]


def msa_dist_matrix(seqs, seq_len, dist_type):
    seqs = np.ascontiguousarray(seqs, np.uint64)
    n = seqs.shape[0]
    D = np.zeros((n, n), np.float64)
    lib().orc_msa_dist_matrix(seqs, n, seq_len, dist_type, D)
    return D


def murmur3_x64_128(key, seed=0):
    out = np.zeros(2, np.uint64)
    lib().orc_murmur3_x64_128(key, len(key), seed, out)
    return int(out[0]), int(out[1])


def sketch(seq2, length, k=15, s=1000):
    out = np.zeros(s, np.uint64)
    lib().orc_sketch(np.ascontiguousarray(seq2, np.uint64), length, k, s, out)
    return out


def sketch_all(flat, offsets, lens, k=15, s=1000):
    n = len(lens)
    out = np.zeros((n, s), np.uint64)
    lib().orc_sketch_all(np.ascontiguousarray(flat, np.uint64), np.ascontiguousarray(offsets, np.uint64),
                         np.ascontiguousarray(lens, np.uint64), n, k, s, out)
    return out


def mash_dist(A, B, k=15):
    A = np.ascontiguousarray(A, np.uint64)
    B = np.ascontiguousarray(B, np.uint64)
    return float(lib().orc_mash_dist(A, B, len(A), k))


def mash_inter_uni(A, B):
    i, u = C.c_int(), C.c_int()
    lib().orc_mash_inter_uni(np.ascontiguousarray(A, np.uint64), np.ascontiguousarray(B, np.uint64), len(A),
                             C.byref(i), C.byref(u))
    return i.value, u.value


def mash_dist_matrix(sk, k=15):
    sk = np.ascontiguousarray(sk, np.uint64)
    n, s = sk.shape
    D = np.zeros((n, n), np.float64)
    lib().orc_mash_dist_matrix(sk, n, s, k, D)
    return D


def canon_sum(v):
    v = np.ascontiguousarray(v, np.float64)
    return float(lib().orc_canon_sum(v, len(v)))


def nj(D):
    """Returns (child0, child1, len0, len1) for internal nodes n..2n-2."""
    D = np.ascontiguousarray(D, np.float64)
    n = D.shape[0]
    c0 = np.zeros(n - 1, np.int32)
    c1 = np.zeros(n - 1, np.int32)
    l0 = np.zeros(n - 1, np.float64)
    l1 = np.zeros(n - 1, np.float64)
    rc = lib().orc_nj(D, n, c0, c1, l0, l1)
    if rc != 0:
        raise MemoryError("orc_nj")
    return c0, c1, l0, l1


def _names_arr(names):
    arr = (C.c_char_p * len(names))()
    arr[:] = [x.encode() if isinstance(x, str) else x for x in names]
    return arr


def _take_str(ptr):
    s = C.string_at(ptr).decode()
    lib().orc_free(ptr)
    return s


def nj_newick(c0, c1, l0, l1, names):
    n = len(c0) + 1
    return _take_str(lib().orc_nj_newick(n, c0, c1, l0, l1, _names_arr(names)))


class PTree:
    """k-closest placement tree arrays (src/mash_placement.cuh:167-197)."""

    def __init__(self, handle, n):
        self.h, self.n = handle, n

    def arrays(self):
        L, n = lib(), self.n
        g = lambda nm, cnt, dt: np.ctypeslib.as_array(getattr(L, "orc_ptree_" + nm)(self.h), (cnt,)).astype(dt).copy()
        return dict(head=g("head", 2 * n, np.int32), e=g("e", 8 * n, np.int32), nxt=g("nxt", 8 * n, np.int32),
                    belong=g("belong", 8 * n, np.int32), len=g("len", 8 * n, np.float64),
                    cid=g("cid", 40 * n, np.int32), cdis=g("cdis", 40 * n, np.float64))

    def newick(self, names, root=None):
        names = list(names) + [""] * (2 * self.n - len(names))
        return _take_str(lib().orc_ptree_newick(self.h, self.n if root is None else root, _names_arr(names)))

    def __del__(self):
        try:
            lib().orc_ptree_free(self.h)
        except Exception:
            pass


def place_all(D):
    D = np.ascontiguousarray(D, np.float64)
    return PTree(lib().orc_place_all_matrix(D, D.shape[0]), D.shape[0])


def place_exact(D):
    """Exact placement mode (src/placement.cu:505-789) from a full distance matrix; print from node n."""
    D = np.ascontiguousarray(D, np.float64)
    return PTree(lib().orc_place_exact_matrix(D, D.shape[0]), D.shape[0])


def place_add(D, B, root, child_off, child_idx, parent, bl):
    """Load a B-leaf backbone (ids: leaves 0..B-1, internals n..) then add tips B..n-1."""
    D = np.ascontiguousarray(D, np.float64)
    n = D.shape[0]
    t = PTree(lib().orc_ptree_new(n), n)
    lib().orc_ptree_load_backbone(t.h, root, np.ascontiguousarray(child_off, np.int32),
                                  np.ascontiguousarray(child_idx, np.int32), np.ascontiguousarray(parent, np.int32),
                                  np.ascontiguousarray(bl, np.float64), B)
    lib().orc_place_add_matrix(t.h, D, n, B)
    return t


def dc_as_shipped(D, B, stale=0.0, stage3_seed=0.0):
    """The reference's aligned D&C with its defect B17 switched on (see orc_dc_matrix_as_shipped) and, optionally, the
    stale stage-3 BFS seed of defect B10 (orc_dc_set_stage3_seed): checker of the restatement against the reference's
    own output, never the product rule."""
    D = np.ascontiguousarray(D, np.float64)
    n = D.shape[0]
    cl = np.zeros(n, np.int32)
    lib().orc_dc_set_stage3_seed(float(stage3_seed), -1)
    try:
        return PTree(lib().orc_dc_matrix_as_shipped(D, n, B, float(stale), cl), n), cl
    finally:
        lib().orc_dc_set_stage3_seed(0.0, -1)


def dc(D, B, stage3_seed=0.0):
    """Divide-and-conquer tree (DC/placement_close_k.cu:731-1535); returns (PTree, cluster ids)."""
    D = np.ascontiguousarray(D, np.float64)
    n = D.shape[0]
    cl = np.zeros(n, np.int32)
    lib().orc_dc_set_stage3_seed(float(stage3_seed), -1)
    try:
        return PTree(lib().orc_dc_matrix(D, n, B, cl), n), cl
    finally:
        lib().orc_dc_set_stage3_seed(0.0, -1)
