"""Host-side mirror of the reference's operator interface (src/mash_placement.cuh).

Same struct names, method names and argument meaning as the reference so parity tests
read like calls into DIPPER itself; everything executes in libdipper_b200.so through
the C ABI.  Differences, all deliberate: instances are not global statics, the device
is a parameter (reference hard-codes device 1, src/tree_generation.cu:240), errors
raise DipperError instead of exit(1), and device rows come back as numpy arrays.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import DipperError, DistSource, check, lib

DIST_UNCORRECTED, DIST_JUKESCANTOR, DIST_TAJIMANEI, DIST_KIMURA2P, DIST_TAMURA, DIST_JINNEI = 1, 2, 3, 4, 5, 6
NJ_AUTO, NJ_FULLSCAN, NJ_PRUNED, NJ_CLUSTER = 0, 1, 2, 3
T_MSA_UPLOAD, T_MSA_DIST, T_NJ, T_SKETCH, T_MASH_DIST, T_PLACE = 0, 1, 2, 3, 4, 5


class Param:
    """MashPlacement::Param (src/mash_placement.cuh:16-32)."""

    def __init__(self, kmerSize=15, sketchSize=1000, threshold=1, distanceType=1, in_="r", out="t"):
        self.kmerSize, self.sketchSize, self.threshold = kmerSize, sketchSize, threshold
        self.distanceType, self.in_, self.out = distanceType, in_, out
        self.batchSize = 0
        self.backboneSize = 0


class Context:
    def __init__(self, device=0):
        h = C.c_void_p()
        check(lib().dipb_init(device, C.byref(h)))
        self.h = h
        self.device = device

    def elapsed_ms(self, what):
        return float(lib().dipb_elapsed_ms(self.h, what))

    def kernel_launches(self):
        return int(lib().dipb_kernel_launches(self.h))

    def sync(self):
        check(lib().dipb_sync(self.h))

    def nj_stats(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(lib().dipb_nj_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(rows_scanned=a.value, bytes_scanned=b.value, iterations=c.value)

    def close(self):
        if self.h:
            lib().dipb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Matrix:
    """Dense symmetric fp64 matrix on the device (NJDeviceArrays::d_mashDist)."""

    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle

    @classmethod
    def from_host(cls, ctx, D):
        D = np.ascontiguousarray(D, np.float64)
        h = C.c_void_p()
        check(lib().dipb_matrix_from_host(ctx.h, D, D.shape[0], 1, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_lower(cls, ctx, tri, n):
        tri = np.ascontiguousarray(tri, np.float64)
        if tri.size != n * (n - 1) // 2:
            raise ValueError("packed lower triangle must hold n(n-1)/2 values")
        h = C.c_void_p()
        check(lib().dipb_matrix_from_host(ctx.h, tri, n, 0, C.byref(h)))
        return cls(ctx, h)

    @property
    def n(self):
        return int(lib().dipb_matrix_n(self.h))

    def to_host(self):
        n = self.n
        out = np.empty((n, n), np.float64)
        check(lib().dipb_matrix_to_host(self.h, out))
        return out

    def free(self):
        if self.h:
            lib().dipb_matrix_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class MSADeviceArrays:
    """MashPlacement::MSADeviceArrays (src/mash_placement.cuh:89-99)."""

    def __init__(self, ctx):
        self.ctx, self.h, self.numSequences, self.seqLen = ctx, None, 0, 0

    def allocateDeviceArrays(self, h_compressedSeqs, h_seqLengths, num, params=None):
        """h_compressedSeqs: [num, ceil(L/16)] uint64 (or list of rows); h_seqLengths: [num]."""
        lens = np.ascontiguousarray(h_seqLengths, np.uint64)
        if len(lens) != num:
            raise ValueError("h_seqLengths must have num entries")
        h = C.c_void_p()
        if isinstance(h_compressedSeqs, np.ndarray) and h_compressedSeqs.ndim == 2:
            if np.any(lens != lens[0]):
                raise DipperError("aligned input requires equal sequence lengths")
            flat = np.ascontiguousarray(h_compressedSeqs, np.uint64)
            check(lib().dipb_msa_upload_flat(self.ctx.h, flat, num, int(lens[0]), C.byref(h)))
        else:
            rows = [np.ascontiguousarray(r, np.uint64) for r in h_compressedSeqs]
            ptrs = (C.c_void_p * num)(*[r.ctypes.data for r in rows])
            check(lib().dipb_msa_upload(self.ctx.h, ptrs, lens, num, C.byref(h)))
        self.h, self.numSequences, self.seqLen = h, num, int(lens[0])

    def distConstructionOnGpu(self, params, rowId):
        """d(rowId, j) for j < rowId (src/MSA.cu:271-282), returned as a host array."""
        out = np.zeros(max(rowId, 1), np.float64)
        check(lib().dipb_msa_dist_row_host(self.h, params.distanceType, rowId, out))
        return out[:rowId]

    def distBlock(self, params, r0, r1, ncols):
        """d(i, j) for i in [r0, r1), j in [0, ncols): the batched rows placement and the divide-and-conquer
        stages consume (DC/msa.cu:269-372).  Returned as a host array [r1 - r0, ncols]."""
        import torch
        ld = (ncols + 127) // 128 * 128
        buf = torch.empty((r1 - r0, ld), dtype=torch.float64, device=f"cuda:{self.ctx.device}")
        check(lib().dipb_msa_dist_block(self.h, params.distanceType, r0, r1, ncols, buf.data_ptr(), ld))
        self.ctx.sync()
        return buf[:, :ncols].cpu().numpy()

    def counts(self, i0, i1, j0, j1):
        m = np.zeros((i1 - i0, j1 - j0), np.int32)
        u = np.zeros((i1 - i0, j1 - j0), np.int32)
        check(lib().dipb_msa_counts(self.h, i0, i1, j0, j1, m, u))
        return m, u

    def distMatrix(self, params, row_begin=None, row_end=None):
        h = C.c_void_p()
        if row_begin is None:
            check(lib().dipb_msa_dist_matrix(self.h, params.distanceType, C.byref(h)))
        else:
            check(lib().dipb_msa_dist_matrix_rows(self.h, params.distanceType, row_begin, row_end, C.byref(h)))
        return Matrix(self.ctx, h)

    def dropOperands(self):
        """Release the tensor-core operand expansion (rebuilt by the next p / JC matrix or row block)."""
        check(lib().dipb_msa_drop_operands(self.h))

    def deallocateDeviceArrays(self):
        if self.h:
            lib().dipb_msa_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.deallocateDeviceArrays()
        except Exception:
            pass


class MashDeviceArrays:
    """MashPlacement::MashDeviceArrays (src/mash_placement.cuh:34-51)."""

    def __init__(self, ctx):
        self.ctx, self.h, self.numSequences, self.params = ctx, None, 0, None

    def allocateDeviceArrays(self, h_compressedSeqs, h_seqLengths, num, params):
        lens = np.ascontiguousarray(h_seqLengths, np.uint64)
        rows = [np.ascontiguousarray(r, np.uint64) for r in h_compressedSeqs]
        ptrs = (C.c_void_p * num)(*[r.ctypes.data for r in rows])
        h = C.c_void_p()
        check(lib().dipb_mash_upload(self.ctx.h, ptrs, lens, num, int(params.kmerSize), int(params.sketchSize), C.byref(h)))
        self.h, self.numSequences, self.params = h, num, params

    def setSketches(self, sketches, params):
        sk = np.ascontiguousarray(sketches, np.uint64)
        h = C.c_void_p()
        check(lib().dipb_mash_set_sketches(self.ctx.h, sk, sk.shape[0], int(params.kmerSize), sk.shape[1], C.byref(h)))
        self.h, self.numSequences, self.params = h, sk.shape[0], params

    def sketchConstructionOnGpu(self, params=None):
        check(lib().dipb_mash_sketch(self.h))

    def sketches(self):
        out = np.zeros((self.numSequences, int(self.params.sketchSize)), np.uint64)
        check(lib().dipb_mash_get_sketches(self.h, out))
        return out

    def distConstructionOnGpu(self, params, rowId):
        out = np.zeros(max(rowId, 1), np.float64)
        check(lib().dipb_mash_dist_row_host(self.h, rowId, out))
        return out[:rowId]

    def distMatrix(self):
        h = C.c_void_p()
        check(lib().dipb_mash_dist_matrix(self.h, C.byref(h)))
        return Matrix(self.ctx, h)

    def deallocateDeviceArrays(self):
        if self.h:
            lib().dipb_mash_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.deallocateDeviceArrays()
        except Exception:
            pass


class MatrixReader:
    """MashPlacement::MatrixReader (src/mash_placement.cuh:124-134, src/matrix_reader.cu):
    PHYLIP rows parsed one at a time, numbers through float32 like the reference's stof."""

    def __init__(self):
        self.numSequences, self.name, self.f = 0, [], None

    def allocateDeviceArrays(self, num, fptr):
        self.numSequences, self.name, self.f = num, [""] * num, fptr

    def distConstructionOnGpu(self, params, rowId):
        parts = self.f.readline().split()
        self.name[rowId] = parts[0]
        vals = np.array([np.float32(x) for x in parts[1:1 + rowId]], np.float64)
        if len(vals) < rowId:
            raise DipperError("PHYLIP row %d has %d distances, expected >= %d" % (rowId, len(vals), rowId))
        return vals

    @classmethod
    def read_all(cls, path):
        """Returns (names, packed lower triangle) of a PHYLIP file (full or lower-triangular)."""
        with open(path) as f:
            n = int(f.readline().split()[0])
            r = cls()
            r.allocateDeviceArrays(n, f)
            rows = [r.distConstructionOnGpu(None, i) for i in range(n)]
        tri = np.concatenate(rows) if n > 1 else np.zeros(0)
        return r.name, tri


class NJDeviceArrays:
    """MashPlacement::NJDeviceArrays (src/mash_placement.cuh:199-214)."""

    def __init__(self, ctx):
        self.ctx, self.matrix, self.d_numSequences = ctx, None, 0
        self.result = None

    def getDismatrix(self, numSequences, params, mashDeviceArrays=None, matrixReader=None, msaDeviceArrays=None):
        """src/neighborJoining.cu:35-85: build the dense matrix from the selected provider."""
        self.d_numSequences = numSequences
        if params.in_ == "m":
            self.matrix = msaDeviceArrays.distMatrix(params)
        elif params.in_ == "r":
            self.matrix = mashDeviceArrays.distMatrix()
        elif params.in_ == "d":
            rows = [matrixReader.distConstructionOnGpu(params, i) for i in range(numSequences)]
            tri = np.concatenate(rows) if numSequences > 1 else np.zeros(0)
            self.matrix = Matrix.from_lower(self.ctx, tri, numSequences)
        else:
            raise DipperError("unknown input format %r" % params.in_)

    def setMatrix(self, D):
        self.matrix = Matrix.from_host(self.ctx, D)
        self.d_numSequences = self.matrix.n

    def findNeighbourJoiningTree(self, name, algo=NJ_AUTO):
        """src/neighborJoining.cu:197-271; returns the Newick string the reference writes."""
        n = self.d_numSequences
        c0 = np.zeros(n - 1, np.int32)
        c1 = np.zeros(n - 1, np.int32)
        l0 = np.zeros(n - 1, np.float64)
        l1 = np.zeros(n - 1, np.float64)
        check(lib().dipb_nj(self.matrix.h, algo, c0, c1, l0, l1))
        self.result = (c0, c1, l0, l1)
        return _lib.take_str(lib().dipb_nj_newick(n, c0, c1, l0, l1, _lib.names_array(name)))

    def deallocateDeviceArrays(self):
        if self.matrix is not None:
            self.matrix.free()
            self.matrix = None


class KPlacementDeviceArrays:
    """MashPlacement::KPlacementDeviceArrays (src/mash_placement.cuh:167-197)."""

    def __init__(self, ctx):
        self.ctx, self.h, self.numSequences, self.backboneSize = ctx, None, 0, -1
        self._backbone = None

    def allocateDeviceArrays(self, num, backboneSize=-1):
        self.numSequences, self.backboneSize = num, backboneSize

    def initializeDeviceArrays(self, newick):
        """Reference takes a parsed Tree*; here the Newick string (src/tree.cpp:216-361)."""
        n = self.numSequences
        head = np.zeros(2 * n, np.int32)
        e, nxt, belong = (np.zeros(8 * n, np.int32) for _ in range(3))
        ln = np.zeros(8 * n, np.float64)
        names = C.c_void_p()
        B = lib().dipb_backbone_from_newick(newick.encode(), n, head, e, nxt, belong, ln, C.byref(names))
        if B < 0:
            check(B)
        self.backbone_names = _lib.take_str(names.value).split("\n")[:-1]
        self.backboneSize = B
        self._backbone = (head, e, nxt, belong, ln)
        return B

    def _source(self, params, mash, matrix, msa):
        s = DistSource()
        s.dist_type = int(params.distanceType)
        if params.in_ == "m":
            s.msa = msa.h
        elif params.in_ == "r":
            s.mash = mash.h
        else:
            s.matrix = matrix.h
        return s

    def findPlacementTree(self, params, mashDeviceArrays=None, matrix=None, msaDeviceArrays=None):
        s = self._source(params, mashDeviceArrays, matrix, msaDeviceArrays)
        h = C.c_void_p()
        check(lib().dipb_place_kclosest(self.ctx.h, C.byref(s), self.numSequences, C.byref(h)))
        self.h = h

    def addQuery(self, params, mashDeviceArrays=None, matrix=None, msaDeviceArrays=None):
        s = self._source(params, mashDeviceArrays, matrix, msaDeviceArrays)
        head, e, nxt, belong, ln = self._backbone
        h = C.c_void_p()
        check(lib().dipb_place_add(self.ctx.h, C.byref(s), self.numSequences, self.backboneSize, head, e, nxt, belong,
                                   ln, C.byref(h)))
        self.h = h

    def findTreeDC(self, params, backboneSize=None, mashDeviceArrays=None, matrix=None, msaDeviceArrays=None):
        """-m 3: findBackboneTreeDC + findClustersDC + findClusterTreeDC (DC/placement_close_k.cu:731-1535);
        backbone = numSequences / 20 unless given (src/tree_generation.cu:425,545)."""
        s = self._source(params, mashDeviceArrays, matrix, msaDeviceArrays)
        B = self.numSequences // 20 if backboneSize is None else backboneSize
        h = C.c_void_p()
        check(lib().dipb_dc(self.ctx.h, C.byref(s), self.numSequences, B, C.byref(h)))
        self.h = h
        cl = np.zeros(self.numSequences, np.int32)
        check(lib().dipb_dc_cluster_ids(self.ctx.h, cl, self.numSequences))
        self.clusterID = cl

    def findTreeDC_sharded(self, params, rank, world, all_gather, gather_to_root, backboneSize=None,
                           mashDeviceArrays=None, matrix=None, msaDeviceArrays=None, shard_clusters=True):
        """-m 3 with stage 2 (queries) and stage 3 (clusters) sharded over `world` ranks, one process
        per GPU.  `all_gather(np.int32 array)` must return the list of every rank's array;
        `gather_to_root(bytes)` must return the list of every rank's blob on rank 0 (None elsewhere).
        With torch.distributed these are all_gather_object / gather_object on a gloo group.
        Only rank 0 ends up holding the merged tree (self.h); other ranks get self.h = None.
        shard_clusters=False shards stage 2 only (93 % of the work at 2 M tips) and lets rank 0 place all clusters: the
        only exchange is then the all-gather of one int per tip instead of the slot slices of every cluster."""
        from . import sharding
        L = lib()
        s = self._source(params, mashDeviceArrays, matrix, msaDeviceArrays)
        n = self.numSequences
        B = n // 20 if backboneSize is None else backboneSize
        st = C.c_void_p()
        check(L.dipb_dc_begin(self.ctx.h, C.byref(s), n, B, C.byref(st)))
        q0, q1 = sharding.split_units(n - B, world)[rank]
        mine = np.zeros(max(q1 - q0, 1), np.int32)
        check(L.dipb_dc_assign(st, B + q0, B + q1, mine))
        parts = all_gather(mine[: q1 - q0])
        cl = np.concatenate([np.full(B, -1, np.int32)] + [np.asarray(p, np.int32) for p in parts])
        nc = C.c_int()
        check(L.dipb_dc_set_clusters(st, np.ascontiguousarray(cl), C.byref(nc)))
        sizes = np.zeros(max(nc.value, 1), np.int32)
        check(L.dipb_dc_cluster_sizes(st, sizes))
        self.clusterID = cl
        if not shard_clusters:
            if rank == 0:
                check(L.dipb_dc_run_clusters(st, 0, nc.value))
                h = C.c_void_p()
                check(L.dipb_dc_finish(st, C.byref(h)))
                self.h = h
            else:
                check(L.dipb_dc_finish(st, None))
                self.h = None
            return
        ranges = sharding.balance_clusters(sizes[: nc.value], world)
        c0, c1 = ranges[rank]
        check(L.dipb_dc_run_clusters(st, c0, c1))
        nbytes = C.c_size_t()
        check(L.dipb_dc_export_slice(st, c0, c1, None, 0, C.byref(nbytes)))
        buf = (C.c_char * max(nbytes.value, 1))()
        check(L.dipb_dc_export_slice(st, c0, c1, buf, nbytes.value, C.byref(nbytes)))
        blobs = gather_to_root(bytes(buf[: nbytes.value]))
        self.clusterID = cl
        if rank == 0:
            for r, blob in enumerate(blobs):
                if r == 0:
                    continue
                check(L.dipb_dc_import_slice(st, blob, len(blob)))
            h = C.c_void_p()
            check(L.dipb_dc_finish(st, C.byref(h)))
            self.h = h
        else:
            check(L.dipb_dc_finish(st, None))
            self.h = None

    def export(self):
        n = self.numSequences
        head = np.zeros(2 * n, np.int32)
        e, nxt, belong = (np.zeros(8 * n, np.int32) for _ in range(3))
        ln = np.zeros(8 * n, np.float64)
        check(lib().dipb_tree_export(self.h, head, e, nxt, belong, ln))
        return dict(head=head, e=e, nxt=nxt, belong=belong, len=ln)

    def export_closest(self):
        n = self.numSequences
        cid = np.zeros(40 * n, np.int32)
        cdis = np.zeros(40 * n, np.float64)
        check(lib().dipb_tree_export_closest(self.h, cid, cdis))
        return cid, cdis

    def printTree(self, name):
        a = self.export()
        n = self.numSequences
        names = list(name) + [""] * (2 * n - len(name))
        return _lib.take_str(lib().dipb_tree_newick(2 * n, n, a["head"], a["e"], a["nxt"], a["len"],
                                                    _lib.names_array(names)))

    def deallocateDeviceArrays(self):
        if self.h:
            lib().dipb_tree_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.deallocateDeviceArrays()
        except Exception:
            pass


class PlacementDeviceArrays(KPlacementDeviceArrays):
    """MashPlacement::PlacementDeviceArrays, the exact placement mode (src/mash_placement.cuh:137-165,
    src/placement.cu:505-789); printTree starts at node numSequences (src/placement.cu:500) like the k-closest one."""

    def findPlacementTree(self, params, mashDeviceArrays=None, matrix=None, msaDeviceArrays=None):
        s = self._source(params, mashDeviceArrays, matrix, msaDeviceArrays)
        h = C.c_void_p()
        check(lib().dipb_place_exact(self.ctx.h, C.byref(s), self.numSequences, C.byref(h)))
        self.h = h

    @staticmethod
    def maxTips():
        return lib().dipb_place_exact_max_tips()


def read_fasta_packed(path, bits, threads=0):
    """Parallel FASTA ingest + packing (dipb_fasta_open): returns (names, lengths[n], word_offsets[n+1], words).
    bits = 4 for aligned input (-i m), 2 for unaligned (-i r).  Replaces readSequences + the packing loops
    (src/tree_generation.cu:132-154, 350-362, 478-490)."""
    h = C.c_void_p()
    check(lib().dipb_fasta_open(os.fsencode(path), bits, threads, C.byref(h)))
    try:
        L = lib()
        n = L.dipb_fasta_count(h)
        names = [L.dipb_fasta_name(h, i).decode() for i in range(n)]
        off = np.ctypeslib.as_array(L.dipb_fasta_word_offsets(h), shape=(n + 1,)).copy()
        lens = np.ctypeslib.as_array(L.dipb_fasta_lengths(h), shape=(n,)).copy() if n else np.zeros(0, np.uint64)
        nw = int(off[-1])
        words = np.ctypeslib.as_array(L.dipb_fasta_words(h), shape=(nw,)).copy() if nw else np.zeros(0, np.uint64)
    finally:
        lib().dipb_fasta_close(h)
    return names, lens, off, words


def pack4(seq):
    b = seq if isinstance(seq, bytes) else seq.encode()
    out = np.zeros((len(b) + 15) // 16, np.uint64)
    lib().dipb_pack4(b, len(b), out)
    return out


def pack2(seq):
    b = seq if isinstance(seq, bytes) else seq.encode()
    out = np.zeros((len(b) + 31) // 32, np.uint64)
    lib().dipb_pack2(b, len(b), out)
    return out


class MultiDevice:
    """Several GPUs of one box in ONE process (dipb_multi_*, csrc/multi.cu): one context and one host thread per device,
    shards exchanged by peer copies.  The reference is single GPU (src/tree_generation.cu:240)."""

    def __init__(self, devices):
        devs = np.ascontiguousarray(devices, np.int32)
        h = C.c_void_p()
        check(lib().dipb_multi_init(devs, len(devs), C.byref(h)))
        self.h, self.devices, self.n = h, list(devices), 0

    def context(self, d=0):
        c = Context.__new__(Context)
        c.h, c.device = C.c_void_p(lib().dipb_multi_ctx(self.h, d)), self.devices[d]
        c.close = lambda: None          # owned by the multi-device handle
        return c

    def allocateDeviceArrays(self, flat, seq_len):
        flat = np.ascontiguousarray(flat, np.uint64)
        self.n = flat.shape[0]
        check(lib().dipb_multi_msa_upload_flat(self.h, flat, self.n, int(seq_len)))

    def distMatrix(self, params):
        h = C.c_void_p()
        check(lib().dipb_multi_msa_dist_matrix(self.h, params.distanceType, C.byref(h)))
        return Matrix(self.context(0), h)

    def findTreeDC(self, params, backboneSize=None):
        """-m 3 with stage 2 split over the devices; returns a KPlacementDeviceArrays holding the tree (device 0)."""
        B = self.n // 20 if backboneSize is None else backboneSize
        h = C.c_void_p()
        check(lib().dipb_multi_dc(self.h, params.distanceType, B, C.byref(h)))
        kp = KPlacementDeviceArrays(self.context(0))
        kp.allocateDeviceArrays(self.n)
        kp.h = h
        cl = np.zeros(self.n, np.int32)
        check(lib().dipb_multi_dc_cluster_ids(self.h, cl, self.n))
        kp.clusterID = cl
        return kp

    def elapsed_ms(self, what):
        return float(lib().dipb_multi_elapsed_ms(self.h, what))

    def close(self):
        if self.h:
            lib().dipb_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
