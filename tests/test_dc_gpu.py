"""GPU parity: divide-and-conquer (-m 3) through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

from dipper_b200 import api, newick, synth
from conftest import make_msa
from test_placement_gpu import compare_trees

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,B", [(40, 4), (200, 10), (600, 30), (900, 300)])
def test_dc_from_matrix_matches_oracle(ctx, oracle, n, B):
    codes, P, _ = make_msa(n, 900, seed=300 + n)
    D = oracle.msa_dist_matrix(P, 900, 2)
    M = api.Matrix.from_host(ctx, D)
    kp = api.KPlacementDeviceArrays(ctx)
    kp.allocateDeviceArrays(n)
    kp.findTreeDC(api.Param(in_="d"), backboneSize=B, matrix=M)
    ot, ocl = oracle.dc(D, B)
    assert np.array_equal(kp.clusterID, ocl)
    compare_trees(kp, ot, n, 4 * n - 4)
    assert kp.printTree(synth.names(n)) == ot.newick(synth.names(n))


def test_dc_from_msa_default_backbone(ctx, oracle):
    n, L = 800, 1200
    codes, P, _ = make_msa(n, L, seed=51)
    prm = api.Param(distanceType=2, in_="m")
    msa = api.MSADeviceArrays(ctx)
    msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
    kp = api.KPlacementDeviceArrays(ctx)
    kp.allocateDeviceArrays(n)
    kp.findTreeDC(prm, msaDeviceArrays=msa)          # backbone = n / 20
    D = msa.distMatrix(prm).to_host()
    ot, ocl = oracle.dc(D, n // 20)
    assert np.array_equal(kp.clusterID, ocl)
    compare_trees(kp, ot, n, 4 * n - 4)
    nwk = kp.printTree(synth.names(n))
    assert nwk.count("T") == n and newick.rf_distance(nwk, ot.newick(synth.names(n))) == 0


def test_dc_from_mash(ctx, oracle):
    n = 240
    codes, _ = synth.evolve(n, 3000, seed=52, gap_cols=0.01)
    seqs = synth.unaligned(codes)
    prm = api.Param(kmerSize=15, sketchSize=1000, in_="r")
    m = api.MashDeviceArrays(ctx)
    m.allocateDeviceArrays([synth.pack2_np(s) for s in seqs], np.array([len(s) for s in seqs], np.uint64), n, prm)
    m.sketchConstructionOnGpu()
    kp = api.KPlacementDeviceArrays(ctx)
    kp.allocateDeviceArrays(n)
    kp.findTreeDC(prm, backboneSize=24, mashDeviceArrays=m)
    D = m.distMatrix().to_host()
    ot, ocl = oracle.dc(D, 24)
    assert np.array_equal(kp.clusterID, ocl)
    compare_trees(kp, ot, n, 4 * n - 4)


@pytest.mark.parametrize("dist_type", [4, 5])
def test_dc_other_models_in_cluster_distances(ctx, oracle, dist_type):
    n, L = 160, 1500
    codes, P, _ = make_msa(n, L, seed=53)
    prm = api.Param(distanceType=dist_type, in_="m")
    msa = api.MSADeviceArrays(ctx)
    msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
    kp = api.KPlacementDeviceArrays(ctx)
    kp.allocateDeviceArrays(n)
    kp.findTreeDC(prm, backboneSize=16, msaDeviceArrays=msa)
    D = msa.distMatrix(prm).to_host()
    ot, ocl = oracle.dc(D, 16)
    assert np.array_equal(kp.clusterID, ocl)
    assert newick.rf_distance(kp.printTree(synth.names(n)), ot.newick(synth.names(n))) == 0


def test_dc_sharded_over_ranks_equals_single_gpu(ctx, oracle):
    """Stage 2 / stage 3 sharding emulated on ONE GPU: every 'rank' runs its share of the
    queries and clusters on its own state, rank 0 imports the other slices -> same tree as the oracle."""
    import ctypes as C
    from dipper_b200 import sharding
    from dipper_b200._lib import lib, check
    n, B, world = 700, 35, 3
    codes, P, _ = make_msa(n, 900, seed=71)
    D = oracle.msa_dist_matrix(P, 900, 2)
    M = api.Matrix.from_host(ctx, D)
    prm = api.Param(in_="d")
    kps = [api.KPlacementDeviceArrays(ctx) for _ in range(world)]
    for k in kps:
        k.allocateDeviceArrays(n)
    L = lib()
    # drive the staged C ABI by hand, interleaving the "ranks"
    states, parts = [], []
    for r in range(world):
        s = kps[r]._source(prm, None, M, None)
        st = C.c_void_p()
        check(L.dipb_dc_begin(ctx.h, C.byref(s), n, B, C.byref(st)))
        states.append(st)
        q0, q1 = sharding.split_units(n - B, world)[r]
        mine = np.zeros(q1 - q0, np.int32)
        check(L.dipb_dc_assign(st, B + q0, B + q1, mine))
        parts.append(mine)
    cl = np.concatenate([np.full(B, -1, np.int32)] + parts)
    blobs = []
    for r in range(world):
        nc = C.c_int()
        check(L.dipb_dc_set_clusters(states[r], cl, C.byref(nc)))
        sizes = np.zeros(nc.value, np.int32)
        check(L.dipb_dc_cluster_sizes(states[r], sizes))
        c0, c1 = sharding.balance_clusters(sizes, world)[r]
        check(L.dipb_dc_run_clusters(states[r], c0, c1))
        nb = C.c_size_t()
        check(L.dipb_dc_export_slice(states[r], c0, c1, None, 0, C.byref(nb)))
        buf = (C.c_char * nb.value)()
        check(L.dipb_dc_export_slice(states[r], c0, c1, buf, nb.value, C.byref(nb)))
        blobs.append(bytes(buf))
    for r in range(1, world):
        check(L.dipb_dc_import_slice(states[0], blobs[r], len(blobs[r])))
        check(L.dipb_dc_finish(states[r], None))
    h = C.c_void_p()
    check(L.dipb_dc_finish(states[0], C.byref(h)))
    kps[0].h = h
    ot, ocl = oracle.dc(D, B)
    assert np.array_equal(cl, ocl)
    compare_trees(kps[0], ot, n, 4 * n - 4)
    assert kps[0].printTree(synth.names(n)) == ot.newick(synth.names(n))
