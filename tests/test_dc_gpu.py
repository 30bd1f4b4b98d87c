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
