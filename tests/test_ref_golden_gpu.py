"""CUDA path vs the frozen outputs of the reference's OWN objects (tests/golden/ref_*.npz, written on a B200 by
tools/make_ref_golden.py): distance models 3-6, divide and conquer (aligned and Mash), add-tips incl. BASELINE config 4b
on dataset/t2.backbone.nwk.  The live twins of these checks are in test_ref_parity.py."""
import glob
import os

import numpy as np
import pytest

from dipper_b200 import api, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(kind):
    out = []
    for fn in sorted(glob.glob(os.path.join(GOLD, "ref_*.npz"))):
        z = np.load(fn)
        if str(z["kind"]) == kind:
            out.append((fn, z))
    assert out, "no committed fixture of kind " + kind
    return out


def _same_tree_arrays(kp, z, n):
    a = kp.export()
    ns = 4 * n - 4
    assert np.array_equal(a["head"][: 2 * n], z["head"][: 2 * n])
    for k in ("e", "nxt", "belong"):
        assert np.array_equal(a[k][:ns], z[k][:ns]), k
    assert np.allclose(a["len"][:ns], z["len"][:ns], rtol=0, atol=1e-12)


def test_distance_models_vs_reference_fixture(ctx):
    for fn, z in _load("ref_models"):
        P, L = z["packed"], int(z["seq_len"])
        n = P.shape[0]
        msa = api.MSADeviceArrays(ctx)
        msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, api.Param(in_="m"))
        low = np.tril_indices(n, -1)
        for t in range(1, 7):
            D = msa.distMatrix(api.Param(distanceType=t, in_="m")).to_host()
            assert np.allclose(D[low], z["rows_%d" % t][low], rtol=1e-6, atol=0), (fn, t)


def test_dc_aligned_vs_reference_fixture(ctx, monkeypatch):
    for fn, z in _load("ref_dc_msa"):
        P, L, B = z["packed"], int(z["seq_len"]), int(z["backbone"])
        n = P.shape[0]
        prm = api.Param(distanceType=2, in_="m")
        msa = api.MSADeviceArrays(ctx)
        msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
        monkeypatch.setenv("DIPB_DC_REF_B17", "1")     # the reference as shipped (defect B17), see dc.cu
        kp = api.KPlacementDeviceArrays(ctx)
        kp.allocateDeviceArrays(n)
        kp.findTreeDC(prm, backboneSize=B, msaDeviceArrays=msa)
        assert np.array_equal(kp.clusterID, z["clusters"]), fn
        _same_tree_arrays(kp, z, n)
        assert kp.printTree(synth.names(n)) == str(z["newick"]), fn
        monkeypatch.delenv("DIPB_DC_REF_B17")


def test_dc_mash_vs_reference_fixture(ctx):
    for fn, z in _load("ref_dc_mash"):
        lens, offs, flat, B = z["lens"], z["offsets"], z["flat"], int(z["backbone"])
        n = len(lens)
        prm = api.Param(kmerSize=int(z["k"]), sketchSize=int(z["s"]), in_="r")
        m = api.MashDeviceArrays(ctx)
        bounds = list(offs.astype(np.int64)) + [len(flat)]
        m.allocateDeviceArrays([flat[bounds[i]:bounds[i + 1]] for i in range(n)], lens, n, prm)
        m.sketchConstructionOnGpu()
        kp = api.KPlacementDeviceArrays(ctx)
        kp.allocateDeviceArrays(n)
        kp.findTreeDC(prm, backboneSize=B, mashDeviceArrays=m)
        assert np.array_equal(kp.clusterID, z["clusters"]), fn
        _same_tree_arrays(kp, z, n)
        assert kp.printTree(synth.names(n)) == str(z["newick"]), fn


@pytest.mark.parametrize("kind", ["ref_add_msa", "ref_add_t2"])
def test_add_tips_vs_reference_fixture(ctx, kind):
    for fn, z in _load(kind):
        P, L, B = z["packed"], int(z["seq_len"]), int(z["backbone"])
        n = P.shape[0]
        bb = str(z["backbone_newick"]) if "backbone_newick" in z else open(os.path.join(GOLD, "t2.backbone.nwk")).readline().strip()
        prm = api.Param(distanceType=2, in_="m")
        msa = api.MSADeviceArrays(ctx)
        msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
        kp = api.KPlacementDeviceArrays(ctx)
        kp.allocateDeviceArrays(n)
        assert kp.initializeDeviceArrays(bb) == B
        kp.addQuery(prm, msaDeviceArrays=msa)
        _same_tree_arrays(kp, z, n)
        names = kp.backbone_names + ["Q%d" % (i + 1) for i in range(n - B)]
        assert kp.printTree(names) == str(z["newick"]), fn
