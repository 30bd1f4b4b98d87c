"""GPU parity against the reference's OWN CUDA objects (oracle/_ref/dipper_ref, built by
oracle/build_ref.sh from /root/reference in place; the binary travels to the GPU box)."""
import json
import os
import subprocess

import numpy as np
import pytest

from dipper_b200 import api, newick, synth
from conftest import make_msa, ROOT

pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, "oracle", "_ref", "dipper_ref")


def write_bin(path, rows, lens, bits):
    with open(path, "wb") as f:
        np.array([len(lens), bits], np.int64).tofile(f)
        np.asarray(lens, np.uint64).tofile(f)
        for r in rows:
            np.ascontiguousarray(r, np.uint64).tofile(f)


def run_ref(mode, inp, out, *extra):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/dipper_ref not built (reference not mounted at build time)")
    p = subprocess.run([REF, mode, inp, out, *map(str, extra)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return json.loads(p.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("dist_type", [1, 2])
def test_msa_rows_vs_reference_cuda(ctx, oracle, tmp_path, dist_type):
    n, L = 200, 3000
    codes, P, _ = make_msa(n, L, seed=41, gap_cols=0.05)
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "o")
    write_bin(inp, P, [L] * n, 4)
    run_ref("msa_rows", inp, out, dist_type)
    ref = np.fromfile(out + ".rows", np.float64)
    msa = api.MSADeviceArrays(ctx)
    msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, api.Param(in_="m"))
    D = msa.distMatrix(api.Param(distanceType=dist_type, in_="m")).to_host()
    mine = np.concatenate([D[i, :i] for i in range(1, n)])
    assert np.array_equal(mine, ref)        # same counts, same fp64 expression, same libdevice log
    orc = oracle.msa_dist_matrix(P, L, dist_type)
    assert np.allclose(np.concatenate([orc[i, :i] for i in range(1, n)]), ref, rtol=1e-6, atol=0)


def test_msa_nj_tree_vs_reference_cuda(ctx, oracle, tmp_path):
    n, L = 400, 4000
    codes, P, _ = make_msa(n, L, seed=42)
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "o")
    write_bin(inp, P, [L] * n, 4)
    run_ref("msa_nj", inp, out, 2)
    ref_nwk = open(out + ".nwk").read()
    prm = api.Param(distanceType=2, in_="m")
    msa = api.MSADeviceArrays(ctx)
    msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
    nj = api.NJDeviceArrays(ctx)
    nj.getDismatrix(n, prm, msaDeviceArrays=msa)
    nwk = nj.findNeighbourJoiningTree(synth.names(n))
    assert newick.rf_distance(nwk, ref_nwk) == 0
    assert newick.max_branch_diff(nwk, ref_nwk) < 1e-5
    o = oracle.nj(oracle.msa_dist_matrix(P, L, 2))
    assert newick.rf_distance(oracle.nj_newick(*o, synth.names(n)), ref_nwk) == 0


def test_mash_sketches_and_rows_vs_reference_cuda(ctx, oracle, tmp_path):
    codes, _ = synth.evolve(64, 3000, seed=43, regime="tiefree", gap_cols=0.02)
    seqs = synth.unaligned(codes)
    packed = [synth.pack2_np(s) for s in seqs]
    lens = np.array([len(s) for s in seqs], np.uint64)
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "o")
    write_bin(inp, packed, lens, 2)
    run_ref("mash_sketch", inp, out, 2, 15)
    ref_sk = np.fromfile(out + ".sk", np.uint64).reshape(64, 1000)
    m = api.MashDeviceArrays(ctx)
    m.allocateDeviceArrays(packed, lens, 64, api.Param(kmerSize=15, sketchSize=1000, in_="r"))
    m.sketchConstructionOnGpu()
    assert np.array_equal(m.sketches(), ref_sk)                      # bit-exact vs the reference's kernel
    flat, offs, ln = synth.flatten2(seqs)
    assert np.array_equal(oracle.sketch_all(flat, offs, ln, 15, 1000), ref_sk)   # pins the oracle too
    run_ref("mash_rows", inp, out, 2, 15)
    ref_rows = np.fromfile(out + ".rows", np.float64)
    D = m.distMatrix().to_host()
    mine = np.concatenate([D[i, :i] for i in range(1, 64)])
    assert np.array_equal(mine, ref_rows)


def test_msa_placement_tree_vs_reference_cuda(ctx, oracle, tmp_path):
    """-m 1 k-closest placement: our tree vs the reference's own placement kernels."""
    n, L = 300, 3000
    codes, P, _ = make_msa(n, L, seed=44)
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "o")
    write_bin(inp, P, [L] * n, 4)
    run_ref("msa_place", inp, out, 2)
    ref_nwk = open(out + ".nwk").read()
    prm = api.Param(distanceType=2, in_="m")
    msa = api.MSADeviceArrays(ctx)
    msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
    kp = api.KPlacementDeviceArrays(ctx)
    kp.allocateDeviceArrays(n)
    kp.findPlacementTree(prm, msaDeviceArrays=msa)
    nwk = kp.printTree(synth.names(n))
    assert newick.rf_distance(nwk, ref_nwk) == 0
    assert newick.max_branch_diff(nwk, ref_nwk) < 1e-5
    assert nwk == ref_nwk      # same slots, same adjacency order, same %g text


def test_msa_exact_placement_tree_vs_reference_cuda(ctx, oracle, tmp_path):
    """-p 0 exact placement (src/placement.cu): our tree and the oracle's vs the reference's own kernels."""
    n, L = 500, 3000
    codes, P, _ = make_msa(n, L, seed=47)
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "o")
    write_bin(inp, P, [L] * n, 4)
    run_ref("msa_place_exact", inp, out, 2)
    ref_nwk = open(out + ".nwk").read()
    prm = api.Param(distanceType=2, in_="m")
    msa = api.MSADeviceArrays(ctx)
    msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
    pl = api.PlacementDeviceArrays(ctx)
    pl.allocateDeviceArrays(n)
    pl.findPlacementTree(prm, msaDeviceArrays=msa)
    nwk = pl.printTree(synth.names(n))
    assert newick.rf_distance(nwk, ref_nwk) == 0
    assert newick.max_branch_diff(nwk, ref_nwk) < 1e-5
    assert nwk == ref_nwk      # same slots, same adjacency order, same %g text
    D = msa.distMatrix(prm).to_host()
    assert oracle.place_exact(D).newick(synth.names(n)) == ref_nwk   # pins the oracle restatement
