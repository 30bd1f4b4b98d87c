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


def _read_arrays(path, n):
    raw = np.fromfile(path, np.uint8)
    o = 0
    out = {}
    for k, cnt, dt in (("head", 2 * n, np.int32), ("e", 8 * n, np.int32), ("nxt", 8 * n, np.int32), ("belong", 8 * n, np.int32),
                       ("len", 8 * n, np.float64), ("cid", 20 * n, np.int32), ("cdis", 20 * n, np.float64)):
        nb = cnt * np.dtype(dt).itemsize
        if o + nb > raw.size:
            break                      # closest lists are only dumped by the D&C modes
        out[k] = raw[o:o + nb].view(dt).copy()
        o += nb
    return out


def _stage3_seed(ref, B):
    """Reference defect B10 in its D&C twin: the stage-3 closest-leaf BFS reads dis[0] / from[0] of queue arrays that
    findClusterTreeDC cudaMalloc's and never initialises (src/divide_and_conquer/placement_close_k.cu:326-331,
    1261-1275).  Zero-filled memory gives the intended result (seed 0); recycled memory adds the stale value to every
    closest-list distance of stage 3.  The value is visible in the reference's own output: the first placed tip's
    slot tip->middle (index 4B-4+2) lists the tip itself at that distance."""
    return float(ref["cdis"][(4 * B - 4 + 2) * 5])


def _same_tree_arrays(kp, ref, n):
    _same_arrays(kp.export(), ref, n)


def _same_arrays(a, ref, n):
    ns = 4 * n - 4
    assert np.array_equal(a["head"][: 2 * n], ref["head"][: 2 * n])
    for k in ("e", "nxt", "belong"):
        assert np.array_equal(a[k][:ns], ref[k][:ns]), k
    assert np.allclose(a["len"][:ns], ref["len"][:ns], rtol=0, atol=1e-12)


def _singleton_at_B(cl, B):
    """see tools/make_ref_golden.py: keeps reference defect B12 (tip id == B read from the wrong buffer) without effect"""
    cnt = np.bincount(cl[B:], minlength=int(cl.max()) + 1)
    return next(q for q in range(B, len(cl)) if cnt[cl[q]] == 1)


@pytest.mark.parametrize("dist_type", [3, 4, 5, 6])
def test_msa_rows_models_3_to_6_vs_reference_cuda(ctx, oracle, tmp_path, dist_type):
    """TN / K2P / Tamura / Jin-Nei: rows of the reference's well-formed DC twins (src/divide_and_conquer/msa.cu:219-264;
    src/MSA.cu:239-265 indexes out of bounds) vs our bit-plane kernel and the oracle.  Tolerance 1e-6 relative
    (north star): nvcc may contract the Tajima-Nei / Tamura products into FMAs differently."""
    n, L = 150, 3000
    codes, P, _ = make_msa(n, L, seed=48, gap_cols=0.05)
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "o")
    write_bin(inp, P, [L] * n, 4)
    run_ref("msa_dc_rows", inp, out, dist_type)
    ref = np.fromfile(out + ".rows", np.float64)
    msa = api.MSADeviceArrays(ctx)
    msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, api.Param(in_="m"))
    D = msa.distMatrix(api.Param(distanceType=dist_type, in_="m")).to_host()
    mine = np.concatenate([D[i, :i] for i in range(1, n)])
    ok = np.isfinite(ref)
    assert ok.mean() > 0.99
    assert np.allclose(mine[ok], ref[ok], rtol=1e-6, atol=0)
    orc = oracle.msa_dist_matrix(P, L, dist_type)
    assert np.allclose(np.concatenate([orc[i, :i] for i in range(1, n)])[ok], ref[ok], rtol=1e-6, atol=0)


def test_dc_aligned_vs_reference_cuda(ctx, oracle, tmp_path, monkeypatch):
    """-m 3 on aligned input vs findBackboneTreeDC / findClustersDC / findClusterTreeDC of the reference's own objects
    (src/divide_and_conquer/placement_close_k.cu:731-1535).  The reference as shipped carries defect B17 (stale
    d(query, tip B-1) in the assignment stage); DIPB_DC_REF_B17=1 switches the same behaviour on here so that all
    cluster ids and slot arrays can be compared exactly; the default (intended) rule is checked against the oracle."""
    # (the reference exits when a cluster reaches the backbone size, src/divide_and_conquer/placement_close_k.cu:1334-1337;
    # with its stale distance to tip B-1 that tip attracts many queries, so the backbone is taken large here)
    n, L, B = 2000, 2000, 400
    codes, P, _ = make_msa(n, L, seed=61)
    D0 = oracle.msa_dist_matrix(P, L, 2)
    _, cl0 = oracle.dc_as_shipped(D0, B, 0.0)
    q = _singleton_at_B(cl0, B)
    P[[B, q]] = P[[q, B]]
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "o")
    write_bin(inp, P, [L] * n, 4)
    run_ref("msa_dc", inp, out, 2, 15, B)
    ref_cl = np.fromfile(out + ".clusters", np.int32)
    ref = _read_arrays(out + ".arrays", n)
    prm = api.Param(distanceType=2, in_="m")
    msa = api.MSADeviceArrays(ctx)
    msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
    monkeypatch.setenv("DIPB_DC_REF_B17", "1")
    kp = api.KPlacementDeviceArrays(ctx)
    kp.allocateDeviceArrays(n)
    kp.findTreeDC(prm, backboneSize=B, msaDeviceArrays=msa)
    assert np.array_equal(kp.clusterID, ref_cl)
    # the oracle with the reference's defects on reproduces the reference run (pins orc_dc): B17 and, when this run's
    # stage-3 queue memory was not zero-filled, the stale BFS seed B10 (see _stage3_seed)
    D = msa.distMatrix(prm).to_host()
    seed = _stage3_seed(ref, B)
    ot, ocl = oracle.dc_as_shipped(D, B, 0.0, stage3_seed=seed)
    assert np.array_equal(ocl, ref_cl)
    _same_arrays(ot.arrays(), ref, n)
    assert ot.newick(synth.names(n)) == open(out + ".nwk").read()
    if seed == 0.0:      # memory was clean: the product (B17 switch on) must equal the reference slot for slot
        _same_tree_arrays(kp, ref, n)
        assert kp.printTree(synth.names(n)) == open(out + ".nwk").read()
    ot0, _ = oracle.dc_as_shipped(D, B, 0.0)
    _same_tree_arrays(kp, ot0.arrays(), n)
    # default = intended rule: equal to the oracle's intended rule, a few queries near tip B-1 differ from the reference
    monkeypatch.delenv("DIPB_DC_REF_B17")
    kp2 = api.KPlacementDeviceArrays(ctx)
    kp2.allocateDeviceArrays(n)
    kp2.findTreeDC(prm, backboneSize=B, msaDeviceArrays=msa)
    _, icl = oracle.dc(D, B)
    assert np.array_equal(kp2.clusterID, icl)
    assert np.count_nonzero(icl != ref_cl) < 0.05 * (n - B)


def test_dc_mash_vs_reference_cuda(ctx, oracle, tmp_path):
    """-m 3 on unaligned input vs the reference's own DC objects (src/divide_and_conquer/mash.cu + placement_close_k.cu)."""
    n, B = 300, 60
    codes, _ = synth.evolve(n, 3000, seed=62, regime="tiefree", gap_cols=0.01)
    seqs = synth.unaligned(codes)
    flat, offs, lens = synth.flatten2(seqs)
    _, cl0 = oracle.dc(oracle.mash_dist_matrix(oracle.sketch_all(flat, offs, lens, 15, 1000), 15), B)
    q = _singleton_at_B(cl0, B)
    seqs[B], seqs[q] = seqs[q], seqs[B]
    packed = [synth.pack2_np(s) for s in seqs]
    lens = np.array([len(s) for s in seqs], np.uint64)
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "o")
    write_bin(inp, packed, lens, 2)
    run_ref("mash_dc", inp, out, 1, 15, B)
    ref_cl = np.fromfile(out + ".clusters", np.int32)
    ref = _read_arrays(out + ".arrays", n)
    prm = api.Param(kmerSize=15, sketchSize=1000, in_="r")
    m = api.MashDeviceArrays(ctx)
    m.allocateDeviceArrays(packed, lens, n, prm)
    m.sketchConstructionOnGpu()
    kp = api.KPlacementDeviceArrays(ctx)
    kp.allocateDeviceArrays(n)
    kp.findTreeDC(prm, backboneSize=B, mashDeviceArrays=m)
    assert np.array_equal(kp.clusterID, ref_cl)
    D = m.distMatrix().to_host()
    seed = _stage3_seed(ref, B)
    ot, ocl = oracle.dc(D, B, stage3_seed=seed)          # the reference run incl. its stale BFS seed, if any
    assert np.array_equal(ocl, ref_cl)
    _same_arrays(ot.arrays(), ref, n)
    assert ot.newick(synth.names(n)) == open(out + ".nwk").read()
    if seed == 0.0:
        _same_tree_arrays(kp, ref, n)
        assert kp.printTree(synth.names(n)) == open(out + ".nwk").read()
    ot0, _ = oracle.dc(D, B)
    _same_tree_arrays(kp, ot0.arrays(), n)
    assert kp.printTree(synth.names(n)) == ot0.newick(synth.names(n))


def test_add_tips_onto_t2_backbone_vs_reference_cuda(ctx, oracle, tmp_path):
    """BASELINE config 4b: -m 1 --add onto the reference's own fixture dataset/t2.backbone.nwk (1000 tips; copy under
    tests/golden/), 2000 queries evolved on a tree that contains it; vs initializeDeviceArrays(Tree*) + addQuery of the
    reference's objects (src/placement_close_k.cu:126-264,858-990)."""
    bbt = open(os.path.join(ROOT, "tests", "golden", "t2.backbone.nwk")).readline().strip()
    nq, L = 2000, 3000
    tree, bb_leaves, bb_names, qnodes = synth.tree_with_queries(bbt, nq, seed=63, scale=50.0)
    B = len(bb_leaves)
    n = B + nq
    codes, info = synth.evolve(n, L, seed=64, tree=tree, gap_cols=0.02)
    row_of = {int(v): i for i, v in enumerate(info["order"])}
    order = [row_of[v] for v in bb_leaves] + [row_of[v] for v in qnodes]
    P = np.ascontiguousarray(synth.pack4_np(codes[order]))
    inp, out, nwk = str(tmp_path / "in.bin"), str(tmp_path / "o"), str(tmp_path / "bb.nwk")
    write_bin(inp, P, [L] * n, 4)
    open(nwk, "w").write(bbt + "\n")
    run_ref("msa_add", inp, out, 2, 15, nwk)
    ref = _read_arrays(out + ".arrays", n)
    prm = api.Param(distanceType=2, in_="m")
    msa = api.MSADeviceArrays(ctx)
    msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
    kp = api.KPlacementDeviceArrays(ctx)
    kp.allocateDeviceArrays(n)
    assert kp.initializeDeviceArrays(bbt) == B and kp.backbone_names == bb_names
    kp.addQuery(prm, msaDeviceArrays=msa)
    _same_tree_arrays(kp, ref, n)
    names = bb_names + ["Q%d" % (i + 1) for i in range(nq)]
    assert kp.printTree(names) == open(out + ".nwk").read()
