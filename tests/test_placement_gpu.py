"""GPU parity: k-closest placement and add-tips through the C ABI vs the CPU oracle.
Tree arrays (slots, lengths, closest lists) must be identical, hence the Newick too."""
import numpy as np
import pytest

from dipper_b200 import api, newick, synth
from conftest import make_msa

pytestmark = pytest.mark.gpu


def compare_trees(kp, otree, n, live_slots):
    mine = kp.export()
    ref = otree.arrays()
    for k in ("head", "e", "nxt", "belong"):
        assert np.array_equal(mine[k], ref[k]), k
    assert np.array_equal(mine["len"][:live_slots], ref["len"][:live_slots])
    cid, cdis = kp.export_closest()
    assert np.array_equal(cid[: 5 * live_slots], ref["cid"][: 5 * live_slots])
    assert np.array_equal(cdis[: 5 * live_slots], ref["cdis"][: 5 * live_slots])


@pytest.mark.parametrize("n", [2, 3, 4, 10, 129, 700])
def test_placement_from_matrix_matches_oracle(ctx, oracle, n):
    codes, P, _ = make_msa(n, 800, seed=200 + n)
    D = oracle.msa_dist_matrix(P, 800, 2)
    M = api.Matrix.from_host(ctx, D)
    kp = api.KPlacementDeviceArrays(ctx)
    kp.allocateDeviceArrays(n)
    kp.findPlacementTree(api.Param(in_="d"), matrix=M)
    ot = oracle.place_all(D)
    compare_trees(kp, ot, n, 4 * n - 4)
    assert kp.printTree(synth.names(n)) == ot.newick(synth.names(n))


def test_placement_tie_heavy(ctx, oracle):
    n = 300
    codes, P, _ = make_msa(n, 1500, seed=6, regime="alisim", gap_cols=0.0, gap_runs=False)
    D = oracle.msa_dist_matrix(P, 1500, 1)
    M = api.Matrix.from_host(ctx, D)
    kp = api.KPlacementDeviceArrays(ctx)
    kp.allocateDeviceArrays(n)
    kp.findPlacementTree(api.Param(in_="d"), matrix=M)
    compare_trees(kp, oracle.place_all(D), n, 4 * n - 4)


def test_placement_large_additions_hit_the_0_0_2_tuple(ctx, oracle):
    # distances > 2 everywhere: every candidate has addLen >= 2 and the reference's default tuple wins
    n = 12
    rng = np.random.default_rng(1)
    D = rng.uniform(6, 9, (n, n)); D = np.tril(D, -1); D = D + D.T
    M = api.Matrix.from_host(ctx, D)
    kp = api.KPlacementDeviceArrays(ctx)
    kp.allocateDeviceArrays(n)
    kp.findPlacementTree(api.Param(in_="d"), matrix=M)
    compare_trees(kp, oracle.place_all(D), n, 4 * n - 4)


def test_placement_from_msa_batches(ctx, oracle):
    """-i m -m 1: rows come from the tiled distance kernel in batches (crosses the 512-row batch)."""
    n, L = 1100, 1000
    codes, P, _ = make_msa(n, L, seed=17)
    prm = api.Param(distanceType=2, in_="m")
    msa = api.MSADeviceArrays(ctx)
    msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
    kp = api.KPlacementDeviceArrays(ctx)
    kp.allocateDeviceArrays(n)
    kp.findPlacementTree(prm, msaDeviceArrays=msa)
    D = msa.distMatrix(prm).to_host()
    ot = oracle.place_all(D)
    compare_trees(kp, ot, n, 4 * n - 4)
    assert newick.rf_distance(kp.printTree(synth.names(n)), ot.newick(synth.names(n))) == 0


def test_placement_from_mash(ctx, oracle):
    n = 200
    codes, _ = synth.evolve(n, 3000, seed=23, gap_cols=0.01)
    seqs = synth.unaligned(codes)
    prm = api.Param(kmerSize=15, sketchSize=1000, in_="r")
    m = api.MashDeviceArrays(ctx)
    m.allocateDeviceArrays([synth.pack2_np(s) for s in seqs], np.array([len(s) for s in seqs], np.uint64), n, prm)
    m.sketchConstructionOnGpu()
    kp = api.KPlacementDeviceArrays(ctx)
    kp.allocateDeviceArrays(n)
    kp.findPlacementTree(prm, mashDeviceArrays=m)
    D = m.distMatrix().to_host()
    compare_trees(kp, oracle.place_all(D), n, 4 * n - 4)


def _backbone_struct(nwk, total):
    """Independent (Python) construction of the reference's node numbering from a Newick."""
    children, length, name = newick.parse(nwk)
    nn = len(children)
    internal_order = [v for v in range(nn) if children[v]]        # creation order == order of '('
    leaf_order = [v for v in range(nn) if not children[v]]         # order of appearance
    idx = {}
    for k, v in enumerate(leaf_order):
        idx[v] = k
    for k, v in enumerate(internal_order):
        idx[v] = total + k
    size = 2 * total + 2
    parent = np.full(size, -1, np.int32)
    bl = np.zeros(size, np.float64)
    ch = [[] for _ in range(size)]
    for v in range(nn):
        for c in children[v]:
            parent[idx[c]] = idx[v]
            bl[idx[c]] = float(np.float32(length[c]))
            ch[idx[v]].append(idx[c])
    off = np.zeros(size + 1, np.int32)
    flat = []
    for v in range(size):
        off[v + 1] = off[v] + len(ch[v])
        flat += ch[v]
    return idx[0], off, np.array(flat if flat else [0], np.int32), parent, bl, [name[v] for v in leaf_order]


def test_add_tips_onto_backbone(ctx, oracle):
    """--add -t backbone.nwk (src/tree_generation.cu:252-332): backbone = placement tree of the first B tips."""
    n, B = 260, 120
    codes, P, _ = make_msa(n, 1200, seed=29)
    D = oracle.msa_dist_matrix(P, 1200, 2)
    names = synth.names(n)
    bb = oracle.place_all(np.ascontiguousarray(D[:B, :B])).newick(names[:B])
    # backbone leaves are renumbered by order of appearance; queries follow (idMap :271-282)
    root, off, flat, parent, bl, leaf_names = _backbone_struct(bb, n)
    order = [names.index(x) for x in leaf_names] + list(range(B, n))
    Dp = np.ascontiguousarray(D[np.ix_(order, order)])
    kp = api.KPlacementDeviceArrays(ctx)
    kp.allocateDeviceArrays(n)
    assert kp.initializeDeviceArrays(bb) == B and kp.backbone_names == leaf_names
    M = api.Matrix.from_host(ctx, Dp)
    kp.addQuery(api.Param(in_="d"), matrix=M)
    ot = oracle.place_add(Dp, B, root, off, flat, parent, bl)
    compare_trees(kp, ot, n, 4 * n - 4)
    new_names = [names[i] for i in order]
    assert kp.printTree(new_names) == ot.newick(new_names)


@pytest.mark.parametrize("case", ["tiefree", "alisim", "add"])
def test_speculative_batch_path_is_exact(ctx, oracle, monkeypatch, case):
    """DIPB_PLACE_SPEC=1 (batch scoring + one-CTA sequential phase, placement.cu): same arrays as the oracle, including
    tie-heavy data and add-tips onto a loaded backbone.  Off by default (not faster), kept exact."""
    monkeypatch.setenv("DIPB_PLACE_SPEC", "1")
    monkeypatch.setenv("DIPB_PLACE_BATCH", "16")
    if case == "add":
        n, B = 260, 120
        codes, P, _ = make_msa(n, 1200, seed=29)
        D = oracle.msa_dist_matrix(P, 1200, 2)
        names = synth.names(n)
        bb = oracle.place_all(np.ascontiguousarray(D[:B, :B])).newick(names[:B])
        root, off, flat, parent, bl, leaf_names = _backbone_struct(bb, n)
        order = [names.index(x) for x in leaf_names] + list(range(B, n))
        Dp = np.ascontiguousarray(D[np.ix_(order, order)])
        kp = api.KPlacementDeviceArrays(ctx)
        kp.allocateDeviceArrays(n)
        assert kp.initializeDeviceArrays(bb) == B
        kp.addQuery(api.Param(in_="d"), matrix=api.Matrix.from_host(ctx, Dp))
        compare_trees(kp, oracle.place_add(Dp, B, root, off, flat, parent, bl), n, 4 * n - 4)
        return
    n = 700 if case == "tiefree" else 300
    codes, P, _ = make_msa(n, 1000, seed=207, regime=case, gap_cols=0.0, gap_runs=False)
    D = oracle.msa_dist_matrix(P, 1000, 2 if case == "tiefree" else 1)
    kp = api.KPlacementDeviceArrays(ctx)
    kp.allocateDeviceArrays(n)
    kp.findPlacementTree(api.Param(in_="d"), matrix=api.Matrix.from_host(ctx, D))
    compare_trees(kp, oracle.place_all(D), n, 4 * n - 4)
