"""GPU parity: exact placement mode (-p 0, src/placement.cu) through the C ABI vs the CPU oracle.
The slot arrays must be identical (bit-equal lengths), hence the Newick text too."""
import numpy as np
import pytest

from dipper_b200 import api, newick, synth
from conftest import make_msa

pytestmark = pytest.mark.gpu


def compare_trees(pl, otree, n):
    mine = pl.export()
    ref = otree.arrays()
    live = 4 * n - 4
    for k in ("head",):
        assert np.array_equal(mine[k][: 2 * n - 1], ref[k][: 2 * n - 1]), k
    for k in ("e", "nxt", "belong"):
        assert np.array_equal(mine[k][:live], ref[k][:live]), k
    assert np.array_equal(mine["len"][:live], ref["len"][:live])


@pytest.mark.parametrize("n", [2, 3, 4, 10, 33, 129, 700])
def test_exact_placement_from_matrix_matches_oracle(ctx, oracle, n):
    codes, P, _ = make_msa(n, 800, seed=300 + n)
    D = oracle.msa_dist_matrix(P, 800, 2)
    M = api.Matrix.from_host(ctx, D)
    pl = api.PlacementDeviceArrays(ctx)
    pl.allocateDeviceArrays(n)
    pl.findPlacementTree(api.Param(in_="d"), matrix=M)
    ot = oracle.place_exact(D)
    compare_trees(pl, ot, n)
    assert pl.printTree(synth.names(n)) == ot.newick(synth.names(n))


def test_exact_placement_recovers_an_additive_tree(ctx, oracle):
    """Path metric of a random tree: every limit is attained exactly, the true tree comes back (RF = 0)."""
    n, rng = 400, np.random.default_rng(11)
    adj = {0: {1: 0.05}, 1: {0: 0.05}}
    nxt = n
    for leaf in range(2, n):
        a = int(rng.choice(list(adj.keys())))
        b = int(rng.choice(list(adj[a].keys())))
        L = adj[a][b]
        f = rng.uniform(0.2, 0.8) * L
        m, nxt = nxt, nxt + 1
        del adj[a][b]; del adj[b][a]
        pend = rng.uniform(0.01, 0.1)
        adj[m] = {a: f, b: L - f, leaf: pend}
        adj[a][m] = f; adj[b][m] = L - f; adj[leaf] = {m: pend}
    D = np.zeros((n, n))
    for s in range(n):
        st = [(s, -1, 0.0)]
        while st:
            v, p, d = st.pop()
            if v < n:
                D[s, v] = d
            st.extend((w, v, d + l) for w, l in adj[v].items() if w != p)
    M = api.Matrix.from_host(ctx, D)
    pl = api.PlacementDeviceArrays(ctx)
    pl.allocateDeviceArrays(n)
    pl.findPlacementTree(api.Param(in_="d"), matrix=M)
    names = synth.names(n)

    def nwk(v, p):
        ch = [w for w in adj[v] if w != p]
        return names[v] if not ch else "(" + ",".join(nwk(w, v) + ":%g" % adj[v][w] for w in ch) + ")"
    import sys
    sys.setrecursionlimit(10000)
    truth = nwk(n, -1) + ";"
    mine = pl.printTree(names)
    assert newick.rf_distance(mine, truth) == 0
    assert newick.max_branch_diff(mine, truth) < 1e-6   # %g text: 6 significant digits
    compare_trees(pl, oracle.place_exact(D), n)


def test_exact_placement_tie_heavy(ctx, oracle):
    n = 300
    codes, P, _ = make_msa(n, 1500, seed=6, regime="alisim", gap_cols=0.0, gap_runs=False)
    D = oracle.msa_dist_matrix(P, 1500, 1)
    M = api.Matrix.from_host(ctx, D)
    pl = api.PlacementDeviceArrays(ctx)
    pl.allocateDeviceArrays(n)
    pl.findPlacementTree(api.Param(in_="d"), matrix=M)
    compare_trees(pl, oracle.place_exact(D), n)


def test_exact_placement_from_msa_batches(ctx, oracle):
    """-i m -p 0: rows arrive in 512-tip batches, the tree state is saved / restored between launches."""
    n, L = 1300, 1000
    codes, P, _ = make_msa(n, L, seed=19)
    prm = api.Param(distanceType=2, in_="m")
    msa = api.MSADeviceArrays(ctx)
    msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
    pl = api.PlacementDeviceArrays(ctx)
    pl.allocateDeviceArrays(n)
    pl.findPlacementTree(prm, msaDeviceArrays=msa)
    D = msa.distMatrix(prm).to_host()
    ot = oracle.place_exact(D)
    compare_trees(pl, ot, n)
    assert pl.printTree(synth.names(n)) == ot.newick(synth.names(n))


def test_exact_placement_from_mash(ctx, oracle):
    n = 200
    codes, _ = synth.evolve(n, 3000, seed=29, gap_cols=0.01)
    seqs = synth.unaligned(codes)
    prm = api.Param(kmerSize=15, sketchSize=1000, in_="r")
    m = api.MashDeviceArrays(ctx)
    m.allocateDeviceArrays([synth.pack2_np(s) for s in seqs], np.array([len(s) for s in seqs], np.uint64), n, prm)
    m.sketchConstructionOnGpu()
    pl = api.PlacementDeviceArrays(ctx)
    pl.allocateDeviceArrays(n)
    pl.findPlacementTree(prm, mashDeviceArrays=m)
    D = m.distMatrix().to_host()
    # placement consumes row i, column j < i (A = column, B = row in the asymmetric merge)
    Dl = np.tril(D, -1)
    compare_trees(pl, oracle.place_exact(Dl + Dl.T), n)


@pytest.mark.parametrize("mode", ["0", "1"])
def test_exact_placement_earlier_kernel_variants(ctx, oracle, monkeypatch, mode):
    """DIPB_EXACT_FLOW=0: level steps with cluster barriers; 1: data flow over all nodes (the default flows over
    internal nodes only).  Same arrays."""
    monkeypatch.setenv("DIPB_EXACT_FLOW", mode)
    n, L = 900, 900
    codes, P, _ = make_msa(n, L, seed=78)
    prm = api.Param(distanceType=2, in_="m")
    msa = api.MSADeviceArrays(ctx)
    msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
    pl = api.PlacementDeviceArrays(ctx)
    pl.allocateDeviceArrays(n)
    pl.findPlacementTree(prm, msaDeviceArrays=msa)
    compare_trees(pl, oracle.place_exact(msa.distMatrix(prm).to_host()), n)


@pytest.mark.parametrize("n", [2, 3, 50, 1300])
def test_exact_placement_global_memory_kernel(ctx, oracle, monkeypatch, n):
    """The kernel used beyond one cluster's shared memory (> 49 152 tips), forced here on small inputs: whole grid,
    state in global memory, values travel through L2.  Same arrays; 1300 tips cross the 512-tip launch boundary."""
    monkeypatch.setenv("DIPB_EXACT_GLOBAL", "1")
    L = 900
    codes, P, _ = make_msa(n, L, seed=400 + n)
    prm = api.Param(distanceType=2, in_="m")
    msa = api.MSADeviceArrays(ctx)
    msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
    pl = api.PlacementDeviceArrays(ctx)
    pl.allocateDeviceArrays(n)
    pl.findPlacementTree(prm, msaDeviceArrays=msa)
    ot = oracle.place_exact(msa.distMatrix(prm).to_host())
    compare_trees(pl, ot, n)
    assert pl.printTree(synth.names(n)) == ot.newick(synth.names(n))


def test_exact_placement_8_cta_cluster(ctx, oracle, monkeypatch):
    monkeypatch.setenv("DIPB_EXACT_CLUSTER", "8")
    n = 500
    codes, P, _ = make_msa(n, 900, seed=77)
    D = oracle.msa_dist_matrix(P, 900, 2)
    M = api.Matrix.from_host(ctx, D)
    pl = api.PlacementDeviceArrays(ctx)
    pl.allocateDeviceArrays(n)
    pl.findPlacementTree(api.Param(in_="d"), matrix=M)
    compare_trees(pl, oracle.place_exact(D), n)


def test_exact_placement_reports_the_default_tuple_case(ctx):
    # distances > 4 everywhere: every pendant length is >= 2, the reference's (0,0,2) default tuple would win
    n = 12
    rng = np.random.default_rng(1)
    D = rng.uniform(6, 9, (n, n)); D = np.tril(D, -1); D = D + D.T
    M = api.Matrix.from_host(ctx, D)
    pl = api.PlacementDeviceArrays(ctx)
    pl.allocateDeviceArrays(n)
    with pytest.raises(Exception, match="pendant length"):
        pl.findPlacementTree(api.Param(in_="d"), matrix=M)


def test_exact_placement_size_limit_of_the_cluster_kernel(ctx, monkeypatch):
    monkeypatch.setenv("DIPB_EXACT_GLOBAL", "0")     # without the global-memory kernel the cluster's capacity is reported
    n, L = api.PlacementDeviceArrays.maxTips() + 64, 64
    P = np.zeros((n, L // 16), np.uint64)
    prm = api.Param(distanceType=1, in_="m")
    msa = api.MSADeviceArrays(ctx)
    msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
    pl = api.PlacementDeviceArrays(ctx)
    pl.allocateDeviceArrays(n)
    with pytest.raises(Exception, match="exceed"):
        pl.findPlacementTree(prm, msaDeviceArrays=msa)
