"""GPU parity: neighbor joining through the C ABI vs the CPU oracle (bit-identical trees)."""
import numpy as np
import pytest

from dipper_b200 import api, newick, synth
from conftest import make_msa

pytestmark = pytest.mark.gpu

ALGOS = [api.NJ_FULLSCAN, api.NJ_PRUNED, api.NJ_CLUSTER]


def run_nj(ctx, D, algo):
    nj = api.NJDeviceArrays(ctx)
    nj.setMatrix(D)
    nwk = nj.findNeighbourJoiningTree(synth.names(D.shape[0]), algo)
    res = nj.result
    nj.deallocateDeviceArrays()
    return nwk, res


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("n", [2, 3, 4, 5, 33, 255, 256, 257, 700])
def test_nj_matches_oracle_tiefree(ctx, oracle, n, algo):
    codes, P, _ = make_msa(n, 900, seed=100 + n)
    D = oracle.msa_dist_matrix(P, 900, 2)
    nwk, (c0, c1, l0, l1) = run_nj(ctx, D, algo)
    o0, o1, ol0, ol1 = oracle.nj(D)
    assert np.array_equal(c0, o0) and np.array_equal(c1, o1)
    assert np.array_equal(l0, ol0) and np.array_equal(l1, ol1)      # same fp64 order -> bit-identical
    assert nwk == oracle.nj_newick(o0, o1, ol0, ol1, synth.names(n))


@pytest.mark.parametrize("algo", ALGOS)
def test_nj_tie_heavy_alisim_regime(ctx, oracle, algo):
    # many identical sequences -> exact ties; the reference's scan-order tie-break decides
    n = 300
    codes, P, _ = make_msa(n, 2000, seed=5, regime="alisim", gap_cols=0.0, gap_runs=False)
    D = oracle.msa_dist_matrix(P, 2000, 1)
    assert (D[np.triu_indices(n, 1)] == 0).sum() > 10
    nwk, (c0, c1, l0, l1) = run_nj(ctx, D, algo)
    o0, o1, ol0, ol1 = oracle.nj(D)
    assert np.array_equal(c0, o0) and np.array_equal(c1, o1)
    assert np.array_equal(l0, ol0) and np.array_equal(l1, ol1)


@pytest.mark.parametrize("algo", ALGOS)
def test_nj_quantised_mash_like_distances(ctx, oracle, algo):
    # Mash distances take <= 1001 values: heavy exact ties of another kind
    rng = np.random.default_rng(3)
    n = 200
    j = rng.integers(1, 1001, (n, n)) / 1000.0
    D = np.minimum(1.0, np.abs(np.log(2 * j / (1 + j)) / 15))
    D = np.tril(D, -1)
    D = D + D.T
    nwk, (c0, c1, l0, l1) = run_nj(ctx, D, algo)
    o0, o1, ol0, ol1 = oracle.nj(D)
    assert np.array_equal(c0, o0) and np.array_equal(c1, o1)
    assert np.array_equal(l0, ol0) and np.array_equal(l1, ol1)


def test_pruned_equals_fullscan_at_scale(ctx):
    # beyond what the CPU oracle finishes in seconds: the exhaustive GPU search is the yard-stick
    n = 3000
    codes, P, _ = make_msa(n, 1500, seed=77, gap_runs=False)
    msa = api.MSADeviceArrays(ctx)
    msa.allocateDeviceArrays(P, np.full(n, 1500, np.uint64), n, api.Param(in_="m"))
    prm = api.Param(distanceType=2, in_="m")
    res = []
    for algo in ALGOS:
        nj = api.NJDeviceArrays(ctx)
        nj.getDismatrix(n, prm, msaDeviceArrays=msa)
        nj.findNeighbourJoiningTree(synth.names(n), algo)
        res.append(nj.result)
        nj.deallocateDeviceArrays()
    for other in res[1:]:
        for a, b in zip(res[0], other):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("env", [{"DIPB_NJ_CLUSTER": "8"}, {"DIPB_NJ_THREADS": "256"}, {"DIPB_NJ_THREADS": "1024"}, {"DIPB_NJ_HELPERS": "0"},
                                 {"DIPB_NJ_HELPERS": "1"}, {"DIPB_NJ_SLACK": "0"}, {"DIPB_NJ_DBG": "1"}])
@pytest.mark.parametrize("n", [130, 1500])
def test_cluster_kernel_variants_equal_fullscan(ctx, env, n, monkeypatch):
    """The launch shapes and modes of nj_cluster_kernel that the default run does not take (8-CTA cluster, 256 / 1024 threads,
    no / one helper cluster, no early unit refresh, whole-row rescans): all must give the exhaustive search's arrays.  n = 130
    also crosses the 128-row staging tile on the first search."""
    codes, P, _ = make_msa(n, 1200, seed=300 + n, gap_runs=False)
    msa = api.MSADeviceArrays(ctx)
    msa.allocateDeviceArrays(P, np.full(n, 1200, np.uint64), n, api.Param(in_="m"))
    prm = api.Param(distanceType=2, in_="m")

    def run(algo):
        nj = api.NJDeviceArrays(ctx)
        nj.getDismatrix(n, prm, msaDeviceArrays=msa)
        nj.findNeighbourJoiningTree(synth.names(n), algo)
        r = nj.result
        nj.deallocateDeviceArrays()
        return r

    ref = run(api.NJ_FULLSCAN)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    got = run(api.NJ_CLUSTER)
    for a, b in zip(ref, got):
        assert np.array_equal(a, b)


def test_msa_to_tree_end_to_end_rf_zero(ctx, oracle):
    """-i m -d 2 -m 2: CUDA distances + CUDA NJ vs oracle distances + oracle NJ."""
    n, L = 500, 3000
    codes, P, _ = make_msa(n, L, seed=8)
    msa = api.MSADeviceArrays(ctx)
    prm = api.Param(distanceType=2, in_="m")
    msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
    nj = api.NJDeviceArrays(ctx)
    nj.getDismatrix(n, prm, msaDeviceArrays=msa)
    nwk = nj.findNeighbourJoiningTree(synth.names(n))
    o = oracle.nj(oracle.msa_dist_matrix(P, L, 2))
    onwk = oracle.nj_newick(*o, synth.names(n))
    assert newick.rf_distance(nwk, onwk) == 0                      # north_star: RF = 0
    assert newick.max_branch_diff(nwk, onwk) < 1e-5                # north_star: 1e-5


def test_phylip_input_path(ctx, oracle, tmp_path):
    """-i d: rows parsed through float32 like the reference's stof (src/matrix_reader.cu:42)."""
    n = 60
    codes, P, _ = make_msa(n, 800, seed=14)
    D = oracle.msa_dist_matrix(P, 800, 2)
    names = synth.names(n)
    path = str(tmp_path / "m.phy")
    synth.write_phylip(path, names, D, lower=True)
    rd = api.MatrixReader()
    f = open(path)
    nseq = int(f.readline())
    rd.allocateDeviceArrays(nseq, f)
    nj = api.NJDeviceArrays(ctx)
    nj.getDismatrix(nseq, api.Param(in_="d"), matrixReader=rd)
    got = nj.matrix.to_host()
    D32 = np.array([[np.float32("%.6f" % v) for v in row] for row in D], np.float64)
    exp = np.tril(D32, -1) + np.tril(D32, -1).T
    assert np.array_equal(got, exp) and rd.name == names
    nwk = nj.findNeighbourJoiningTree(rd.name)
    o = oracle.nj(exp)
    assert nwk == oracle.nj_newick(*o, names)
