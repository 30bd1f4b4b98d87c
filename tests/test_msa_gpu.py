"""GPU parity: aligned-MSA counts / distances through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

from dipper_b200 import api, synth
from conftest import make_msa

pytestmark = pytest.mark.gpu


def upload(ctx, P, L):
    m = api.MSADeviceArrays(ctx)
    m.allocateDeviceArrays(P, np.full(P.shape[0], L, np.uint64), P.shape[0], api.Param(distanceType=2, in_="m"))
    return m


@pytest.mark.parametrize("n,L", [(2, 1), (3, 15), (17, 16), (17, 17), (5, 31), (130, 33), (257, 1000), (300, 515)])
def test_counts_bit_exact(ctx, oracle, n, L):
    codes, P, _ = make_msa(n, L, seed=n * 1000 + L, gap_cols=0.1)
    msa = upload(ctx, P, L)
    m, u = msa.counts(0, n, 0, n)
    om, ou = oracle.msa_counts(P, L, 0, n, 0, n)
    assert np.array_equal(m, om)
    assert np.array_equal(u, ou)


def test_counts_with_lowercase_and_N(ctx, oracle):
    rng = np.random.default_rng(2)
    n, L = 20, 777
    seqs = ["".join(rng.choice(list("ACGTUacgtN-RY"), L, p=[.2, .2, .2, .15, .05, .02, .02, .02, .02, .05, .05, .01, .01]))
            for _ in range(n)]
    P = np.stack([oracle.pack4(s) for s in seqs])
    msa = upload(ctx, P, L)
    m, u = msa.counts(0, n, 0, n)
    om, ou = oracle.msa_counts(P, L, 0, n, 0, n)
    assert np.array_equal(m, om) and np.array_equal(u, ou)


def test_long_alignment_segments(ctx, oracle):
    # > 65 024 sites forces the multi-segment (int32-accumulating) path
    n, L = 6, 70001
    codes, P, _ = make_msa(n, L, seed=9, gap_cols=0.01, gap_runs=False)
    msa = upload(ctx, P, L)
    m, u = msa.counts(0, n, 0, n)
    om, ou = oracle.msa_counts(P, L, 0, n, 0, n)
    assert np.array_equal(m, om) and np.array_equal(u, ou)
    D = msa.distMatrix(api.Param(distanceType=2, in_="m")).to_host()
    O = oracle.msa_dist_matrix(P, L, 2)
    assert np.allclose(D, O, rtol=1e-6, atol=0)


@pytest.mark.parametrize("dist_type", [1, 2, 3, 4, 5, 6])
def test_distance_rows_all_models(ctx, oracle, dist_type):
    n, L = 150, 2000
    codes, P, _ = make_msa(n, L, seed=21, gap_cols=0.05)
    msa = upload(ctx, P, L)
    prm = api.Param(distanceType=dist_type, in_="m")
    for row in (1, 2, 77, 128, 149):
        got = msa.distConstructionOnGpu(prm, row)
        exp = oracle.msa_dist_row(P, L, row, dist_type)
        ok = np.isfinite(exp)
        assert np.array_equal(np.isfinite(got), ok)
        assert np.allclose(got[ok], exp[ok], rtol=1e-6, atol=0), (row, dist_type)   # north_star: 1e-6 relative


@pytest.mark.parametrize("dist_type", [1, 2, 4])
@pytest.mark.parametrize("n", [2, 3, 129, 400])
def test_full_matrix(ctx, oracle, n, dist_type):
    L = 1200
    codes, P, _ = make_msa(n, L, seed=n, gap_cols=0.03)
    msa = upload(ctx, P, L)
    D = msa.distMatrix(api.Param(distanceType=dist_type, in_="m")).to_host()
    O = oracle.msa_dist_matrix(P, L, dist_type)
    assert np.array_equal(D, D.T) and np.all(np.diag(D) == 0)
    ok = np.isfinite(O)
    assert np.allclose(D[ok], O[ok], rtol=1e-6, atol=0)
    if dist_type == 1:
        assert np.array_equal(D, O)     # p-distance is one fp64 division: bit-exact


def test_all_gap_pair_is_nan_like_reference(ctx, oracle):
    # useful == 0 -> 0/0 (src/MSA.cu:233 has no guard)
    P = np.stack([oracle.pack4("----"), oracle.pack4("NNNN"), oracle.pack4("ACGT")])
    msa = upload(ctx, P, 4)
    row = msa.distConstructionOnGpu(api.Param(distanceType=1, in_="m"), 1)
    assert np.isnan(row[0])
    row2 = msa.distConstructionOnGpu(api.Param(distanceType=1, in_="m"), 2)
    assert row2[0] == 1.0 and row2[1] == 1.0   # base opposite a gap counts as a mismatch


def test_golden_fixtures_through_cuda(ctx):
    import glob, os
    for fn in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*msa*.npz"))):
        z = np.load(fn)
        if str(z["kind"]) not in ("msa", "ref_msa"):
            continue                       # D&C / add-tips fixtures: tests/test_ref_golden_gpu.py
        P, L = z["packed"], int(z["seq_len"])
        msa = upload(ctx, P, L)
        if str(z["kind"]) == "ref_msa":
            # produced by the reference's own CUDA objects (tools/make_ref_golden.py): rows, NJ, k-closest and exact-mode trees
            from dipper_b200 import newick
            n = P.shape[0]
            low = np.tril_indices(n, -1)
            for t in (1, 2):
                D = msa.distMatrix(api.Param(distanceType=t, in_="m")).to_host()
                assert np.allclose(D[low], z["rows_%d" % t][low], rtol=1e-6, atol=0), (fn, t)
            prm = api.Param(distanceType=2, in_="m")
            nj = api.NJDeviceArrays(ctx)
            nj.getDismatrix(n, prm, msaDeviceArrays=msa)
            nwk = nj.findNeighbourJoiningTree(synth.names(n))
            assert newick.rf_distance(nwk, str(z["nj_newick"])) == 0 and newick.max_branch_diff(nwk, str(z["nj_newick"])) < 1e-5, fn
            kp = api.KPlacementDeviceArrays(ctx); kp.allocateDeviceArrays(n)
            kp.findPlacementTree(prm, msaDeviceArrays=msa)
            assert kp.printTree(synth.names(n)) == str(z["place_newick"]), fn
            pl = api.PlacementDeviceArrays(ctx); pl.allocateDeviceArrays(n)
            pl.findPlacementTree(prm, msaDeviceArrays=msa)
            assert pl.printTree(synth.names(n)) == str(z["place_exact_newick"]), fn
            continue
        m, u = msa.counts(0, P.shape[0], 0, P.shape[0])
        assert np.array_equal(m, z["match"]) and np.array_equal(u, z["useful"]), fn
        for t in z["dist_types"]:
            D = msa.distMatrix(api.Param(distanceType=int(t), in_="m")).to_host()
            G = z["dist_%d" % t]
            ok = np.isfinite(G)
            assert np.allclose(D[ok], G[ok], rtol=1e-6, atol=0), (fn, t)


def test_upload_rejects_ragged(ctx, oracle):
    rows = [oracle.pack4("ACGTACGT"), oracle.pack4("ACGT")]
    m = api.MSADeviceArrays(ctx)
    with pytest.raises(api.DipperError):
        m.allocateDeviceArrays(rows, np.array([8, 4], np.uint64), 2, api.Param(in_="m"))


def test_row_sharded_matrix_sums_to_full(ctx, oracle):
    n, L = 515, 640
    codes, P, _ = make_msa(n, L, seed=31)
    msa = upload(ctx, P, L)
    prm = api.Param(distanceType=2, in_="m")
    full = msa.distMatrix(prm).to_host()
    a = msa.distMatrix(prm, 0, 256).to_host()
    b = msa.distMatrix(prm, 256, n).to_host()
    assert np.array_equal(a + b, full)


@pytest.mark.parametrize("n,L", [(2, 40), (129, 1000), (300, 515), (700, 4000)])
@pytest.mark.parametrize("dist_type", [1, 2])
@pytest.mark.parametrize("pair", ["0", "1"])
def test_tensor_core_path_is_bit_identical(ctx, oracle, n, L, dist_type, pair, monkeypatch):
    """msa_tc.cu (tcgen05 int8 GEMMs; pair = 2-CTA cta_group::2 kernel) must give exactly the matrix of the popcount kernel."""
    monkeypatch.setenv("DIPB_MSA_TC2", pair)
    codes, P, _ = make_msa(n, L, seed=900 + n, gap_cols=0.05)
    msa = upload(ctx, P, L)
    prm = api.Param(distanceType=dist_type, in_="m")
    monkeypatch.setenv("DIPB_MSA_TC", "0")
    ref = msa.distMatrix(prm).to_host()          # popcount kernel
    monkeypatch.setenv("DIPB_MSA_TC", "2")
    got = msa.distMatrix(prm).to_host()          # tcgen05 kernel (the default)
    monkeypatch.delenv("DIPB_MSA_TC")
    assert np.array_equal(got, ref, equal_nan=True)
    assert np.allclose(got, oracle.msa_dist_matrix(P, L, dist_type), rtol=1e-6, atol=0, equal_nan=True)


@pytest.mark.parametrize("fmt", ["0", "2"])
@pytest.mark.parametrize("pair", ["0", "1"])
def test_tensor_core_operand_formats_at_full_length(ctx, fmt, pair, monkeypatch):
    """Both operand encodings of msa_tc.cu (int8 + kind::i8 with s32 accumulators; e2m1, 4 bits per element in HBM, unpacked by
    the TMA, kind::f8f6f4 with f32 accumulators -- the default) against the popcount kernel at the headline sequence length:
    near-identical sequences drive every accumulator to ~ 3 L = 90 000, past 2^16, where an inexact f32 accumulation would show."""
    from dipper_b200 import synth
    n, L = 600, 30000
    rng = np.random.default_rng(5)
    codes = np.tile(rng.integers(0, 4, L).astype(np.uint8), (n, 1))
    mut = rng.random((n, L)) < 0.01
    codes[mut] = rng.integers(0, 4, int(mut.sum())).astype(np.uint8)
    codes[rng.random((n, L)) < 0.002] = 4
    P = synth.pack4_np(codes)
    prm = api.Param(distanceType=2, in_="m")
    monkeypatch.setenv("DIPB_MSA_TC2", pair)
    monkeypatch.setenv("DIPB_MSA_TC", "0")
    ref = upload(ctx, P, L).distMatrix(prm).to_host()
    monkeypatch.setenv("DIPB_MSA_TC", "2")
    monkeypatch.setenv("DIPB_TC_FMT", fmt)
    got = upload(ctx, P, L).distMatrix(prm).to_host()      # (the format is fixed at an alignment's first expansion)
    assert np.array_equal(got, ref, equal_nan=True)
    assert ref[1, 0] > 0 and np.isfinite(ref).all()


@pytest.mark.parametrize("dist_type", [1, 2])
@pytest.mark.parametrize("r0,r1,ncols", [(0, 64, 700), (100, 700, 100), (130, 515, 515), (511, 700, 257)])
@pytest.mark.parametrize("pair", ["0", "1"])
def test_tensor_core_block_rows(ctx, oracle, dist_type, r0, r1, ncols, pair, monkeypatch):
    """Row blocks of >= 64 rows (placement batches, D&C stage 2) go through the tcgen05 kernels too."""
    monkeypatch.setenv("DIPB_MSA_TC2", pair)
    n, L = 700, 2100
    codes, P, _ = make_msa(n, L, seed=77, gap_cols=0.05)
    msa = upload(ctx, P, L)
    prm = api.Param(distanceType=dist_type, in_="m")
    monkeypatch.setenv("DIPB_MSA_TC", "0")
    ref = msa.distBlock(prm, r0, r1, ncols)
    monkeypatch.setenv("DIPB_MSA_TC", "2")       # force: blocks this small normally stay on the popcount kernel
    got = msa.distBlock(prm, r0, r1, ncols)
    monkeypatch.delenv("DIPB_MSA_TC")
    full = oracle.msa_dist_matrix(P, L, dist_type)
    off = np.arange(r0, r1)[:, None] != np.arange(ncols)[None, :]      # the diagonal is unspecified in block mode
    assert np.array_equal(got[off], ref[off], equal_nan=True)
    assert np.allclose(got[off], full[r0:r1, :ncols][off], rtol=1e-6, atol=0, equal_nan=True)


@pytest.mark.parametrize("cut", [1, 128, 300, 699])
@pytest.mark.parametrize("pair", ["0", "1"])
def test_tensor_core_row_sharded_matrix(ctx, oracle, cut, pair, monkeypatch):
    monkeypatch.setenv("DIPB_MSA_TC2", pair)
    n, L = 700, 1300
    codes, P, _ = make_msa(n, L, seed=78)
    msa = upload(ctx, P, L)
    prm = api.Param(distanceType=2, in_="m")
    monkeypatch.setenv("DIPB_MSA_TC", "0")
    full = msa.distMatrix(prm).to_host()
    monkeypatch.setenv("DIPB_MSA_TC", "2")
    a = msa.distMatrix(prm, 0, cut).to_host()
    b = msa.distMatrix(prm, cut, n).to_host()
    monkeypatch.delenv("DIPB_MSA_TC")
    assert np.array_equal(a + b, full)


def test_row_block_gather_and_mirror_equals_full_matrix(ctx, oracle):
    """The multi-GPU gather of bench.py on one device: copy the other shard's (contiguous) rows into rank 0's
    matrix, mirror them (dipb_matrix_mirror_rows) and compare with the unsharded matrix."""
    import torch
    from dipper_b200._lib import check, lib

    class DevView:
        def __init__(self, ptr, count):
            self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}

    n, L = 700, 900
    codes, P, _ = make_msa(n, L, seed=91)
    msa = upload(ctx, P, L)
    prm = api.Param(distanceType=2, in_="m")
    full = msa.distMatrix(prm).to_host()
    cut = 384
    A = msa.distMatrix(prm, 0, cut)
    B = msa.distMatrix(prm, cut, n)
    tA = torch.as_tensor(DevView(lib().dipb_matrix_device_ptr(A.h), n * n), device="cuda")
    tB = torch.as_tensor(DevView(lib().dipb_matrix_device_ptr(B.h), n * n), device="cuda")
    tA[cut * n:].copy_(tB[cut * n:])
    torch.cuda.synchronize()
    check(lib().dipb_matrix_mirror_rows(A.h, cut, n))
    ctx.sync()
    assert np.array_equal(A.to_host(), full)
