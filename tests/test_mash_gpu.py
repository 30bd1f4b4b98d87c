"""GPU parity: MinHash sketches (bit-exact) and Mash distances through the C ABI vs the oracle."""
import glob
import os

import numpy as np
import pytest

from dipper_b200 import api, synth
from conftest import make_msa

pytestmark = pytest.mark.gpu


def unaligned_case(n, L, seed, ragged=True):
    codes, _ = synth.evolve(n, L, seed=seed, regime="tiefree", gap_cols=0.02 if ragged else 0.0, gap_runs=ragged)
    seqs = synth.unaligned(codes)
    return seqs, [synth.pack2_np(s) for s in seqs], np.array([len(s) for s in seqs], np.uint64)


def upload(ctx, packed, lens, k=15, s=1000):
    m = api.MashDeviceArrays(ctx)
    m.allocateDeviceArrays(packed, lens, len(lens), api.Param(kmerSize=k, sketchSize=s, in_="r"))
    m.sketchConstructionOnGpu()
    return m


@pytest.mark.parametrize("k", [15, 8, 16, 21, 32])
def test_sketches_bit_exact(ctx, oracle, k):
    seqs, packed, lens = unaligned_case(12, 3000, seed=k)
    m = upload(ctx, packed, lens, k=k)
    got = m.sketches()
    flat, offs, ln = synth.flatten2(seqs)
    exp = oracle.sketch_all(flat, offs, ln, k, 1000)
    assert np.array_equal(got, exp)


def test_sketch_short_and_long_sequences(ctx, oracle):
    rng = np.random.default_rng(4)
    # fewer k-mers than the sketch size (padding), exactly k bases, and a long one (several buffer sorts)
    seqs = [rng.integers(0, 4, L).astype(np.uint8) for L in (15, 16, 40, 999, 1014, 1015, 5000, 60000)]
    packed = [synth.pack2_np(s) for s in seqs]
    lens = np.array([len(s) for s in seqs], np.uint64)
    m = upload(ctx, packed, lens)
    got = m.sketches()
    flat, offs, ln = synth.flatten2(seqs)
    exp = oracle.sketch_all(flat, offs, ln, 15, 1000)
    assert np.array_equal(got, exp)
    assert got[0, 0] != np.uint64(0xFFFFFFFFFFFFFFFF) and got[0, 1] == np.uint64(0xFFFFFFFFFFFFFFFF)


def test_sketch_keeps_duplicate_hashes(ctx, oracle):
    # a repeat-rich sequence: the reference never de-duplicates (src/mash.cu:326-345)
    unit = np.array([0, 1, 2, 3, 1, 1, 2, 0, 3, 3, 2, 1, 0, 0, 2, 3, 1], np.uint8)
    seq = np.tile(unit, 200)
    m = upload(ctx, [synth.pack2_np(seq)], np.array([len(seq)], np.uint64))
    got = m.sketches()[0]
    exp = oracle.sketch(synth.pack2_np(seq), len(seq), 15, 1000)
    assert np.array_equal(got, exp)
    assert len(np.unique(got)) < 100


def test_mash_distance_matrix_and_rows(ctx, oracle):
    seqs, packed, lens = unaligned_case(150, 4000, seed=7)
    m = upload(ctx, packed, lens)
    sk = m.sketches()
    D = m.distMatrix().to_host()
    O = oracle.mash_dist_matrix(sk, 15)
    assert np.array_equal(D, D.T) and np.all(np.diag(D) == 0)
    assert np.allclose(D, O, rtol=1e-6, atol=0)            # north_star tolerance
    assert np.abs(D - O).max() < 1e-15                      # in practice only libm vs libdevice log
    prm = api.Param(in_="r")
    for row in (1, 7, 16, 17, 149):
        assert np.array_equal(m.distConstructionOnGpu(prm, row), D[row, :row])


def test_mash_merge_rule_on_crafted_sketches(ctx, oracle):
    """inter / uni corner cases: identical, disjoint, duplicates, padding, asymmetric roles."""
    s = 1000
    rng = np.random.default_rng(9)
    base = np.sort(rng.integers(0, 2**62, s, dtype=np.uint64))
    other = np.sort(rng.integers(0, 2**62, s, dtype=np.uint64))
    dup = np.sort(np.concatenate([base[:300], base[:300], base[300:700]]))
    pad = base.copy(); pad[400:] = np.uint64(0xFFFFFFFFFFFFFFFF)
    pad2 = other.copy(); pad2[10:] = np.uint64(0xFFFFFFFFFFFFFFFF)
    half = np.sort(np.concatenate([base[::2], other[::2]]))
    sk = np.stack([base, other, dup, pad, pad2, half, base])
    m = api.MashDeviceArrays(ctx)
    m.setSketches(sk, api.Param(kmerSize=15, sketchSize=s, in_="r"))
    D = m.distMatrix().to_host()
    O = oracle.mash_dist_matrix(sk, 15)
    assert np.allclose(D, O, rtol=1e-12, atol=0)
    assert D[6, 0] == 0.0
    assert D[1, 0] == min(1.0, abs(np.log(2 * (1 / 1000) / (1 + 1 / 1000)) / 15))


@pytest.mark.parametrize("s", [2, 10, 64, 500, 1000, 1024, 1100])
def test_warp_merge_equals_thread_merge_and_oracle(ctx, oracle, monkeypatch, s):
    """mash_rank_kernel (default: rank-compressed keys, interleaved conflict-free tiles) vs mash_tile_kernel (the reference's
    sequential loop per thread on the 64-bit hashes) vs mash_warp_kernel (merge path, 32 lanes per pair) and the oracle:
    overlapping, duplicated and padded sketches of every density."""
    rng = np.random.default_rng(100 + s)
    pool = np.sort(rng.integers(0, 2**63, 3 * s, dtype=np.uint64))
    rows = []
    for r in range(37):
        kind = r % 6
        if kind == 0:
            v = rng.choice(pool, s, replace=False)
        elif kind == 1:
            v = rng.choice(pool[: s + s // 2], s, replace=False)           # heavy overlap
        elif kind == 2:
            v = rng.choice(pool[: max(2, s // 3)], s, replace=True)        # many duplicates
        elif kind == 3:
            v = rng.choice(pool, s, replace=False); v[rng.integers(1, s):] = np.uint64(0xFFFFFFFFFFFFFFFF)   # padded
        elif kind == 4:
            v = pool[:s].copy()                                            # identical lists
        else:
            v = rng.integers(0, 2**63, s, dtype=np.uint64)                 # disjoint
        rows.append(np.sort(v.astype(np.uint64)))
    sk = np.stack(rows)
    prm = api.Param(kmerSize=15, sketchSize=s, in_="r")
    m = api.MashDeviceArrays(ctx)
    m.setSketches(sk, prm)
    D = m.distMatrix().to_host()
    R = m.distConstructionOnGpu(prm, 29)
    monkeypatch.setenv("DIPB_MASH_RANKS", "0")     # 64-bit hashes, thread-per-pair (the reference's loop verbatim)
    D0 = m.distMatrix().to_host()
    monkeypatch.setenv("DIPB_MASH_WARP", "1")      # 64-bit hashes, merge-path warp-per-pair
    D1 = m.distMatrix().to_host()
    assert np.array_equal(D, D0) and np.array_equal(D1, D0)
    assert np.array_equal(R, D0[29, :29])
    O = oracle.mash_dist_matrix(sk, 15)
    assert np.allclose(D, O, rtol=1e-12, atol=0)


def test_mash_golden_fixture(ctx):
    for fn in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*mash*.npz"))):
        z = np.load(fn)
        if str(z["kind"]) not in ("mash", "ref_mash"):
            continue                       # D&C fixtures: tests/test_ref_golden_gpu.py
        m = api.MashDeviceArrays(ctx)
        lens, offs, flat = z["lens"], z["offsets"], z["flat"]
        rows = [flat[int(offs[i]): int(offs[i]) + (int(lens[i]) + 31) // 32] for i in range(len(lens))]
        m.allocateDeviceArrays(rows, lens, len(lens), api.Param(kmerSize=int(z["k"]), sketchSize=int(z["s"]), in_="r"))
        m.sketchConstructionOnGpu()
        assert np.array_equal(m.sketches(), z["sketches"]), fn
        D = m.distMatrix().to_host()
        if str(z["kind"]) == "ref_mash":
            # produced by the reference's own CUDA objects (tools/make_ref_golden.py): rows, k-closest and exact-mode trees
            n = len(lens)
            low = np.tril_indices(n, -1)
            assert np.allclose(D[low], z["rows"][low], rtol=1e-6, atol=0), fn
            prm = api.Param(kmerSize=int(z["k"]), sketchSize=int(z["s"]), in_="r")
            kp = api.KPlacementDeviceArrays(ctx); kp.allocateDeviceArrays(n)
            kp.findPlacementTree(prm, mashDeviceArrays=m)
            assert kp.printTree(synth.names(n)) == str(z["place_newick"]), fn
            pl = api.PlacementDeviceArrays(ctx); pl.allocateDeviceArrays(n)
            pl.findPlacementTree(prm, mashDeviceArrays=m)
            assert pl.printTree(synth.names(n)) == str(z["place_exact_newick"]), fn
        else:
            assert np.allclose(D, z["dist"], rtol=1e-6, atol=0), fn


def test_mash_to_nj_tree(ctx, oracle):
    """-i r -o t -m 2 (config C2): sketches -> Mash matrix -> NJ, all on the GPU."""
    seqs, packed, lens = unaligned_case(120, 5000, seed=11)
    m = upload(ctx, packed, lens)
    nj = api.NJDeviceArrays(ctx)
    nj.getDismatrix(120, api.Param(in_="r"), mashDeviceArrays=m)
    D = nj.matrix.to_host()
    nwk = nj.findNeighbourJoiningTree(synth.names(120))
    o = oracle.nj(D)
    assert nwk == oracle.nj_newick(*o, synth.names(120))


def test_distances_before_sketching_fail_loudly(ctx):
    seqs, packed, lens = unaligned_case(4, 500, seed=1)
    m = api.MashDeviceArrays(ctx)
    m.allocateDeviceArrays(packed, lens, 4, api.Param(in_="r"))
    with pytest.raises(api.DipperError):
        m.distMatrix()
