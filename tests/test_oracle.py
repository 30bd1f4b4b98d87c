"""CPU tests: the oracle against known answers, golden fixtures and its own invariants."""
import glob
import os

import numpy as np
import pytest

from dipper_b200 import newick, synth
from conftest import make_msa

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_murmur_known_answers(oracle):
    # public MurmurHash3_x64_128 vectors
    assert oracle.murmur3_x64_128(b"", 0) == (0, 0)
    assert oracle.murmur3_x64_128(b"hello", 0) == (0xCBD8A7B341BD9B02, 0x5B1E906A48AE1D19)
    assert oracle.murmur3_x64_128(b"The quick brown fox jumps over the lazy dog", 0) == (
        0xE34BBC7BBC071B6C, 0x7A433CA9C49A9347)


def test_murmur_smhasher_verification(oracle):
    # SMHasher VerificationTest: keys {0..i-1} with seed 256-i, hash of all digests, first 32 bits
    blob = b""
    for i in range(256):
        h = oracle.murmur3_x64_128(bytes(range(i)), 256 - i)
        blob += int(h[0]).to_bytes(8, "little") + int(h[1]).to_bytes(8, "little")
    final = oracle.murmur3_x64_128(blob, 0)
    assert final[0] & 0xFFFFFFFF == 0x6384BA69


def test_pack4_pack2_layout(oracle):
    s = "ACGTUNacgt-RYK" * 5 + "A"
    p4 = oracle.pack4(s)
    p2 = oracle.pack2(s)
    lut4 = {"A": 0, "C": 1, "G": 2, "T": 3, "U": 3}
    for j, ch in enumerate(s):
        assert (int(p4[j // 16]) >> (4 * (j % 16))) & 15 == lut4.get(ch, 4)
        assert (int(p2[j // 32]) >> (2 * (j % 32))) & 3 == lut4.get(ch, 0)
    codes = np.array([lut4.get(ch, 4) for ch in s], np.uint8)
    assert np.array_equal(synth.pack4_np(codes[None, :])[0], p4)
    assert np.array_equal(synth.pack2_np(codes), p2)


@pytest.mark.parametrize("L", [1, 15, 16, 17, 33, 1000])
def test_fast_counts_equals_nibble_loop(oracle, L):
    import ctypes as C
    codes, P, _ = make_msa(6, L, seed=L)
    for i in range(6):
        for j in range(i):
            st = oracle.pair_stats(P[i], P[j], L)
            u, m = C.c_int(), C.c_int()
            oracle.lib().orc_fast_counts(P[i], P[j], L, C.byref(u), C.byref(m))
            assert (u.value, m.value) == (st.useful, st.match)
            a, b = codes[i], codes[j]
            assert st.useful == int(((a < 4) | (b < 4)).sum())
            assert st.match == int(((a < 4) & (a == b)).sum())
            assert st.tot == int(((a < 4) & (b < 4)).sum())
            assert st.ts + st.tv + st.match == st.tot


def test_distance_models_formulas(oracle):
    codes, P, _ = make_msa(8, 3000, seed=3, gap_cols=0.05)
    st = oracle.pair_stats(P[5], P[2], 3000)
    p = 1 - st.match / st.useful
    assert oracle.dist_from_stats(st, 1) == p
    assert oracle.dist_from_stats(st, 2) == -0.75 * np.log(1.0 - p / 0.75)
    pp, qq = st.ts / st.tot, st.tv / st.tot
    assert np.isclose(oracle.dist_from_stats(st, 4), -0.5 * np.log((1 - 2 * pp - qq) * np.sqrt(1 - 2 * qq)), rtol=1e-14)
    assert np.isclose(oracle.dist_from_stats(st, 6), 0.5 * (1 / (1 - 2 * pp - qq) + 0.5 / (1 - 2 * qq) - 1.5), rtol=1e-14)
    for t in (3, 5):
        assert np.isfinite(oracle.dist_from_stats(st, t))
    D = oracle.msa_dist_matrix(P, 3000, 2)
    assert np.array_equal(D, D.T) and np.all(np.diag(D) == 0)
    assert np.array_equal(oracle.msa_dist_row(P, 3000, 6, 2), D[6, :6])


def test_sketch_semantics(oracle):
    rng = np.random.default_rng(5)
    seq = rng.integers(0, 4, 700)
    p2 = synth.pack2_np(seq)
    sk = oracle.sketch(p2, 700, k=15, s=1000)
    nk = 700 - 15 + 1
    hashes = sorted(int(oracle.lib().orc_kmer_hash(p2, j, 15)) for j in range(nk))
    assert [int(x) for x in sk[:nk]] == hashes          # multiset kept, ascending
    assert np.all(sk[nk:] == np.uint64(0xFFFFFFFFFFFFFFFF))  # padded
    # canonical: reverse complement gives the same sketch
    rc = (3 - seq)[::-1]
    assert np.array_equal(oracle.sketch(synth.pack2_np(rc), 700, 15, 1000), sk)
    # bottom-s of a longer sequence
    seq2 = rng.integers(0, 4, 5000)
    sk2 = oracle.sketch(synth.pack2_np(seq2), 5000, 15, 1000)
    assert np.all(sk2[:-1] <= sk2[1:]) and sk2[-1] != np.uint64(0xFFFFFFFFFFFFFFFF)


def test_mash_distance_properties(oracle):
    rng = np.random.default_rng(7)
    a = np.sort(rng.integers(0, 2**63, 1000, dtype=np.uint64))
    assert oracle.mash_dist(a, a, 15) == 0.0 or oracle.mash_inter_uni(a, a) == (1000, 1000)
    i, u = oracle.mash_inter_uni(a, a)
    assert (i, u) == (1000, 1000)
    b = np.sort(rng.integers(0, 2**63, 1000, dtype=np.uint64))
    i, u = oracle.mash_inter_uni(a, b)
    assert u == 1000 and i == 0
    assert oracle.mash_dist(a, b, 15) == min(1.0, abs(np.log(2 * (1 / 1000) / (1 + 1 / 1000)) / 15))
    # shared half
    c = np.sort(np.concatenate([a[:500], b[:500]]))
    i, u = oracle.mash_inter_uni(a, c)
    merged = np.sort(np.concatenate([a, c]))
    assert u == 1000 and 0 < i <= 500


def test_canonical_sum_is_exact_on_integers(oracle):
    v = np.arange(5000, dtype=np.float64)
    assert oracle.canon_sum(v) == v.sum()
    r = np.random.default_rng(1).random(3000)
    assert abs(oracle.canon_sum(r) - r.sum()) < 1e-9


def _true_newick(info, names):
    ch, bl, order = info["children"], info["bl"], info["order"]
    row = {int(v): i for i, v in enumerate(order)}
    out = {}
    stack = [(0, False)]
    while stack:
        v, done = stack.pop()
        if not ch[v]:
            out[v] = names[row[v]]
        elif done:
            out[v] = "(" + ",".join("%s:%.10g" % (out[c], bl[c]) for c in ch[v]) + ")"
        else:
            stack.append((v, True))
            stack.extend((c, False) for c in ch[v])
    return out[0] + ";"


def _additive_matrix(info, n):
    """Exact path-length distances of the generating tree (NJ must recover it: RF = 0)."""
    parent, bl, order = info["parent"], info["bl"], info["order"]
    depth = {}
    def anc(v):
        path = []
        while v != -1:
            path.append(v)
            v = int(parent[v])
        return path
    paths = [anc(int(v)) for v in order]
    D = np.zeros((n, n))
    for i in range(n):
        di = {v: 0.0 for v in []}
        acc = 0.0
        for v in paths[i]:
            di[v] = acc
            acc += bl[v]
        for j in range(i):
            accj = 0.0
            for v in paths[j]:
                if v in di:
                    D[i, j] = D[j, i] = di[v] + accj
                    break
                accj += bl[v]
    return D


def test_nj_recovers_additive_tree(oracle):
    from dipper_b200 import synth
    n = 60
    rng = np.random.default_rng(11)
    tree = synth.yule_tree(n, rng, "tiefree")
    codes, info = synth.evolve(n, 10, seed=2, tree=tree)
    names = synth.names(n)
    D = _additive_matrix(info, n)
    c0, c1, l0, l1 = oracle.nj(D)
    nw = oracle.nj_newick(c0, c1, l0, l1, names)
    assert newick.rf_distance(nw, _true_newick(info, names)) == 0
    assert newick.max_branch_diff(nw, _true_newick(info, names)) < 1e-6  # %g prints 6 significant digits


def test_placement_recovers_additive_tree(oracle):
    from dipper_b200 import synth
    n = 40
    rng = np.random.default_rng(12)
    tree = synth.yule_tree(n, rng, "tiefree")
    codes, info = synth.evolve(n, 10, seed=3, tree=tree)
    names = synth.names(n)
    D = _additive_matrix(info, n)
    t = oracle.place_all(D)
    nw = t.newick(names)
    assert newick.rf_distance(nw, _true_newick(info, names)) == 0


def test_nj_small_cases(oracle):
    D = np.array([[0, 3.0], [3.0, 0]])
    c0, c1, l0, l1 = oracle.nj(D)
    assert (c0[0], c1[0], l0[0], l1[0]) == (0, 1, 1.5, 1.5)
    D = np.array([[0, 2.0, 4.0], [2.0, 0, 4.0], [4.0, 4.0, 0]])
    c0, c1, l0, l1 = oracle.nj(D)
    nw = oracle.nj_newick(c0, c1, l0, l1, ["a", "b", "c"])
    assert nw.endswith(";\n") and nw.count("(") == 2


def test_golden_fixtures(oracle):
    """Oracle reproduces every committed fixture (tests/golden/make_golden.py wrote them;
    fixtures named ref_* were produced by the reference's own CUDA objects on a B200)."""
    files = sorted(glob.glob(os.path.join(GOLD, "*.npz")))
    assert files, "no golden fixtures committed"
    for fn in files:
        z = np.load(fn)
        kind = str(z["kind"])
        if kind == "msa":
            P, L = z["packed"], int(z["seq_len"])
            m, u = oracle.msa_counts(P, L, 0, P.shape[0], 0, P.shape[0])
            assert np.array_equal(m, z["match"]) and np.array_equal(u, z["useful"]), fn
            for t in z["dist_types"]:
                D = oracle.msa_dist_matrix(P, L, int(t))
                G = z["dist_%d" % t]
                ok = np.isfinite(G)
                assert np.allclose(D[ok], G[ok], rtol=1e-6, atol=0), fn
            if "nj_child0" in z:
                c0, c1, l0, l1 = oracle.nj(z["dist_2"])
                assert np.array_equal(c0, z["nj_child0"]) and np.array_equal(c1, z["nj_child1"]), fn
                assert np.allclose(l0, z["nj_len0"], atol=1e-5) and np.allclose(l1, z["nj_len1"], atol=1e-5), fn
        elif kind == "mash":
            sk = oracle.sketch_all(z["flat"], z["offsets"], z["lens"], int(z["k"]), int(z["s"]))
            assert np.array_equal(sk, z["sketches"]), fn
            D = oracle.mash_dist_matrix(sk, int(z["k"]))
            assert np.allclose(D, z["dist"], rtol=1e-6, atol=0), fn
        elif kind == "ref_msa":
            # outputs of the reference's own CUDA kernels (tools/make_ref_golden.py on a B200)
            P, L = z["packed"], int(z["seq_len"])
            n = P.shape[0]
            names = synth.names(n)
            low = np.tril_indices(n, -1)
            for t in (1, 2):
                D = oracle.msa_dist_matrix(P, L, t)
                assert np.allclose(D[low], z["rows_%d" % t][low], rtol=1e-6, atol=0), fn
            D = oracle.msa_dist_matrix(P, L, 2)
            nj = oracle.nj_newick(*oracle.nj(D), names)
            assert newick.rf_distance(nj, str(z["nj_newick"])) == 0 and newick.max_branch_diff(nj, str(z["nj_newick"])) < 1e-5, fn
            assert oracle.place_all(D).newick(names) == str(z["place_newick"]), fn
            assert oracle.place_exact(D).newick(names) == str(z["place_exact_newick"]), fn
        elif kind == "ref_mash":
            sk = oracle.sketch_all(z["flat"], z["offsets"], z["lens"], int(z["k"]), int(z["s"]))
            assert np.array_equal(sk, z["sketches"]), fn
            n = sk.shape[0]
            names = synth.names(n)
            low = np.tril_indices(n, -1)
            D = oracle.mash_dist_matrix(sk, int(z["k"]))
            assert np.allclose(D[low], z["rows"][low], rtol=1e-6, atol=0), fn
            assert oracle.place_all(D).newick(names) == str(z["place_newick"]), fn
            assert oracle.place_exact(D).newick(names) == str(z["place_exact_newick"]), fn
        elif kind == "ref_models":
            # the six distance models, rows from the reference's well-formed DC twins (src/divide_and_conquer/msa.cu:219-264)
            P, L = z["packed"], int(z["seq_len"])
            low = np.tril_indices(P.shape[0], -1)
            for t in range(1, 7):
                D = oracle.msa_dist_matrix(P, L, t)
                assert np.allclose(D[low], z["rows_%d" % t][low], rtol=1e-6, atol=0), (fn, t)
        elif kind == "ref_dc_msa":
            # aligned D&C as the reference ships it: defect B17 on (see orc_dc_matrix_as_shipped) -> identical output
            P, L, B = z["packed"], int(z["seq_len"]), int(z["backbone"])
            n = P.shape[0]
            D = oracle.msa_dist_matrix(P, L, 2)
            t, cl = oracle.dc_as_shipped(D, B, 0.0)
            assert np.array_equal(cl, z["clusters"]), fn
            _same_slot_arrays(t.arrays(), z, n, fn)
            assert t.newick(synth.names(n)) == str(z["newick"]), fn
            # the intended rule (what the Mash twin does and the product follows) differs from the shipped one only in
            # queries whose candidate lists see backbone tip B-1
            _, cl2 = oracle.dc(D, B)
            assert 0 < np.count_nonzero(cl2 != cl) < 0.05 * (n - B), fn
        elif kind == "ref_dc_mash":
            B = int(z["backbone"])
            sk = oracle.sketch_all(z["flat"], z["offsets"], z["lens"], int(z["k"]), int(z["s"]))
            n = sk.shape[0]
            t, cl = oracle.dc(oracle.mash_dist_matrix(sk, int(z["k"])), B)
            assert np.array_equal(cl, z["clusters"]), fn
            _same_slot_arrays(t.arrays(), z, n, fn)
            assert t.newick(synth.names(n)) == str(z["newick"]), fn
        elif kind in ("ref_add_msa", "ref_add_t2"):
            # add-tips (initializeDeviceArrays(Tree*) + addQuery, src/placement_close_k.cu:126-264,858-990)
            from dipper_b200 import api
            P, L, B = z["packed"], int(z["seq_len"]), int(z["backbone"])
            n = P.shape[0]
            bb = str(z["backbone_newick"]) if "backbone_newick" in z else open(os.path.join(GOLD, "t2.backbone.nwk")).readline().strip()
            kp = api.KPlacementDeviceArrays(None)
            kp.allocateDeviceArrays(n)
            assert kp.initializeDeviceArrays(bb) == B
            root, off, flat, parent, bl = _backbone_tables(bb, n)
            t = oracle.place_add(oracle.msa_dist_matrix(P, L, 2), B, root, off, flat, parent, bl)
            _same_slot_arrays(t.arrays(), z, n, fn)
            names = kp.backbone_names + ["Q%d" % (i + 1) for i in range(n - B)]
            assert t.newick(names) == str(z["newick"]), fn


def _same_slot_arrays(a, z, n, fn):
    ns = 4 * n - 4
    assert np.array_equal(a["head"][: 2 * n], z["head"][: 2 * n]), fn
    for k in ("e", "nxt", "belong"):
        assert np.array_equal(a[k][:ns], z[k][:ns]), (fn, k)
    assert np.allclose(a["len"][:ns], z["len"][:ns], rtol=0, atol=1e-12), fn   # libm vs libdevice log: <= 1 ulp of a distance


def _backbone_tables(nwk, total):
    """Newick -> the node tables orc_ptree_load_backbone takes (leaf ids by order of appearance, internal ids
    total, total+1, ... in order of '(', src/tree.cpp:308-341)."""
    children, length, name = newick.parse(nwk)
    nn = len(children)
    idx = [0] * nn
    leaf = 0
    k = 0
    # parse() numbers nodes in order of appearance: node 0 = root = first '(', internal nodes in order of '('
    for v in range(nn):
        if children[v]:
            idx[v] = total + k
            k += 1
        else:
            idx[v] = leaf
            leaf += 1
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
    return idx[0], off, np.array(flat if flat else [0], np.int32), parent, bl


def test_exact_placement_oracle_recovers_an_additive_tree(oracle):
    """Exact placement mode (src/placement.cu) on the path metric of a random tree returns that tree."""
    from dipper_b200 import newick as nw
    n, rng = 120, np.random.default_rng(5)
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
    names = ["T%d" % (i + 1) for i in range(n)]

    def nwk(v, p):
        ch = [w for w in adj[v] if w != p]
        return names[v] if not ch else "(" + ",".join(nwk(w, v) + ":%g" % adj[v][w] for w in ch) + ")"
    truth = nwk(n, -1) + ";"
    mine = oracle.place_exact(D).newick(names)
    assert nw.rf_distance(mine, truth) == 0
    assert nw.max_branch_diff(mine, truth) < 1e-6   # %g text: 6 significant digits
    # the k-closest rule agrees on an additive metric
    assert nw.rf_distance(oracle.place_all(D).newick(names), truth) == 0
