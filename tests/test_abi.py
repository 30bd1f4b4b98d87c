"""CPU tests of the C-ABI boundary: the library loads, exports every declared symbol,
fails loudly without a GPU, and its host-side code agrees with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from dipper_b200 import _lib, api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for h in ("dipper_b200.h", "dipper_host.h"):
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names |= set(re.findall(r"\b(dipb_[a-z0-9_]+)\s*\(", txt))
    return names


def test_library_exports_every_declared_symbol():
    L = C.CDLL(_lib.LIB_PATH)
    decl = _declared()
    assert len(decl) >= 40
    for nm in decl:
        assert hasattr(L, nm), "missing export " + nm
    assert decl == set(_lib.SIGNATURES), decl ^ set(_lib.SIGNATURES)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.DipperError) as e:
        api.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_host_packers_match_oracle(oracle):
    rng = np.random.default_rng(0)
    for L in (1, 15, 16, 17, 31, 32, 33, 64, 1000):
        s = "".join(rng.choice(list("ACGTUN-acgtRY"), L))
        assert np.array_equal(api.pack4(s), oracle.pack4(s))
        assert np.array_equal(api.pack2(s), oracle.pack2(s))


def test_nj_newick_writer_matches_oracle(oracle):
    codes, _ = synth.evolve(30, 500, seed=4)
    D = oracle.msa_dist_matrix(synth.pack4_np(codes), 500, 2)
    c0, c1, l0, l1 = oracle.nj(D)
    names = synth.names(30)
    mine = _lib.take_str(_lib.lib().dipb_nj_newick(30, c0, c1, l0, l1, _lib.names_array(names)))
    assert mine == oracle.nj_newick(c0, c1, l0, l1, names)


def test_tree_newick_writer_matches_oracle(oracle):
    codes, _ = synth.evolve(25, 500, seed=5)
    D = oracle.msa_dist_matrix(synth.pack4_np(codes), 500, 2)
    t = oracle.place_all(D)
    names = synth.names(25)
    a = t.arrays()
    padded = names + [""] * 25
    mine = _lib.take_str(_lib.lib().dipb_tree_newick(50, 25, a["head"], a["e"], a["nxt"], a["len"],
                                                     _lib.names_array(padded)))
    assert mine == t.newick(names)


def test_backbone_loader_round_trip(oracle):
    """Newick -> adjacency arrays (src/placement_close_k.cu:160-183) -> Newick."""
    nwk = "((A:0.1,B:0.2):0.05,(C:0.3,(D:0.15,E:0.25):0.02):0.07);"
    n = 8  # 5 backbone leaves + 3 queries to come
    kp = api.KPlacementDeviceArrays(None)
    kp.allocateDeviceArrays(n)
    B = kp.initializeDeviceArrays(nwk)
    assert B == 5 and kp.backbone_names == ["A", "B", "C", "D", "E"]
    head, e, nxt, belong, ln = kp._backbone
    assert (belong[: 4 * B - 4] >= 0).all() and (belong[4 * B - 4:] == -1).all()
    # leaves have exactly one adjacency, the root (idx n) two, other internals three
    deg = np.bincount(belong[: 4 * B - 4], minlength=2 * n)
    assert list(deg[:B]) == [1] * B and deg[n] == 2 and set(deg[n + 1: n + B - 1]) == {3}
    # float32 parse like the reference's stof
    assert ln[0] == float(np.float32(0.1))
    names = kp.backbone_names + [""] * (2 * n - B)
    out = _lib.take_str(_lib.lib().dipb_tree_newick(2 * n, n, head, e, nxt, ln, _lib.names_array(names)))
    from dipper_b200 import newick
    assert newick.rf_distance(out, nwk) == 0
    assert newick.max_branch_diff(out, nwk) < 1e-6


def test_reference_backbone_file_parses():
    path = "/root/reference/dataset/t2.backbone.nwk"
    if not os.path.exists(path):
        pytest.skip("reference not mounted")
    nwk = open(path).readline()
    kp = api.KPlacementDeviceArrays(None)
    kp.allocateDeviceArrays(10000)
    B = kp.initializeDeviceArrays(nwk)
    # 1000 leaves, first leaf T9326 idx 0, T342 idx 1
    assert B == 1000 and kp.backbone_names[0] == "T9326" and kp.backbone_names[1] == "T342"


def test_backbone_loader_rejects_non_binary_trees():
    """dipb_place_add walks exactly 4B-4 slots (rooted binary backbone, src/placement_close_k.cu:887):
    an unrooted (trifurcating root) or unary-node Newick must be refused, not walked out of bounds."""
    for bad in ("(A:0.1,B:0.2,(C:0.3,D:0.1):0.2);", "((A:0.1,B:0.2):0.3);"):
        kp = api.KPlacementDeviceArrays(None)
        kp.allocateDeviceArrays(8)
        with pytest.raises(api.DipperError) as e:
            kp.initializeDeviceArrays(bad)
        assert "rooted binary" in str(e.value)


def test_phylip_writer_round_trips_through_the_reference_reader_rule(tmp_path):
    """-o d (docs/index.md:114): what dipb_phylip_write emits is what MatrixReader (src/matrix_reader.cu:15-44) reads:
    n, then `name v0 v1 ...` rows whose values survive a float32 parse."""
    rng = np.random.default_rng(7)
    n = 9
    D = rng.random((n, n))
    D = np.tril(D, -1) + np.tril(D, -1).T
    names = ["tx%d" % i for i in range(n)]
    for lower in (1, 0):
        path = str(tmp_path / ("m%d.phy" % lower))
        rc = _lib.lib().dipb_phylip_write(path.encode(), n, np.ascontiguousarray(D), _lib.names_array(names), lower)
        assert rc == 0
        lines = open(path).read().splitlines()
        assert int(lines[0]) == n and len(lines) == n + 1
        for i in range(n):
            tok = lines[1 + i].split()
            assert tok[0] == names[i] and len(tok) == 1 + (i if lower else n)
            got = np.array([np.float32(t) for t in tok[1:1 + i]], np.float32)
            assert np.array_equal(got, D[i, :i].astype(np.float32))


def test_branch_length_formatter_equals_printf_g():
    """dipb_format_g (the Newick writers' number formatter) against C's "%g" (Python's % operator formats the same way)."""
    import ctypes as C
    import math
    from dipper_b200._lib import lib
    rng = np.random.default_rng(11)
    vals = [0.0, -0.0, 1.0, 10.0, 0.1, 0.5, 0.0078125, 999999.0, 999999.5, 999999.4999, 1e6, 123456.5, 1e-4, 1e-5, 9.9999949e-5,
            9.99999951e-5, 0.000099999949, 1.5e-10, 1e-15, 9e-16, 1e-300, 5e-324, 1e300, float("inf"), float("-inf"), 2.5e-7, -3.75e-3]
    vals += list(rng.random(60000) * 2.0)                                     # branch-length sized
    vals += list(np.ldexp(0.5 + rng.random(60000) * 0.5, rng.integers(-70, 30, 60000)))
    vals += list(-np.ldexp(0.5 + rng.random(5000) * 0.5, rng.integers(-40, 10, 5000)))
    vals += [k / 2.0 ** j for j in range(0, 30, 3) for k in range(1, 3000, 7)]     # exact decimal ties
    vals += [(999990 + d * 0.25) * 10.0 ** x for x in range(-18, 2) for d in range(0, 44)]   # carries into the next decade
    buf = C.create_string_buffer(40)
    f = lib().dipb_format_g
    for v in vals:
        n = f(float(v), buf)
        assert buf.value.decode() == "%g" % v and n == len(buf.value), (v, buf.value)
    n = f(float("nan"), buf)
    assert "nan" in buf.value.decode()
