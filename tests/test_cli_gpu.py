"""GPU end-to-end: the `dipper` binary (reference CLI contract) vs the library API / oracle."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from dipper_b200 import api, newick, synth
from conftest import make_msa, ROOT

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "dipper_b200", "dipper")


def run(*args):
    p = subprocess.run([EXE, *map(str, args)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stderr


def test_aligned_fasta_to_nj_tree(ctx, oracle, tmp_path):
    n, L = 150, 2000
    codes, P, _ = make_msa(n, L, seed=61)
    names = synth.names(n)
    fa, out = str(tmp_path / "a.fa.gz"), str(tmp_path / "o.nwk")
    with gzip.open(fa, "wt") as f:
        for nm, s in zip(names, synth.codes_to_strings(codes)):
            f.write(">%s some description\n%s\n%s\n" % (nm, s[:700], s[700:]))     # wrapped lines, header comment
    err = run("-i", "m", "-I", fa, "-O", out, "-o", "t", "-d", 2, "-m", 2, "--no-shuffle")
    assert "Using conventional NJ" in err
    o = oracle.nj(oracle.msa_dist_matrix(P, L, 2))
    exp = oracle.nj_newick(*o, names)
    got = open(out).read()
    assert newick.rf_distance(got, exp) == 0 and newick.max_branch_diff(got, exp) < 1e-5
    # a seeded shuffle changes row order, not the leaf set; default -d is 1 (uncorrected) like the reference
    run("-i", "m", "-I", fa, "-O", out, "-m", 2, "--seed", 7)
    got2 = open(out).read()
    assert sorted(newick.parse(got2)[2]) == sorted(newick.parse(exp)[2])


def test_unaligned_fasta_placement_and_dc(ctx, oracle, tmp_path):
    n = 120
    codes, _ = synth.evolve(n, 2500, seed=62, gap_cols=0.01)
    seqs = synth.unaligned(codes)
    names = synth.names(n)
    fa, out = str(tmp_path / "u.fa"), str(tmp_path / "o.nwk")
    lut = np.frombuffer(b"ACGT", np.uint8)
    synth.write_fasta(fa, names, [lut[s].tobytes().decode() for s in seqs])
    prm = api.Param(kmerSize=15, sketchSize=1000, in_="r")
    m = api.MashDeviceArrays(ctx)
    m.allocateDeviceArrays([synth.pack2_np(s) for s in seqs], np.array([len(s) for s in seqs], np.uint64), n, prm)
    m.sketchConstructionOnGpu()
    D = m.distMatrix().to_host()
    err = run("-i", "r", "-I", fa, "-O", out, "-m", 1, "--no-shuffle")
    assert "k-closest placement mode" in err
    assert open(out).read() == oracle.place_all(D).newick(names)
    err = run("-i", "r", "-I", fa, "-O", out, "-m", 1, "-p", 0, "--no-shuffle")
    assert "exact placement mode" in err
    assert open(out).read() == oracle.place_exact(D).newick(names)
    err = run("-i", "r", "-I", fa, "-O", out, "-m", 3, "--no-shuffle")
    assert "divide-and-conquer" in err
    assert open(out).read() == oracle.dc(D, n // 20)[0].newick(names)
    run("-i", "r", "-I", fa, "-O", out, "-m", 2, "--no-shuffle")          # Mash + NJ (reference bug B3 fixed)
    o = oracle.nj(D)
    assert open(out).read() == oracle.nj_newick(*o, names)


def test_phylip_input_and_add_tips(ctx, oracle, tmp_path):
    n, B, L = 90, 40, 1500
    codes, P, _ = make_msa(n, L, seed=63)
    names = synth.names(n)
    D = oracle.msa_dist_matrix(P, L, 1)
    phy, out = str(tmp_path / "m.phy"), str(tmp_path / "o.nwk")
    synth.write_phylip(phy, names, D, lower=True)
    run("-i", "d", "-I", phy, "-O", out, "-m", 2)
    D32 = np.array([[np.float32("%.6f" % v) for v in row] for row in D], np.float64)
    Dq = np.tril(D32, -1) + np.tril(D32, -1).T
    assert open(out).read() == oracle.nj_newick(*oracle.nj(Dq), names)
    # --add: backbone = placement tree of the first B tips, written as Newick, then all sequences as input
    bb = oracle.place_all(np.ascontiguousarray(D[:B, :B])).newick(names[:B])
    tree, fa = str(tmp_path / "bb.nwk"), str(tmp_path / "all.fa")
    open(tree, "w").write(bb)
    synth.write_fasta(fa, names, synth.codes_to_strings(codes))
    err = run("-i", "m", "-I", fa, "-O", out, "-d", 1, "--add", "-t", tree)
    got = open(out).read()
    leaves = sorted(x for x in newick.parse(got)[2] if x)
    assert leaves == sorted(names)
    # the backbone's topology is preserved among backbone tips
    kp = api.KPlacementDeviceArrays(ctx)
    kp.allocateDeviceArrays(n)
    kp.initializeDeviceArrays(bb)
    order = [names.index(x) for x in kp.backbone_names] + list(range(B, n))
    msa = api.MSADeviceArrays(ctx)
    prm = api.Param(distanceType=1, in_="m")
    msa.allocateDeviceArrays(np.ascontiguousarray(P[order]), np.full(n, L, np.uint64), n, prm)
    kp.addQuery(prm, msaDeviceArrays=msa)
    assert got == kp.printTree([names[i] for i in order])


def test_cli_errors(tmp_path):
    p = subprocess.run([EXE, "-i", "m"], capture_output=True, text=True)
    assert p.returncode == 1 and "required" in p.stderr
    p = subprocess.run([EXE, "-i", "m", "-I", "/nonexistent.fa", "-O", str(tmp_path / "x")], capture_output=True, text=True)
    assert p.returncode == 1 and "cant open file" in p.stderr
    p = subprocess.run([EXE, "-h"], capture_output=True, text=True)
    assert p.returncode == 0 and "--input-format" in p.stderr


def test_output_format_d_writes_a_matrix_the_cli_reads_back(ctx, oracle, tmp_path):
    """-o d (docs/index.md:114, not implemented by the reference): PHYLIP out, then -i d in -> the same tree as -i m."""
    n, L = 90, 1500
    codes, P, _ = make_msa(n, L, seed=66)
    names = synth.names(n)
    fa, phy, t1, t2 = (str(tmp_path / x) for x in ("a.fa", "m.phy", "a.nwk", "b.nwk"))
    synth.write_fasta(fa, names, synth.codes_to_strings(codes))
    run("-i", "m", "-I", fa, "-O", phy, "-o", "d", "-d", 2, "--no-shuffle")
    lines = open(phy).read().splitlines()
    assert int(lines[0]) == n and [ln.split()[0] for ln in lines[1:]] == names
    D = oracle.msa_dist_matrix(P, L, 2)
    row = np.array([float(t) for t in lines[1 + 40].split()[1:]])
    assert len(row) == 40 and np.allclose(row, D[40, :40], rtol=1e-6, atol=0)
    run("-i", "m", "-I", fa, "-O", t1, "-d", 2, "-m", 2, "--no-shuffle")
    run("-i", "d", "-I", phy, "-O", t2, "-m", 2)
    a, b = open(t1).read(), open(t2).read()
    assert newick.rf_distance(a, b) == 0 and newick.max_branch_diff(a, b) < 1e-5     # (values pass through float32 on the way back)
