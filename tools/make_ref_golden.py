"""Golden vectors from the reference's OWN CUDA objects (oracle/_ref/dipper_ref, built by oracle/build_ref.sh from the
reference sources in place).  Runs on a GPU box:  python tools/make_ref_golden.py   -> gpurun_out/ref_*.npz, which are
then committed under tests/golden/ and checked by the CPU-only suite (tests/test_oracle.py::test_golden_fixtures).
Inputs are seeded synthetic data (dipper_b200/synth.py); every output array / tree text is the reference's."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from dipper_b200 import synth

REF = os.path.join(ROOT, "oracle", "_ref", "dipper_ref")
OUT = os.path.join(ROOT, "gpurun_out")
TMP = "/tmp/ref_golden"
os.makedirs(TMP, exist_ok=True)
os.makedirs(OUT, exist_ok=True)


def write_bin(path, rows, lens, bits):
    with open(path, "wb") as f:
        np.array([len(lens), bits], np.int64).tofile(f)
        np.asarray(lens, np.uint64).tofile(f)
        for r in rows:
            np.ascontiguousarray(r, np.uint64).tofile(f)


def run(mode, inp, out, *extra):
    p = subprocess.run([REF, mode, inp, out, *map(str, extra)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return json.loads(p.stdout.strip().splitlines()[-1])


def rows_to_lower(flat, n):
    D = np.zeros((n, n))
    k = 0
    for i in range(1, n):
        D[i, :i] = flat[k:k + i]
        k += i
    return D


# ---- aligned: 60 tips x 1500 sites, tie-free regime
n, L = 60, 1500
codes, _ = synth.evolve(n, L, seed=11, regime="tiefree", gap_cols=0.05)
P = synth.pack4_np(codes)
inp, out = TMP + "/msa.bin", TMP + "/msa"
write_bin(inp, P, [L] * n, 4)
rec = dict(kind="ref_msa", packed=P, seq_len=L)
for t in (1, 2):
    run("msa_rows", inp, out, t)
    rec["rows_%d" % t] = rows_to_lower(np.fromfile(out + ".rows", np.float64), n)
run("msa_nj", inp, out, 2); rec["nj_newick"] = open(out + ".nwk").read()
run("msa_place", inp, out, 2); rec["place_newick"] = open(out + ".nwk").read()
run("msa_place_exact", inp, out, 2); rec["place_exact_newick"] = open(out + ".nwk").read()
np.savez_compressed(os.path.join(OUT, "ref_msa_60x1500.npz"), **rec)

# ---- unaligned: 24 sequences of ~3 kb
n, L = 24, 3000
codes, _ = synth.evolve(n, L, seed=12, regime="tiefree", gap_cols=0.01)
seqs = synth.unaligned(codes)
flat, offs, lens = synth.flatten2(seqs)
inp, out = TMP + "/mash.bin", TMP + "/mash"
write_bin(inp, [synth.pack2_np(s) for s in seqs], lens, 2)
rec = dict(kind="ref_mash", flat=flat, offsets=offs, lens=lens, k=15, s=1000)
run("mash_sketch", inp, out, 1, 15); rec["sketches"] = np.fromfile(out + ".sk", np.uint64).reshape(n, 1000)
run("mash_rows", inp, out, 1, 15); rec["rows"] = rows_to_lower(np.fromfile(out + ".rows", np.float64), n)
run("mash_place", inp, out, 1, 15); rec["place_newick"] = open(out + ".nwk").read()
run("mash_place_exact", inp, out, 1, 15); rec["place_exact_newick"] = open(out + ".nwk").read()
np.savez_compressed(os.path.join(OUT, "ref_mash_24x3000.npz"), **rec)
print("wrote", sorted(f for f in os.listdir(OUT) if f.startswith("ref_")))
