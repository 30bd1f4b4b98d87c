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
    assert p.returncode == 0, (mode, p.returncode, p.stderr[-1500:])
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


def read_arrays(path, n):
    raw = np.fromfile(path, np.uint8)
    o = 0
    def take(cnt, dt):
        nonlocal o
        a = raw[o:o + cnt * np.dtype(dt).itemsize].view(dt).copy()
        o += cnt * np.dtype(dt).itemsize
        return a
    return dict(head=take(2 * n, np.int32), e=take(8 * n, np.int32), nxt=take(8 * n, np.int32), belong=take(8 * n, np.int32),
                len=take(8 * n, np.float64))


def singleton_at_B(cl, B):
    """Reference defect B12 (src/divide_and_conquer/msa.cu:395-397, mash.cu:584-588: `idx > backboneSize` should be
    `>=`) makes every later tip of the cluster that holds tip B read tip B's data from the wrong buffer.  Cluster
    assignment is independent per query, so the inputs are arranged with a query that is ALONE in its cluster at
    index B; the defect then has nothing to act on and the outputs must agree exactly."""
    cnt = np.bincount(cl[B:], minlength=int(cl.max()) + 1)
    for q in range(B, len(cl)):
        if cnt[cl[q]] == 1:
            return q
    raise RuntimeError("no singleton cluster")


from oracle import oracle as ORC   # here only to arrange the input (see singleton_at_B)

# ---- the six distance models through the well-formed DC twins (src/divide_and_conquer/msa.cu:219-264)
n, L = 48, 1500
codes, _ = synth.evolve(n, L, seed=13, regime="tiefree", gap_cols=0.05)
P = synth.pack4_np(codes)
inp, out = TMP + "/models.bin", TMP + "/models"
write_bin(inp, P, [L] * n, 4)
rec = dict(kind="ref_models", packed=P, seq_len=L)
for t in (1, 2, 3, 4, 5, 6):
    run("msa_dc_rows", inp, out, t)
    rec["rows_%d" % t] = rows_to_lower(np.fromfile(out + ".rows", np.float64), n)
np.savez_compressed(os.path.join(OUT, "ref_models_48x1500.npz"), **rec)

# ---- divide and conquer, aligned (JC): 600 tips, backbone 60 (the reference requires every cluster < backbone size,
# src/divide_and_conquer/placement_close_k.cu:1334-1337)
n, L, B = 600, 1200, 60
codes, _ = synth.evolve(n, L, seed=21, regime="tiefree", gap_cols=0.03)
P = synth.pack4_np(codes)
_, cl0 = ORC.dc_as_shipped(ORC.msa_dist_matrix(P, L, 2), B, 0.0)
q = singleton_at_B(cl0, B)
P[[B, q]] = P[[q, B]]
inp, out = TMP + "/dc.bin", TMP + "/dc"
write_bin(inp, P, [L] * n, 4)
run("msa_dc", inp, out, 2, 15, B)
rec = dict(kind="ref_dc_msa", packed=P, seq_len=L, backbone=B, clusters=np.fromfile(out + ".clusters", np.int32),
           newick=open(out + ".nwk").read(), **read_arrays(out + ".arrays", n))
np.savez_compressed(os.path.join(OUT, "ref_dc_msa_600x1200.npz"), **rec)

# ---- divide and conquer, unaligned (Mash): 200 sequences of ~3 kb, backbone 40
n, L, B = 200, 3000, 40
codes, _ = synth.evolve(n, L, seed=22, regime="tiefree", gap_cols=0.01)
seqs = synth.unaligned(codes)
flat, offs, lens = synth.flatten2(seqs)
_, cl0 = ORC.dc(ORC.mash_dist_matrix(ORC.sketch_all(flat, offs, lens, 15, 1000), 15), B)
q = singleton_at_B(cl0, B)
seqs[B], seqs[q] = seqs[q], seqs[B]
flat, offs, lens = synth.flatten2(seqs)
inp, out = TMP + "/dcm.bin", TMP + "/dcm"
write_bin(inp, [synth.pack2_np(s) for s in seqs], lens, 2)
run("mash_dc", inp, out, 1, 15, B)
rec = dict(kind="ref_dc_mash", flat=flat, offsets=offs, lens=lens, k=15, s=1000, backbone=B,
           clusters=np.fromfile(out + ".clusters", np.int32), newick=open(out + ".nwk").read(), **read_arrays(out + ".arrays", n))
np.savez_compressed(os.path.join(OUT, "ref_dc_mash_200x3000.npz"), **rec)

# ---- add-tips (-m 1 --add): backbone = the reference's own placement tree of the first B tips
n, L, B = 220, 1200, 90
codes, _ = synth.evolve(n, L, seed=23, regime="tiefree", gap_cols=0.03)
P = synth.pack4_np(codes)
inp, out = TMP + "/bb.bin", TMP + "/bb"
write_bin(inp, P[:B], [L] * B, 4)
run("msa_place", inp, out, 2)
bb = open(out + ".nwk").read().strip()
from dipper_b200 import newick as NW
_, _, nm = NW.parse(bb)
leaf_names = [x for x in nm if x]                      # order of appearance = the reference's leaf numbering
order = [int(x[1:]) - 1 for x in leaf_names] + list(range(B, n))
Pp = np.ascontiguousarray(P[order])
inp, out = TMP + "/add.bin", TMP + "/add"
write_bin(inp, Pp, [L] * n, 4)
open(TMP + "/bb.nwk", "w").write(bb + "\n")
run("msa_add", inp, out, 2, 15, TMP + "/bb.nwk")
rec = dict(kind="ref_add_msa", packed=Pp, seq_len=L, backbone=B, backbone_newick=bb, newick=open(out + ".nwk").read(),
           **read_arrays(out + ".arrays", n))
np.savez_compressed(os.path.join(OUT, "ref_add_msa_220x1200.npz"), **rec)

# ---- BASELINE config 4b: add tips onto the reference's own fixture dataset/t2.backbone.nwk (copied to tests/golden/):
# 1000 backbone tips + 200 queries evolved on a tree that contains the backbone (branch lengths x 50, see synth)
bbt = open(os.path.join(ROOT, "tests", "golden", "t2.backbone.nwk")).readline().strip()
nq, L = 200, 1000
tree, bb_leaves, bb_names, qnodes = synth.tree_with_queries(bbt, nq, seed=24, scale=50.0)
n = len(bb_leaves) + nq
codes, info = synth.evolve(n, L, seed=25, tree=tree, gap_cols=0.02)
row_of = {int(v): i for i, v in enumerate(info["order"])}
order = [row_of[v] for v in bb_leaves] + [row_of[v] for v in qnodes]
Pp = np.ascontiguousarray(synth.pack4_np(codes[order]))
inp, out = TMP + "/t2.bin", TMP + "/t2"
write_bin(inp, Pp, [L] * n, 4)
open(TMP + "/t2.nwk", "w").write(bbt + "\n")
run("msa_add", inp, out, 2, 15, TMP + "/t2.nwk")
rec = dict(kind="ref_add_t2", packed=Pp, seq_len=L, backbone=len(bb_leaves), backbone_names=np.array(bb_names),
           newick=open(out + ".nwk").read(), **read_arrays(out + ".arrays", n))
np.savez_compressed(os.path.join(OUT, "ref_add_t2_1200x1000.npz"), **rec)
print("wrote", sorted(f for f in os.listdir(OUT) if f.startswith("ref_")))
