"""Exact placement (-p 0): the reference's own CUDA kernels (oracle/_ref/dipper_ref msa_place_exact) vs this library on
the same input, same box.  Prints one JSON line and writes gpurun_out/r1_exact_vs_reference.json."""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from dipper_b200 import api, newick, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
codes, _ = synth.evolve(n, L, seed=5, gap_cols=0.03, gap_runs=False)
P = synth.pack4_np(codes)
inp, out = "/tmp/exact_in.bin", "/tmp/exact_out"
with open(inp, "wb") as f:
    np.array([n, 4], np.int64).tofile(f)
    np.full(n, L, np.uint64).tofile(f)
    for r in P:
        np.ascontiguousarray(r, np.uint64).tofile(f)
ref = os.path.join(ROOT, "oracle", "_ref", "dipper_ref")
p = subprocess.run([ref, "msa_place_exact", inp, out, "2"], capture_output=True, text=True, timeout=1500)
rj = json.loads(p.stdout.strip().splitlines()[-1])
ctx = api.Context(0)
prm = api.Param(distanceType=2, in_="m")
msa = api.MSADeviceArrays(ctx); msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
for rep in range(2):
    pl = api.PlacementDeviceArrays(ctx); pl.allocateDeviceArrays(n)
    t0 = time.time(); pl.findPlacementTree(prm, msaDeviceArrays=msa); wall = time.time() - t0
mine = pl.printTree(synth.names(n))
theirs = open(out + ".nwk").read()
res = {"config": "exact placement -p 0, %d aligned x %d, JC" % (n, L), "reference_cuda_tree_ms": rj["tree_ms"],
       "ours_ms": ctx.elapsed_ms(api.T_PLACE), "ours_wall_s": wall, "speedup": rj["tree_ms"] / ctx.elapsed_ms(api.T_PLACE),
       "newick_text_identical": mine == theirs, "rf": newick.rf_distance(mine, theirs)}
print(json.dumps(res))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "r1_exact_vs_reference.json"), "w"), indent=1)
