"""Parity at the headline size (BASELINE config C3, 30 000 tips x 30 000 sites, JC + NJ), outside any timed region:
  1. the default NJ search (DIPB_NJ_CLUSTER: bound-pruned, one thread-block cluster) returns the SAME arrays as the
     exhaustive full scan (DIPB_NJ_FULLSCAN, the reference algorithm with all state on the device);
  2. the tree equals the reference's own CUDA objects' tree (oracle/_ref/dipper_ref msa_nj on the same input):
     RF distance 0, branch lengths within 1e-5 (north star).
Writes profiles/r2_c3_tree_verify.json (via gpurun_out/).   usage: python tools/verify_c3_tree.py [tips] [sites]"""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from dipper_b200 import api, newick
from bench import gen_data, write_ref_bin

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
P = gen_data(n, L, 1)
names = ["T%d" % (i + 1) for i in range(n)]
ctx = api.Context(0)
prm = api.Param(distanceType=2, in_="m")
msa = api.MSADeviceArrays(ctx)
msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
rec = {"what": "C3 tree parity at full size", "tips": n, "sites": L}
res = {}
for label, algo in (("cluster", api.NJ_CLUSTER), ("fullscan", api.NJ_FULLSCAN)):
    nj = api.NJDeviceArrays(ctx)
    nj.getDismatrix(n, prm, msaDeviceArrays=msa)
    t0 = time.time()
    nwk = nj.findNeighbourJoiningTree(names, algo)
    rec["nj_%s_ms" % label] = ctx.elapsed_ms(api.T_NJ)
    res[label] = (nj.result, nwk)
    nj.deallocateDeviceArrays()
a, b = res["cluster"][0], res["fullscan"][0]
rec["cluster_equals_fullscan_children"] = bool(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]))
rec["cluster_equals_fullscan_lengths_bitwise"] = bool(np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3]))
rec["cluster_equals_fullscan_newick_text"] = bool(res["cluster"][1] == res["fullscan"][1])
exe = os.path.join(ROOT, "oracle", "_ref", "dipper_ref")
if os.path.exists(exe) and not os.environ.get("SKIP_REF"):
    tmp = tempfile.mkdtemp(prefix="dipb_verify_")
    inp = os.path.join(tmp, "in.bin")
    write_ref_bin(inp, P, L)
    t0 = time.time()
    p = subprocess.run([exe, "msa_nj", inp, os.path.join(tmp, "o"), "2"], capture_output=True, text=True)
    rec["reference_wall_s"] = time.time() - t0
    rec["reference_rc"] = p.returncode
    if p.returncode == 0:
        rec["reference_phases"] = json.loads(p.stdout.strip().splitlines()[-1])
        ref = open(os.path.join(tmp, "o.nwk")).read()
        mine = res["cluster"][1]
        rec["rf_vs_reference"] = newick.rf_distance(mine, ref)
        rec["max_branch_diff_vs_reference"] = newick.max_branch_diff(mine, ref) if rec["rf_vs_reference"] == 0 else None
    else:
        rec["reference_stderr"] = p.stderr[-500:]
print(json.dumps(rec))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rec, open(os.path.join(ROOT, "gpurun_out", "r2_c3_tree_verify.json"), "w"), indent=1)
ok = rec["cluster_equals_fullscan_children"] and rec["cluster_equals_fullscan_lengths_bitwise"] and rec.get("rf_vs_reference", 0) == 0
sys.exit(0 if ok else 1)
