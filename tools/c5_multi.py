"""BASELINE configs[4] (divide and conquer, -m 3, aligned n tips x L sites, backbone n / 20) on 1, 2, 4, 8 GPUs of one box
through the native multi-device entry (dipb_multi_dc, csrc/multi.cu: one process, one host thread per device).
usage: python tools/c5_multi.py [tips] [sites] [device counts, e.g. 1,2,4,8]   -> gpurun_out/r2_c5_multi_<tips>.json"""
import hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dipper_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
have = torch.cuda.device_count()
counts = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 4, 8]
counts = [c for c in counts if c <= have]
t0 = time.time()
P = synth.evolve_parallel_packed(n, L, seed=4)
t_gen = time.time() - t0
print("generated %d x %d in %.1f s on %d host cores" % (n, L, t_gen, os.cpu_count()), flush=True)
prm = api.Param(distanceType=2, in_="m")
names = synth.names(n)
rec = {"config": "C5: divide-and-conquer -m 3, %d aligned tips x %d sites, backbone %d, one process / %s devices" % (n, L, n // 20, counts),
       "host_gen_s": t_gen, "runs": []}
base = None
for nd in counts:
    md = api.MultiDevice(list(range(nd)))
    t0 = time.time(); md.allocateDeviceArrays(P, L); t_up = time.time() - t0
    t0 = time.time(); kp = md.findTreeDC(prm); wall = time.time() - t0
    nwk = kp.printTree(names)
    h = hashlib.sha256(nwk.encode()).hexdigest()[:16]
    r = {"devices": nd, "upload_s": t_up, "dc_wall_s": wall, "sharded_stage_ms": md.elapsed_ms(0), "rest_ms": md.elapsed_ms(2),
         "tips_per_s": n / wall, "newick_sha256_16": h}
    if base is None:
        base = r
    r["speedup_vs_1"] = base["dc_wall_s"] / wall
    r["efficiency"] = r["speedup_vs_1"] / nd * base["devices"]
    r["tree_identical_to_first"] = h == base["newick_sha256_16"]
    rec["runs"].append(r)
    print(json.dumps(r), flush=True)
    kp.deallocateDeviceArrays(); md.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rec, open(os.path.join("gpurun_out", "r2_c5_multi_%d.json" % n), "w"), indent=1)
