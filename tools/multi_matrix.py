"""C3 distance matrix (30 000 x 30 000, JC) through the native multi-device entry on 1, 2, 4, 8 GPUs of one box:
row blocks balanced by triangle area, peer-read gather with fused mirror (csrc/multi.cu).  -> gpurun_out/r2_multi_matrix.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dipper_b200 import api
from bench import gen_data

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
P = gen_data(n, L, 1)
prm = api.Param(distanceType=2, in_="m")
have = torch.cuda.device_count()
rec = {"config": "C3 distance matrix, %d x %d JC, native multi-device entry (one process)" % (n, L), "runs": []}
ref = None
for nd in [d for d in (1, 2, 4, 8) if d <= have]:
    md = api.MultiDevice(list(range(nd)))
    md.allocateDeviceArrays(P, L)
    best = None
    for rep in range(4):
        M = md.distMatrix(prm)
        t = [md.elapsed_ms(k) for k in range(4)]
        if rep and (best is None or t[3] < best[3]):
            best = t
        if rep == 3:
            D = M.to_host() if n <= 12000 else None
        M.free()
    pairs = n * (n - 1) / 2
    r = {"devices": nd, "compute_ms": best[0], "gather_mirror_ms": best[1], "total_ms": best[3], "pairs_per_s": pairs / best[3] * 1e3}
    if ref is None:
        ref = r
    r["speedup_vs_1"] = ref["total_ms"] / r["total_ms"]
    r["efficiency"] = r["speedup_vs_1"] / nd
    rec["runs"].append(r)
    print(json.dumps(r), flush=True)
    md.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rec, open("gpurun_out/r2_multi_matrix.json", "w"), indent=1)
