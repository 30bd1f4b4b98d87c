"""Divide-and-conquer (-m 3, BASELINE config C5 scaled) with stage 2 / stage 3 sharded over ranks.
torchrun --nproc-per-node N tools/dc_multi_gpu.py [tips] [sites]   (one process per GPU)
Prints one JSON line on rank 0: wall time, tips/s, and a check against the single-GPU tree."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from dipper_b200 import api, synth

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
check = "--check" in sys.argv
stage2_only = "--stage2-only" in sys.argv   # shard the query assignment only; rank 0 places all clusters
reps = 1 if "--once" in sys.argv else 2
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gloo = dist.new_group(backend="gloo")

cache = "/tmp/dipb_dc_%d_%d.npy" % (n, L)
if rank == 0 and not os.path.exists(cache):
    codes, _ = synth.evolve(n, L, seed=4, gap_cols=0.03, gap_runs=False)
    np.save(cache, synth.pack4_np(codes))
if world > 1:
    dist.barrier()
P = np.load(cache)
ctx = api.Context(local)
prm = api.Param(distanceType=2, in_="m")
msa = api.MSADeviceArrays(ctx); msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)


def all_gather(arr):
    if world == 1:
        return [arr]
    out = [None] * world
    dist.all_gather_object(out, arr, group=gloo)
    return out


def gather_to_root(blob):
    if world == 1:
        return [blob]
    out = [None] * world if rank == 0 else None
    dist.gather_object(blob, out, dst=0, group=gloo)
    return out


times = []
for rep in range(reps):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    kp = api.KPlacementDeviceArrays(ctx); kp.allocateDeviceArrays(n)
    kp.findTreeDC_sharded(prm, rank, world, all_gather, gather_to_root, msaDeviceArrays=msa, shard_clusters=not stage2_only)
    if world > 1:
        dist.barrier()
    times.append(time.time() - t0)
if rank == 0:
    res = {"config": "C5 scaled: D&C %d tips x %d sites, backbone %d" % (n, L, n // 20), "n_gpus": world, "sharded": "stage 2 only" if stage2_only else "stages 2 and 3",
           "wall_s": min(times), "tips_per_s": n / min(times)}
    if check:
        single = api.KPlacementDeviceArrays(ctx); single.allocateDeviceArrays(n)
        t0 = time.time()
        single.findTreeDC(prm, msaDeviceArrays=msa)
        res["single_gpu_wall_s"] = time.time() - t0
        res["speedup_vs_single_gpu"] = res["single_gpu_wall_s"] / res["wall_s"]
        a, b = kp.export(), single.export()
        res["identical_to_single_gpu"] = bool(all(np.array_equal(a[k], b[k]) for k in ("head", "e", "nxt", "belong")) and
                                              np.array_equal(a["len"][: 4 * n - 4], b["len"][: 4 * n - 4]))
    print(json.dumps(res))
    json.dump(res, open(os.path.join("gpurun_out", "r1_dc_%d_%dgpu.json" % (n, world)), "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
