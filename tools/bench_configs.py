"""Secondary measurements for BASELINE.json configs C1, C2, C4 (scaled), C5 (scaled) on one GPU.
Prints one JSON line per config; results are copied to profiles/.  Synthetic data, seeded."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dipper_b200 import api, synth

ctx = api.Context(0)
out = []


def aligned(n, L, seed=1):
    codes, _ = synth.evolve(n, L, seed=seed, gap_cols=0.03, gap_runs=False)
    return synth.pack4_np(codes)


def unaligned(n, L, seed=2):
    codes, _ = synth.evolve(n, L, seed=seed, gap_cols=0.01, gap_runs=False)
    seqs = synth.unaligned(codes)
    return [synth.pack2_np(s) for s in seqs], np.array([len(s) for s in seqs], np.uint64)


def rec(name, **kw):
    kw["config"] = name
    print(json.dumps(kw), flush=True)
    out.append(kw)


which = sys.argv[1:] or ["c1", "c2", "c4", "c5"]
if "c1" in which:       # C1 stand-in: 2000 x 10000 aligned, JC, NJ
    n, L = 2000, 10000
    P = aligned(n, L)
    prm = api.Param(distanceType=2, in_="m")
    for rep in range(2):
        t0 = time.time()
        msa = api.MSADeviceArrays(ctx); msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
        nj = api.NJDeviceArrays(ctx); nj.getDismatrix(n, prm, msaDeviceArrays=msa)
        nj.findNeighbourJoiningTree(synth.names(n)); wall = time.time() - t0
        nj.deallocateDeviceArrays()
    rec("C1 aligned 2000x10000 JC + NJ", dist_ms=ctx.elapsed_ms(api.T_MSA_DIST), nj_ms=ctx.elapsed_ms(api.T_NJ), e2e_s=wall)
if "c2" in which:       # C2 stand-in: 10000 unaligned ~10 kb, Mash + NJ
    n, L = 10000, 10000
    packed, lens = unaligned(n, L)
    prm = api.Param(kmerSize=15, sketchSize=1000, in_="r")
    for rep in range(2):
        t0 = time.time()
        m = api.MashDeviceArrays(ctx); m.allocateDeviceArrays(packed, lens, n, prm); m.sketchConstructionOnGpu()
        nj = api.NJDeviceArrays(ctx); nj.getDismatrix(n, prm, mashDeviceArrays=m)
        nj.findNeighbourJoiningTree(synth.names(n)); wall = time.time() - t0
        nj.deallocateDeviceArrays()
    kmers = float((lens - 14).sum())
    rec("C2 unaligned 10000 x ~10kb Mash + NJ", sketch_ms=ctx.elapsed_ms(api.T_SKETCH), kmers_per_s=kmers / ctx.elapsed_ms(api.T_SKETCH) * 1e3,
        mash_dist_ms=ctx.elapsed_ms(api.T_MASH_DIST), mash_pairs_per_s=n * (n - 1) / 2 / ctx.elapsed_ms(api.T_MASH_DIST) * 1e3,
        nj_ms=ctx.elapsed_ms(api.T_NJ), e2e_s=wall)
    if "c4" in which:   # C4 scaled: k-closest placement of the same 10000 (Mash)
        kp = api.KPlacementDeviceArrays(ctx); kp.allocateDeviceArrays(n)
        t0 = time.time(); kp.findPlacementTree(prm, mashDeviceArrays=m); wall = time.time() - t0
        rec("C4 scaled: placement -m 1, 10000 unaligned (Mash rows on the fly)", place_ms=ctx.elapsed_ms(api.T_PLACE), tips_per_s=n / wall)
if "c4" in which:       # placement from aligned input, 30000 tips x 10000 sites
    n, L = 30000, 10000
    P = aligned(n, L, seed=3)
    prm = api.Param(distanceType=2, in_="m")
    msa = api.MSADeviceArrays(ctx); msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
    kp = api.KPlacementDeviceArrays(ctx); kp.allocateDeviceArrays(n)
    t0 = time.time(); kp.findPlacementTree(prm, msaDeviceArrays=msa); wall = time.time() - t0
    rec("C4 scaled: placement -m 1, 30000 aligned x 10000", place_ms=ctx.elapsed_ms(api.T_PLACE), tips_per_s=n / wall)
    if "c5" in which:   # C5 scaled: divide and conquer on the same 30000 tips (backbone 1500)
        kp2 = api.KPlacementDeviceArrays(ctx); kp2.allocateDeviceArrays(n)
        t0 = time.time(); kp2.findTreeDC(prm, msaDeviceArrays=msa); wall = time.time() - t0
        rec("C5 scaled: divide-and-conquer -m 3, 30000 aligned x 10000, backbone 1500", dc_ms=ctx.elapsed_ms(api.T_PLACE), tips_per_s=n / wall,
            clusters=int(len(set(kp2.clusterID[kp2.clusterID >= 0]))))
if "c5big" in which:    # larger D&C: 200000 tips x 10000 sites, backbone 10000
    n, L = 200000, 10000
    P = aligned(n, L, seed=4)
    prm = api.Param(distanceType=2, in_="m")
    msa = api.MSADeviceArrays(ctx); msa.allocateDeviceArrays(P, np.full(n, L, np.uint64), n, prm)
    kp2 = api.KPlacementDeviceArrays(ctx); kp2.allocateDeviceArrays(n)
    t0 = time.time(); kp2.findTreeDC(prm, msaDeviceArrays=msa); wall = time.time() - t0
    rec("C5 scaled: divide-and-conquer -m 3, 200000 aligned x 10000, backbone 10000", dc_ms=ctx.elapsed_ms(api.T_PLACE), tips_per_s=n / wall,
        clusters=int(len(set(kp2.clusterID[kp2.clusterID >= 0]))))
json.dump(out, open(os.path.join("gpurun_out", "r1_configs.json"), "w"), indent=1)
