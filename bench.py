#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline: pairwise distances/s and NJ wall time at N tips.

Workload (config C3, BASELINE.json configs[2]): synthetic aligned MSA, 30 000 tips x
30 000 sites, JC distance matrix (-d 2) then conventional NJ (-m 2).  One "step" = one
full pass: packed sequences -> operand expansion (e2m1 simplex / validity codes) -> fp64 distance matrix -> NJ tree.

  value   whole-job pairs/s with the packed sequences already resident in HBM: K steps between two
          barriers + synchronize, max over ranks ("phases" are CUDA-event times inside the library).  The
          tensor-core operand expansion is paid in EVERY step (dipb_msa_drop_operands before it).
  e2e     same metric through the public C ABI from HOST (pinned) buffers: H2D of the 4-bit
          sequences + repack + expansion + distances + NJ + D2H of the tree, wall-clocked
  roofline  the NJ kernel (dominant: ~90 % of the step) against the measured HBM copy bandwidth on the
          bytes the kernel itself reads and writes (it is latency / issue bound: frac ~ 0.03); the
          reference's full-scan cost sum_n (n^2+4n)*8 B (SURVEY.md 8d yard-stick) is reported separately as
          "speedup_vs_fullscan_bytes", not as a roofline
  dist_kernel  the tcgen05 distance kernel (8-bit-container MMA rate) against the tensor roofline (int8 / fp8 dense
          peak taken as 2x the measured bf16 peak); traffic from the committed ncu capture
  cpu_baseline  the OpenMP oracle port on a bounded sample (rank 0, N=1 only)

`--impl reference` runs the reference's own CUDA objects (oracle/_ref/dipper_ref; the reference has no CPU
path) on the SAME workload (30 000 tips): one job takes ~55 s there, so the number of timed
jobs is bounded by a time budget (stated in the line) instead of the driver's step count.
Multi-GPU (torchrun): the distance matrix is row-block sharded over ranks and gathered
onto rank 0 (NCCL send/recv of the row blocks + a mirror kernel), NJ runs on rank 0 (BASELINE.json configs[2]).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pairwise distances/sec (JC distance matrix + NJ tree, whole job)"
UNIT = "pairs/s"


def workload_config(n, L):
    """Identical in both arms (the driver compares them)."""
    return {"workload": "C3: aligned MSA %d tips x %d sites, JC distance matrix + neighbor joining, one tree per step" % (n, L),
            "inputs": "larger than L2 (%.0f MB packed sequences, %.1f GB fp64 matrix)" % (n * ((L + 15) // 16) * 8 / 1e6, n * n * 8 / 1e9),
            "timing": "every step builds everything from the packed sequences (no cached operands)"}


def measured_peaks():
    """(HBM GB/s, bf16 dense TFLOP/s burst, source) -- MEASURED_PEAKS.json, else the profiling guide's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), float(j["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def pinned_like(a):
    """Copy of a numpy array in page-locked host memory (the e2e contract: H2D from pinned memory)."""
    import torch
    t = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=True)     # returned too: it owns the memory
    out = t.numpy().view(a.dtype).reshape(a.shape)
    out[...] = a
    return out, t


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def gen_data(n, L, seed):
    from dipper_b200 import synth
    cache = os.path.join(tempfile.gettempdir(), "dipb_bench_%d_%d_%d.npy" % (n, L, seed))
    if os.path.exists(cache):
        return np.load(cache)
    codes, _ = synth.evolve(n, L, seed=seed, regime="tiefree", gap_cols=0.03, gap_runs=False)
    P = synth.pack4_np(codes)
    try:
        np.save(cache, P)
    except Exception:
        pass
    return P


def nj_algorithmic_bytes(n):
    # full-scan definition, (m^2 + 4m)*8 B per iteration with m active rows
    m = np.arange(3, n + 1, dtype=np.float64)
    return float(((m * m + 4 * m) * 8).sum())


def dist_algorithmic_intops(n, L):
    # SURVEY.md §8(d), bit-plane formulation: pairs x ceil(L/32) x (4 LOP3 + 2 POPC + 2 IADD)
    return n * (n - 1) / 2 * ((L + 31) // 32) * 8.0


def dist_tensor_ops(n, L):
    # tensor formulation (msa_tc.cu): per pair one {-1,0,1} dot product of length 3L (simplex codes) and one of
    # length L (validity), 2 ops per multiply-add
    return n * (n - 1) / 2 * (4.0 * L) * 2.0


def write_ref_bin(path, P, L):
    with open(path, "wb") as f:
        np.array([P.shape[0], 4], np.int64).tofile(f)
        np.full(P.shape[0], L, np.uint64).tofile(f)
        P.tofile(f)


def run_reference(args):
    """The reference's own CUDA objects on the same workload (rank 0 only).  A job takes ~55 s at 30 000 tips, so the
    timed jobs are bounded by --ref-budget seconds (at least one), whatever --steps says; one small warm-up job absorbs
    CUDA initialisation."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    exe = os.path.join(ROOT, "oracle", "_ref", "dipper_ref")
    n, L = args.tips, args.sites
    base = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "config": workload_config(n, L)}
    if not os.path.exists(exe):
        print(json.dumps(dict(base, unavailable="oracle/_ref/dipper_ref not built (reference sources absent at build time)")))
        return
    tmp = tempfile.mkdtemp(prefix="dipb_ref_")

    def job(P, tag):
        inp = os.path.join(tmp, tag + ".bin")
        write_ref_bin(inp, P, L)
        t0 = time.time()
        p = subprocess.run([exe, "msa_nj", inp, os.path.join(tmp, tag), "2"], capture_output=True, text=True)
        os.remove(inp)
        if p.returncode != 0:
            raise RuntimeError("dipper_ref failed: " + p.stderr[-200:].replace("\n", " "))
        j = json.loads(p.stdout.strip().splitlines()[-1])
        return j["dist_ms"], j["tree_ms"], j["alloc_ms"], time.time() - t0

    try:
        job(gen_data(min(n, 3000), L, args.seed + 1), "warm")
        P = gen_data(n, L, args.seed)
        times, t_begin = [], time.time()
        while len(times) < args.steps:
            times.append(job(P, "job"))
            per_job = (time.time() - t_begin) / len(times)
            if time.time() - t_begin + per_job > args.ref_budget:
                break
    except RuntimeError as e:
        print(json.dumps(dict(base, unavailable=str(e))))
        return
    t = np.array(times)
    dist_ms, nj_ms = float(t[:, 0].mean()), float(t[:, 1].mean())
    pairs = n * (n - 1) / 2
    val = pairs / ((dist_ms + nj_ms) / 1e3)
    sample = ("the full workload, %d tips x %d sites; %d timed job(s) of the %d requested steps fit the %d s budget "
              "(+ one 3000-tip warm-up job); reference CUDA objects on 1 B200, device phases dist %.1f s + NJ %.1f s"
              % (n, L, len(times), args.steps, args.ref_budget, dist_ms / 1e3, nj_ms / 1e3))
    print(json.dumps(dict(base, value=val, ms_per_step=dist_ms + nj_ms, steps_timed=len(times), scaling="strong", vs_baseline=None,
                          dtype="int32 counts + f64", data="synthetic",
                          phases={"dist_ms": dist_ms, "nj_ms": nj_ms, "alloc_ms": float(t[:, 2].mean()), "wall_s_per_job": float(t[:, 3].mean())},
                          cpu_baseline={"value": val, "unit": UNIT, "cores": 0, "kind": "reference", "sample": sample},
                          e2e={"value": pairs / float(t[:, 3].mean()), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                               "note": "wall clock of the whole dipper_ref process incl. reading its input file"})))


def cpu_baseline(args, L):
    from oracle import oracle as O       # checker / baseline leg only
    n = args.cpu_tips
    P = gen_data(n, L, args.seed)
    t0 = time.time()
    D = O.msa_dist_matrix(P, L, 2)
    t1 = time.time()
    O.nj(D)
    t2 = time.time()
    pairs = n * (n - 1) / 2
    return {"value": pairs / (t2 - t0), "unit": UNIT, "cores": O.num_threads(), "kind": "port",
            "sample": "%d tips x %d sites, OpenMP oracle: dist %.2f s + NJ %.2f s" % (n, L, t1 - t0, t2 - t1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tips", type=int, default=30000)
    ap.add_argument("--sites", type=int, default=30000)
    ap.add_argument("--ref-budget", type=int, default=170, help="seconds of timed reference jobs (--impl reference)")
    ap.add_argument("--cpu-tips", type=int, default=6000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--nj-algo", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from dipper_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n, L = args.tips, args.sites
    P, _pin = pinned_like(gen_data(n, L, args.seed))
    lens = np.full(n, L, np.uint64)
    ctx = api.Context(local)
    prm = api.Param(distanceType=2, in_="m")
    pairs = n * (n - 1) / 2

    from dipper_b200 import sharding

    def shard(r):
        return sharding.row_block_shards(n, world)[r]

    class DevView:   # __cuda_array_interface__ over the library's matrix for NCCL
        def __init__(self, ptr, count):
            self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from dipper_b200._lib import names_array
    tip_names = names_array(["T%d" % (i + 1) for i in range(n)])   # char** built once, as a C++ caller holds its names

    def one_step(msa, timed):
        """resident-input step: distances (sharded) -> reduce -> NJ on rank 0. Returns (dist_ms, comm_ms, nj_ms)."""
        r0, r1 = shard(rank)
        msa.dropOperands()                # the operand expansion is part of every step
        if world == 1:
            M = msa.distMatrix(prm)
        else:
            M = msa.distMatrix(prm, r0, r1)
        d_ms = ctx.elapsed_ms(api.T_MSA_DIST)
        c_ms = 0.0
        if world > 1:
            # gather: every other rank sends its row block (rows r0..r1 are contiguous in the matrix) to rank 0 over
            # NCCL point-to-point; rank 0 mirrors the received rows.  (A sum-reduce of the whole 7.2 GB matrix took
            # 292 ms at 2 GPUs; the row blocks of the other ranks are 2 GB.)
            from dipper_b200._lib import lib, check
            ptr = lib().dipb_matrix_device_ptr(M.h)
            t = torch.as_tensor(DevView(ptr, n * n), device="cuda")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if rank == 0:
                for src in range(1, world):
                    s0, s1 = shard(src)
                    if s1 > s0:
                        dist.recv(t[s0 * n:s1 * n], src=src)
            elif r1 > r0:
                dist.send(t[r0 * n:r1 * n], dst=0)
            e1.record()
            torch.cuda.synchronize()
            c_ms = e0.elapsed_time(e1)
            if rank == 0:
                for src in range(1, world):
                    s0, s1 = shard(src)
                    check(lib().dipb_matrix_mirror_rows(M.h, s0, s1))
                    ctx.sync()
                    c_ms += ctx.elapsed_ms(api.T_MSA_DIST)
        nj_ms = 0.0
        if rank == 0:
            nj = api.NJDeviceArrays(ctx)
            nj.matrix, nj.d_numSequences = M, n
            nj.findNeighbourJoiningTree(tip_names, args.nj_algo)
            nj_ms = ctx.elapsed_ms(api.T_NJ)
            nj.deallocateDeviceArrays()
        else:
            M.free()
        return d_ms, c_ms, nj_ms

    msa = api.MSADeviceArrays(ctx)
    msa.allocateDeviceArrays(P, lens, n, prm)
    for _ in range(args.warmup):
        one_step(msa, False)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.kernel_launches()
    t_wall0 = time.time()
    phases = []
    for _ in range(args.steps):
        phases.append(one_step(msa, True))
    barrier()
    t_wall = time.time() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.kernel_launches() - launches0
    ph = np.array(phases)
    # A step = everything between the two barriers divided by K: device phases (CUDA events inside the library, reported
    # under "phases") plus the host work between them (tree replay, Newick text).
    step_ms = t_wall * 1e3 / args.steps
    if world > 1:
        # Ranks overlap: a rank without NJ work runs one step ahead and then waits in its send, so per-rank phase times
        # do not add up.  The step time is the max over ranks; the phases are those of rank 0, which owns the
        # critical path (its distances, the gather, NJ).
        tt = torch.tensor([ph[:, 0].mean(), ph[:, 1].mean(), ph[:, 2].mean(), float(launches), t_wall * 1e3 / args.steps],
                          device="cuda", dtype=torch.float64)
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        r0v = tt.clone()
        dist.broadcast(r0v, src=0)
        d_ms, c_ms, nj_ms = float(mx[0]), float(r0v[1]), float(r0v[2])
        launches = int(sm[3])
        step_ms = float(mx[4])
    else:
        d_ms, c_ms, nj_ms = ph[:, 0].mean(), ph[:, 1].mean(), ph[:, 2].mean()
    nj_stats = ctx.nj_stats() if rank == 0 else {}

    # ---- e2e from host buffers through the C ABI (rank 0 drives; shards upload their own copy) ----
    e2e_t = []
    msa.deallocateDeviceArrays()
    for it in range(1 + max(1, args.steps)):
        barrier()
        t0 = time.time()
        m2 = api.MSADeviceArrays(ctx)
        m2.allocateDeviceArrays(P, lens, n, prm)              # H2D + repack inside
        one_step(m2, True)                                    # distances + NJ + D2H of the tree
        barrier()
        if it > 0:
            e2e_t.append(time.time() - t0)
        m2.deallocateDeviceArrays()
    e2e_s = float(np.mean(e2e_t))
    if world > 1:
        tt = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt[0])

    if rank == 0:
        peak, peak_bf16, peak_src = measured_peaks()
        m = np.arange(3, n + 1, dtype=np.float64)
        # Bytes the NJ kernel itself moves, per merge with m active rows: phase A reads rows x, y, last
        # (3m x 8), writes the two scratch rows (2m x 8); the helper clusters read them (2m x 8), write rows and columns
        # x, y of D (4m x 8) and fold the two new columns into the unit keys (2m x 4 B read-modify-write); plus the scan
        # units the search really loads (counted by the kernel).
        nj_kernel_bytes = float(nj_stats.get("bytes_scanned", 0)) + float((104.0 * m).sum())
        nj_ach = nj_kernel_bytes / (nj_ms / 1e3) / 1e9
        fullscan_bytes = nj_algorithmic_bytes(n)
        tc_ops = dist_tensor_ops(n, L)
        tc_ach = tc_ops / (d_ms / 1e3) / 1e12 / max(world, 1)

        def ncu_traffic(fn):
            """dram read + write bytes of one launch from a committed ncu summary (profiles/), else None."""
            try:
                mt = json.load(open(os.path.join(ROOT, "profiles", fn)))["metrics"]
                tot = 0.0
                for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    v, u = mt[k].split()[:2]
                    tot += float(v) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
                return tot
            except Exception:
                return None

        out = {
            "metric": METRIC, "value": pairs / (step_ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "e2m1 {-1,0,1} dot products -> exact integer counts in f32 accumulators (< 2^24; bit-identical to the int8/int32 and popcount paths) -> f64", "data": "synthetic",
            "config": workload_config(n, L),
            "sharding": "distance row blocks over %d GPU(s), gathered on rank 0; NJ on rank 0 (BASELINE configs[2])" % world,
            "phases": {"dist_ms": d_ms, "reduce_ms": c_ms, "nj_ms": nj_ms,
                       "dist_pairs_per_sec": pairs / (d_ms / 1e3), "nj_wall_s": nj_ms / 1e3,
                       "nj_us_per_merge": nj_ms * 1e3 / max(n - 2, 1),
                       "note": "dist_ms includes the operand expansion of every step"},
            "roofline": {"bound": "hbm", "achieved": nj_ach, "peak": peak, "unit": "GB/s", "frac": nj_ach / peak,
                         "traffic": ncu_traffic("r2_ncu_nj_cluster_30k.json"), "peak_source": peak_src,
                         "kernel": "nj_cluster_kernel (one launch = all %d merges; %.0f %% of the step)" % (n - 2, 100.0 * nj_ms / step_ms),
                         "algorithmic_bytes": nj_kernel_bytes,
                         "definition": "bytes the kernel itself reads and writes: per merge 104 B per active row (rows x, y, last in; scratch rows, "
                                       "rows and columns x, y of D, unit-key folds out) + the scan units actually loaded (kernel counter)",
                         "rows_selected": nj_stats.get("rows_scanned", 0), "scan_bytes": float(nj_stats.get("bytes_scanned", 0)),
                         "speedup_vs_fullscan_bytes": fullscan_bytes / nj_kernel_bytes,
                         "fullscan_bytes": fullscan_bytes,
                         "note": "not bandwidth bound: a chain of 3 cluster barriers per merge whose phases are instruction-issue and "
                                 "latency bound (profiles/r2_nj_cluster_phases.txt); fullscan_bytes = SURVEY.md 8(d) yard-stick "
                                 "(what the reference's search reads), kept apart from the roofline"},
            "dist_kernel": {"kernel": "msa_tc2_kernel<2> (tcgen05.mma.cta_group::2.kind::f8f6f4 on e2m1 operands unpacked by the TMA, 256x256 tiles per CTA pair; kind::i8 behind DIPB_TC_FMT=0) + msa_tc_expand_kernel", "bound": "tensor",
                            "achieved": tc_ach, "peak": 2.0 * peak_bf16, "unit": "TOP/s", "frac": tc_ach / (2.0 * peak_bf16),
                            "peak_definition": "8-bit dense (int8 = fp8 = f8f6f4 rate) = 2 x measured bf16 dense burst (%s); no measured 8-bit peak exists" % peak_src,
                            "algorithmic_ops": tc_ops, "traffic": ncu_traffic("r2_ncu_tc2_30k.json") or ncu_traffic("r1_ncu_tc2_30k.json"),
                            "traffic_source": "profiles/r2_ncu_tc2_30k.json if present, else r1_ncu_tc2_30k.json (dram read + write of one msa_tc2_kernel launch)",
                            "algorithmic_bytes": float(n) * ((L + 15) // 16) * 8 + float(n) * n * 8,
                            "bitplane_equivalent_Tops": dist_algorithmic_intops(n, L) / (d_ms / 1e3) / 1e12 / max(world, 1)},
            "e2e": {"value": pairs / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(P.nbytes + lens.nbytes),
                    "d2h_bytes_per_step": int((n - 1) * 24), "seconds": e2e_s, "host_buffers": "pinned"},
            "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": t_wall,
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args, L)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
