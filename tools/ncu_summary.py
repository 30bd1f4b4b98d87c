"""Summarise an ncu capture into a small JSON under profiles/ (what bench.py's `traffic` fields read).

usage (on a GPU box, after e.g.
   ncu --set full --clock-control none --import-source on -k regex:msa_tc2_kernel -c 1 -o gpurun_out/r2_tc2 python tools/ncu_target.py):
   python tools/ncu_summary.py gpurun_out/r2_tc2.ncu-rep gpurun_out/r2_ncu_tc2_30k.json "what was captured"
Reads `ncu -i <rep> --page raw --csv`, keeps one launch (the first) and the metrics listed below."""
import csv, io, json, subprocess, sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "launch__grid_size", "launch__cluster_size", "launch__registers_per_thread", "sm__cycles_elapsed.avg.per_second"]


def main():
    rep, out, what = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    launches = []
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        d = dict(zip(names, r))
        m = {}
        for k in KEEP:
            if k in d and d[k] != "":
                m[k] = (d[k] + " " + units[names.index(k)]).strip()
        launches.append({"kernel": d.get("Kernel Name", ""), "metrics": m})
    json.dump({"what": what, "launches": launches[:4], "metrics": launches[0]["metrics"] if launches else {}}, open(out, "w"), indent=1)
    print(json.dumps(launches[0] if launches else {}))


if __name__ == "__main__":
    main()
