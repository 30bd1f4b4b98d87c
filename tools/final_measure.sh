# Round-end measurement pass on one B200 (run through gpurun from the repo root): phase table, bench line, ncu launch list,
# ncu captures of the two dominant kernels, secondary configs.  Outputs land in gpurun_out/ and are copied to profiles/ by hand.
set -x
mkdir -p gpurun_out
(DIPB_NJ_PROFILE=1 NJ_REPS=1 timeout 200 python tools/nj_profile.py 30000 30000; DIPB_NJ_PROFILE=1 NJ_REPS=1 timeout 100 python tools/nj_profile.py 4000 30000) > gpurun_out/r2_nj_cluster_phases.txt 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_30k_ours.json 2> gpurun_out/bench.err; tail -c 300 gpurun_out/bench.err
DIPB_NJ_HELPERS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench30k.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
DIPB_NJ_HELPERS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:nj_cluster_kernel -c 1 -o gpurun_out/r2_nj_cluster python tools/ncu_target.py > gpurun_out/ncu_nj.log 2>&1; tail -3 gpurun_out/ncu_nj.log
python tools/ncu_summary.py gpurun_out/r2_nj_cluster.ncu-rep gpurun_out/r2_ncu_nj_cluster_30k.json "nj_cluster_kernel at C3 (30000 tips), kernel replay, helper clusters off (the cooperative launch does not start under ncu)" | cut -c1-600
timeout 300 python tools/bench_configs.py c1 c2 c4 2>&1 | tail -5 | cut -c1-400
