set -x
timeout 300 python -m pytest tests/test_msa_gpu.py -m gpu -x -q -k "operand_formats or tensor_core" 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msa_tc2_kernel -c 1 -o gpurun_out/r2_tc2_e2m1 python tools/ncu_target.py > gpurun_out/ncu_tc2.log 2>&1; tail -2 gpurun_out/ncu_tc2.log
python tools/ncu_summary.py gpurun_out/r2_tc2_e2m1.ncu-rep gpurun_out/r2_ncu_tc2_30k.json "msa_tc2_kernel<2> (e2m1 operands, kind::f8f6f4) at C3, ncu --set full --clock-control none" | cut -c1-700
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_30k_ours.json 2> gpurun_out/bench.err; tail -c 300 gpurun_out/bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_30k_ours.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['phases'], d['e2e'], d['dist_kernel'])"
