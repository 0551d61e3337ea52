#!/bin/bash
# session T: final numbers of the slice-major pipeline -- parity suite, north-star bench with CPU arm, reference arm,
# ncu launch list + full captures (sweep_kernel, search_fast_kernel) on the same command
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_t.log
tail -3 gpurun_out/pytest_gpu_t.log
timeout 1800 python bench.py > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err
tail -4 gpurun_out/bench_t.err; cat gpurun_out/bench_t.json
timeout 1500 python bench.py --guides-per-step 50000 --no-cpu-baseline > gpurun_out/bench_t_50k.json 2> gpurun_out/bench_t_50k.err
cat gpurun_out/bench_t_50k.json
K='regex:sweep_|search_|locate_score|order_matches|scan_u32|scatter_matches|expand_hits|specificity|threshold'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv --log-file gpurun_out/launches_3100mb_t.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_t.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 1 -c 1 -o gpurun_out/prof_sweep_3100mb_t python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_t.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:search_fast -s 1 -c 1 -o gpurun_out/prof_fast_3100mb_t python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_t2.log 2>&1
timeout 1800 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_t.json 2> gpurun_out/bench_ref_t.err
cat gpurun_out/bench_ref_t.json
ls -la gpurun_out | tail -12
