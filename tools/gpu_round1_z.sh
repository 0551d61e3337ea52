#!/bin/bash
# session Z: final numbers -- full parity suite (incl. 3.1 Gb oracle samples of configs 2-4), north-star bench with CPU arm,
# reference arm, ncu launch list + full captures on the same command
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_z.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_z.log
tail -5 gpurun_out/pytest_gpu_z.log
timeout 1800 python bench.py > gpurun_out/bench_z.json 2> gpurun_out/bench_z.err
tail -4 gpurun_out/bench_z.err; cat gpurun_out/bench_z.json
K='regex:sweep_|search_|locate_score|order_matches|scan_u32|scatter_matches|expand_hits|specificity|threshold'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv --log-file gpurun_out/launches_3100mb_z.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_z.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 1 -c 1 -o gpurun_out/prof_sweep_3100mb_z python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_z.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:search_fast -s 1 -c 1 -o gpurun_out/prof_fast_3100mb_z python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_z2.log 2>&1
timeout 1800 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_z.json 2> gpurun_out/bench_ref_z.err
cat gpurun_out/bench_ref_z.json
timeout 900 python bench.py --genome-mb 120 --n-chr 8 --seed 2 --guides-per-step 100000 --steps 3 --warmup 3 > gpurun_out/bench_z_120mb.json 2> gpurun_out/bench_z_120mb.err
cat gpurun_out/bench_z_120mb.json
