#!/bin/bash
# session AC: final numbers -- full parity suite (incl. 3.1 Gb oracle samples of configs 2-4), north-star bench with CPU arm,
# reference arm, ncu launch list + full captures on the same command
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_ac.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_ac.log
tail -5 gpurun_out/pytest_gpu_ac.log
timeout 1800 python bench.py > gpurun_out/bench_ac.json 2> gpurun_out/bench_ac.err
tail -4 gpurun_out/bench_ac.err; cat gpurun_out/bench_ac.json
K='regex:sweep_|search_|locate_score|order_matches|scan_u32|scatter_matches|expand_hits|specificity|threshold'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv --log-file gpurun_out/launches_3100mb_ac.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_ac.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 1 -c 1 -o gpurun_out/prof_sweep_3100mb_ac python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_ac.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:search_fast -s 1 -c 1 -o gpurun_out/prof_fast_3100mb_ac python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_ac2.log 2>&1
timeout 1800 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_ac.json 2> gpurun_out/bench_ref_ac.err
cat gpurun_out/bench_ref_ac.json
timeout 900 python bench.py --genome-mb 120 --n-chr 8 --seed 2 --guides-per-step 100000 --steps 3 --warmup 3 > gpurun_out/bench_ac_120mb.json 2> gpurun_out/bench_ac_120mb.err
cat gpurun_out/bench_ac_120mb.json
