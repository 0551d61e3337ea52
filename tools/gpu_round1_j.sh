#!/bin/bash
# session J: two-block look-ahead pruning + table early exit; slice-gather probe; ncu launch list + full capture; L=15
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_j.log
tail -4 gpurun_out/pytest_gpu_j.log
timeout 600 tools/_build/slice_gather > gpurun_out/slice_gather_j.jsonl 2> gpurun_out/slice_gather_j.err
cat gpurun_out/slice_gather_j.jsonl
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_3100mb_j.json 2> gpurun_out/bench_3100mb_j.err
grep -E "variant|index|cpu_baseline" gpurun_out/bench_3100mb_j.err; cat gpurun_out/bench_3100mb_j.json
GSX_FTAB=15 timeout 1500 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_3100mb_j_L15.json 2> gpurun_out/bench_3100mb_j_L15.err
cut -c1-200 gpurun_out/bench_3100mb_j_L15.json; grep -o '"lookups_per_guide[^,]*,[^,]*' gpurun_out/bench_3100mb_j_L15.json
K='regex:search_|locate_score|order_matches|scan_u32|scatter_matches|expand_hits|specificity|threshold'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv --log-file gpurun_out/launches_3100mb_j.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_j.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:search_fast -s 1 -c 1 -o gpurun_out/prof_search_3100mb_j python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_j.log 2>&1
ls -la gpurun_out | tail -20
