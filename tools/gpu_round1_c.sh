#!/bin/bash
# third GPU session: north-star bench (3.1 Gb) with CPU arm, reference arm, ncu launch list + full capture at 3.1 Gb
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_c.log
tail -3 gpurun_out/pytest_gpu_c.log
timeout 1500 python bench.py > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
tail -6 gpurun_out/bench_c.err; cat gpurun_out/bench_c.json
timeout 1500 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_c.json 2> gpurun_out/bench_ref_c.err
cat gpurun_out/bench_ref_c.json
K='regex:search_kernel|locate_score|order_matches|scan_u32|scatter_matches|expand_hits|specificity|threshold'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 200 --csv --log-file gpurun_out/launches_3100mb.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_c.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 1 -c 1 -o gpurun_out/prof_search_3100mb python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_c.log 2>&1
ls -la gpurun_out
