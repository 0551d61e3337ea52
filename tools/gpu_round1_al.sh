#!/bin/bash
# session AL: guide records uploaded under the search, match arena copied back under arrange/locate; full GPU suite,
# headline bench (3.1 Gb, 200k guides/step), launch list of the bulge bench
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_al.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_al.log
tail -6 gpurun_out/pytest_gpu_al.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_al.json 2> gpurun_out/bench_al.err
tail -3 gpurun_out/bench_al.err; cat gpurun_out/bench_al.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_al_cfg3.csv python bench.py --rna-bulges 1 --dna-bulges 1 --guides-per-step 2048 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_al_cfg3_ncu.log 2>&1
tail -2 gpurun_out/bench_al_cfg3_ncu.log | cut -c1-600
