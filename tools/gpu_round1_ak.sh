#!/bin/bash
# session AK: sweep kernel parks drained before the lane-parallel step; full GPU suite, 3.1 Gb diff against the general kernel, cfg3 bench
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_ak.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_ak.log
tail -12 gpurun_out/pytest_gpu_ak.log
timeout 900 python tools/diag_variants.py --genome-mb 3100 --n 48 > gpurun_out/diag_ak_3100.log 2>&1; tail -30 gpurun_out/diag_ak_3100.log
timeout 900 python bench.py --rna-bulges 1 --dna-bulges 1 --mismatches 3 --guides-per-step 2048 --steps 2 --warmup 3 --cpu-sample 32 > gpurun_out/bench_ak_cfg3.json 2> gpurun_out/bench_ak_cfg3.err
tail -3 gpurun_out/bench_ak_cfg3.err; cat gpurun_out/bench_ak_cfg3.json
