#!/bin/bash
# session AG: host-expanded task seeds for small batches on the general kernel
set -x
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_ag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_ag.log
tail -4 gpurun_out/pytest_gpu_ag.log
timeout 1500 python bench.py --rna-bulges 1 --dna-bulges 1 --mismatches 3 --guides-per-step 256 --steps 2 --warmup 3 --cpu-sample 32 > gpurun_out/bench_ag_cfg3.json 2> gpurun_out/bench_ag_cfg3.err
tail -3 gpurun_out/bench_ag_cfg3.err; cat gpurun_out/bench_ag_cfg3.json
