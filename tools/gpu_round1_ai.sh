#!/bin/bash
# session AI: bulges through edited guides on the specialised kernels + radix-sort match ordering
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_ai.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_ai.log
tail -15 gpurun_out/pytest_gpu_ai.log
timeout 900 python bench.py --rna-bulges 1 --dna-bulges 1 --mismatches 3 --guides-per-step 2048 --steps 2 --warmup 3 --cpu-sample 32 > gpurun_out/bench_ai_cfg3.json 2> gpurun_out/bench_ai_cfg3.err
tail -5 gpurun_out/bench_ai_cfg3.err; cat gpurun_out/bench_ai_cfg3.json
