#!/bin/bash
# session AF: CTA-per-guide match ordering; configs[3] shape (m=3 + bulges) with enough guides to fill the grid
set -x
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_af.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_af.log
tail -4 gpurun_out/pytest_gpu_af.log
timeout 1500 python bench.py --rna-bulges 1 --dna-bulges 1 --mismatches 3 --guides-per-step 2048 --steps 1 --warmup 1 --cpu-sample 32 > gpurun_out/bench_af_cfg3.json 2> gpurun_out/bench_af_cfg3.err
tail -3 gpurun_out/bench_af_cfg3.err; cat gpurun_out/bench_af_cfg3.json
