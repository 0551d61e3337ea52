#!/bin/bash
# session AA: sweep filter extended by the two levels behind the look-ahead window (sum2: the PAM characters for 20-mers)
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_aa.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_aa.log
tail -5 gpurun_out/pytest_gpu_aa.log
timeout 1800 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sweep-variants s5v2,s5v0,s4v2 > gpurun_out/bench_aa.json 2> gpurun_out/bench_aa.err
grep -E "variant|index" gpurun_out/bench_aa.err; cat gpurun_out/bench_aa.json
timeout 1500 python bench.py --guides-per-step 50000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_aa_50k.json 2> gpurun_out/bench_aa_50k.err
cat gpurun_out/bench_aa_50k.json
