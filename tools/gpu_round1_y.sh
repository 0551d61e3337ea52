#!/bin/bash
# session Y: single-load runs + xor table in shared memory + slim statistics
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "slice_major or sweep_kernel or fast_and_general or golden or kmer" > gpurun_out/pytest_gpu_y.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_y.log
tail -3 gpurun_out/pytest_gpu_y.log
timeout 1500 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --sweep-variants s5v2,s5v5,s4v2,s5v0 > gpurun_out/bench_3100mb_y.json 2> gpurun_out/bench_3100mb_y.err
grep -E "variant|index" gpurun_out/bench_3100mb_y.err
timeout 1500 python bench.py --guides-per-step 50000 --steps 3 --warmup 3 --no-cpu-baseline --sweep-variants s5v2,s5v5 > gpurun_out/bench_3100mb_y50k.json 2> gpurun_out/bench_3100mb_y50k.err
grep -E "variant|index" gpurun_out/bench_3100mb_y50k.err
