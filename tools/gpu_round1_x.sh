#!/bin/bash
# session X: sweep runs with two summary loads in flight per lane, xor table in shared memory
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "slice_major or sweep_kernel or fast_and_general or golden or kmer" > gpurun_out/pytest_gpu_x.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_x.log
tail -3 gpurun_out/pytest_gpu_x.log
timeout 1500 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --sweep-variants s5v2,s5v0,s5v1,s4v0,s4v2 > gpurun_out/bench_3100mb_x.json 2> gpurun_out/bench_3100mb_x.err
grep -E "variant|index" gpurun_out/bench_3100mb_x.err
cat gpurun_out/bench_3100mb_x.json
timeout 1500 python bench.py --guides-per-step 50000 --steps 3 --warmup 3 --no-cpu-baseline --sweep-variants s5v2,s5v0,s5v1 > gpurun_out/bench_3100mb_x50k.json 2> gpurun_out/bench_3100mb_x50k.err
grep -E "variant|index" gpurun_out/bench_3100mb_x50k.err
