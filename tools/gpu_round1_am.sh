#!/bin/bash
# session AM: pipelined whole-file driver + parallel CSV ingest + table-driven formatter; match arena copy after the hit-count read-back
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_am.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_am.log
tail -6 gpurun_out/pytest_gpu_am.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_am.json 2> gpurun_out/bench_am.err
tail -3 gpurun_out/bench_am.err; cat gpurun_out/bench_am.json
