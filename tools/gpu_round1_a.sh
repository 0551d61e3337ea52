#!/bin/bash
# first GPU session: parity tests, smoke, random-gather roofline probe, first bench lines, ncu launch list + full capture
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/host.txt; free -g >> gpurun_out/host.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 tools/_build/gather_bench > gpurun_out/gather.jsonl 2> gpurun_out/gather.err
timeout 600 python bench.py --genome-mb 10 --guides-per-step 5000 --steps 3 --warmup 3 --cpu-sample 1000 > gpurun_out/bench_10mb.json 2> gpurun_out/bench_10mb.err
timeout 900 python bench.py --genome-mb 120 --guides-per-step 20000 --steps 3 --warmup 3 --cpu-sample 2000 > gpurun_out/bench_120mb.json 2> gpurun_out/bench_120mb.err
for v in 0 2 3 4; do
  GSX_SEARCH_VARIANT=$v timeout 300 python bench.py --genome-mb 120 --guides-per-step 20000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_120mb_v$v.json 2> gpurun_out/bench_120mb_v$v.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --genome-mb 120 --guides-per-step 20000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 1 -c 1 -o gpurun_out/prof_search_r01 python bench.py --genome-mb 120 --guides-per-step 20000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1
ls -la gpurun_out
