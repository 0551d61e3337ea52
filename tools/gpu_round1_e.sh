#!/bin/bash
set -x
mkdir -p gpurun_out
tools/_build/gather_width > gpurun_out/gather_width.jsonl 2>&1; cat gpurun_out/gather_width.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum --clock-control none -k regex:^kw -c 14 --csv --log-file gpurun_out/gather_width_ncu.csv tools/_build/gather_width > /dev/null 2>&1
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/pytest_gpu_e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_e.log
tail -5 gpurun_out/pytest_gpu_e.log
timeout 900 python bench.py --genome-mb 120 --n-chr 8 --seed 2 --guides-per-step 20000 --steps 3 --warmup 3 --no-cpu-baseline --sweep-variants g0,f0,f1,f2,f3,f4 > gpurun_out/bench_120mb_e.json 2> gpurun_out/bench_120mb_e.err
grep variant gpurun_out/bench_120mb_e.err
timeout 1500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sweep-variants g0,f0,f1,f2,f3,f4 > gpurun_out/bench_3100mb_e.json 2> gpurun_out/bench_3100mb_e.err
grep variant gpurun_out/bench_3100mb_e.err
cat gpurun_out/bench_3100mb_e.json
