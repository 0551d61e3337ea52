#!/bin/bash
# fourth GPU session: DRAM over-fetch probe (load policy x L2 fetch granularity), fast kernel parity + sweeps
set -x
mkdir -p gpurun_out
for g in 0 32 64 128; do if [ $g = 0 ]; then tools/_build/gather_policy; else tools/_build/gather_policy $g; fi; done > gpurun_out/gather_policy.jsonl 2> gpurun_out/gather_policy.err
cat gpurun_out/gather_policy.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum --clock-control none -k regex:^k -c 9 --csv --log-file gpurun_out/gather_policy_ncu.csv tools/_build/gather_policy > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:^k -c 9 --csv --log-file gpurun_out/gather_policy_ncu32.csv tools/_build/gather_policy 32 > /dev/null 2>&1
timeout 600 tools/_build/gather_bench > gpurun_out/gather2.jsonl 2> gpurun_out/gather2.err
grep grouped gpurun_out/gather2.jsonl; grep hint gpurun_out/gather2.jsonl
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_d.log
tail -15 gpurun_out/pytest_gpu_d.log
timeout 900 python bench.py --genome-mb 120 --n-chr 8 --seed 2 --guides-per-step 20000 --steps 3 --warmup 3 --no-cpu-baseline --sweep-variants g0,g1,f0,f1,f2,f3,f4 > gpurun_out/bench_120mb_d.json 2> gpurun_out/bench_120mb_d.err
tail -9 gpurun_out/bench_120mb_d.err
timeout 1500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sweep-variants g0,f0,f1,f2,f3,f4 > gpurun_out/bench_3100mb_d.json 2> gpurun_out/bench_3100mb_d.err
tail -9 gpurun_out/bench_3100mb_d.err
cat gpurun_out/bench_3100mb_d.json
