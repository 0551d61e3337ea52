#!/bin/bash
# session AO: formatter slices by rows -- whole-file leg of a bulge batch at 120 Mb (short run: the round's GPU budget is nearly spent)
set -x
mkdir -p gpurun_out
timeout 95 python bench.py --genome-mb 120 --n-chr 8 --seed 2 --rna-bulges 1 --dna-bulges 1 --guides-per-step 2048 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ao_120mb_cfg3.json 2> gpurun_out/bench_ao_120mb_cfg3.err
tail -3 gpurun_out/bench_ao_120mb_cfg3.err; cat gpurun_out/bench_ao_120mb_cfg3.json
