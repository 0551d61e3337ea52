#!/bin/bash
# session AE: the other BASELINE configs as bench lines: configs[4] shape (m=4, NGG + NAG) and configs[3] shape (m=3 + bulges)
set -x
mkdir -p gpurun_out
timeout 1500 python bench.py --alt-pam NAG --mismatches 4 --guides-per-step 50000 --steps 2 --warmup 3 --cpu-sample 1000 > gpurun_out/bench_ae_cfg4.json 2> gpurun_out/bench_ae_cfg4.err
tail -3 gpurun_out/bench_ae_cfg4.err; cat gpurun_out/bench_ae_cfg4.json
timeout 1500 python bench.py --rna-bulges 1 --dna-bulges 1 --mismatches 3 --guides-per-step 256 --steps 2 --warmup 3 --cpu-sample 32 > gpurun_out/bench_ae_cfg3.json 2> gpurun_out/bench_ae_cfg3.err
tail -3 gpurun_out/bench_ae_cfg3.err; cat gpurun_out/bench_ae_cfg3.json
