#!/bin/bash
# session AN: bulge bench (BASELINE configs[3] shape) with the final code: match arena copied back under locate/score, file_e2e leg
set -x
mkdir -p gpurun_out
timeout 220 python bench.py --rna-bulges 1 --dna-bulges 1 --mismatches 3 --guides-per-step 2048 --steps 2 --warmup 3 --cpu-sample 16 > gpurun_out/bench_an_cfg3.json 2> gpurun_out/bench_an_cfg3.err
tail -3 gpurun_out/bench_an_cfg3.err; cat gpurun_out/bench_an_cfg3.json
