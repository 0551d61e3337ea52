#!/bin/bash
# first GPU session of round 2: the opt-in paths written at the end of round 1 at full genome size
#  1. all alternative PAMs in one pass (GSX_FUSED_PAMS=1) on the BASELINE configs[4] shape, against the two-pass default
#  2. the specialised kernels on a genome with runs of N (GSX_FAST_ON_N=1), against the general kernel (parity_on_cpu_sample in both lines)
#  3. N = 2 headline line with the current kernels
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r2a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2a.log; tail -4 gpurun_out/pytest_gpu_r2a.log
for fused in 0 1; do
  GSX_FUSED_PAMS=$fused timeout 900 python bench.py --alt-pam NAG --mismatches 4 --guides-per-step 50000 --steps 2 --warmup 3 --cpu-sample 400 --no-file-e2e > gpurun_out/bench_r2a_cfg4_fused$fused.json 2> gpurun_out/bench_r2a_cfg4_fused$fused.err
  tail -2 gpurun_out/bench_r2a_cfg4_fused$fused.err; cut -c1-400 gpurun_out/bench_r2a_cfg4_fused$fused.json
done
for on in 0 1; do
  GSX_FAST_ON_N=$on timeout 900 python bench.py --n-runs 300 --guides-per-step 200000 --steps 2 --warmup 3 --cpu-sample 2000 --no-file-e2e > gpurun_out/bench_r2a_nruns_fast$on.json 2> gpurun_out/bench_r2a_nruns_fast$on.err
  tail -2 gpurun_out/bench_r2a_nruns_fast$on.err; cut -c1-400 gpurun_out/bench_r2a_nruns_fast$on.json
done
#  4. pending paths: forced-position pruning of the sweep (tests, then the bulge bench with and without it)
GSX_TEST_PENDING=1 timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "substituted_insert" > gpurun_out/pytest_gpu_r2a_pending.log 2>&1; tail -3 gpurun_out/pytest_gpu_r2a_pending.log
for forced in 0 1; do
  GSX_FORCED_SWEEP=$forced timeout 600 python bench.py --rna-bulges 1 --dna-bulges 1 --guides-per-step 2048 --steps 2 --warmup 3 --cpu-sample 16 --no-file-e2e > gpurun_out/bench_r2a_cfg3_forced$forced.json 2> gpurun_out/bench_r2a_cfg3_forced$forced.err
  cut -c1-300 gpurun_out/bench_r2a_cfg3_forced$forced.json
done
#  5. headline with two host threads taking turns on the device
timeout 600 python bench.py --steps 6 --warmup 3 --e2e-threads 2 --no-file-e2e --no-cpu-baseline > gpurun_out/bench_r2a_e2e_threads2.json 2> gpurun_out/bench_r2a_e2e_threads2.err; cut -c1-900 gpurun_out/bench_r2a_e2e_threads2.json
