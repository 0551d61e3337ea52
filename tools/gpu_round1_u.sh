#!/bin/bash
# session U (N GPUs of one box): scaling run of the bench exactly as the driver launches it
set -x
N=${1:-2}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_u_n$N.json 2> gpurun_out/bench_u_n$N.err
tail -5 gpurun_out/bench_u_n$N.err; cat gpurun_out/bench_u_n$N.json
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/bench_u_ref_n$N.json 2> gpurun_out/bench_u_ref_n$N.err
cat gpurun_out/bench_u_ref_n$N.json | cut -c1-400
