#!/bin/bash
# session AJ: edited-guide path diagnostics (diff against the general kernel at 120 Mb and 3.1 Gb) + full GPU test suite
set -x
mkdir -p gpurun_out
timeout 600 python tools/diag_variants.py --genome-mb 120 --n 32 > gpurun_out/diag_aj_120.log 2>&1; tail -30 gpurun_out/diag_aj_120.log
timeout 900 python tools/diag_variants.py --genome-mb 3100 --n 32 > gpurun_out/diag_aj_3100.log 2>&1; tail -30 gpurun_out/diag_aj_3100.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_aj.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_aj.log
tail -25 gpurun_out/pytest_gpu_aj.log
