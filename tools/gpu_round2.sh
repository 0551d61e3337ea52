#!/bin/bash
# GPU sessions of round 2 (one parameterised script; the per-session scripts of round 1 were folded into this one).
#   tools/gpu_round2.sh <tag> [pytest] [smoke] [plan <plan.json>] [bench <args...>]
# pytest: the -m gpu suite (with the formerly pending paths);  smoke: the plan at 120 Mb with small batches first, so that a
# Python error costs seconds and not an index build;  plan: tools/gpu_session.py on the 3.1 Gb genome;  bench: bench.py with the
# remaining arguments.  Everything lands in gpurun_out/<tag>*.
set -x
tag=$1; shift
mkdir -p gpurun_out
while [ $# -gt 0 ]; do
  case $1 in
    pytest)
      GSX_TEST_PENDING=1 timeout 1500 python -m pytest tests -m gpu -q --maxfail 8 --timeout 900 > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_gpu.log
      tail -5 gpurun_out/${tag}_pytest_gpu.log; shift;;
    smoke)
      plan=$2
      python - "$plan" > /tmp/smoke_plan.json <<'PY'
import json, sys
p = json.load(open(sys.argv[1]))
for g in p["genomes"]:
    for e in g["experiments"]:
        if "guides" in e: e["guides"] = 256 if (e.get("rna_bulges") or e.get("dna_bulges")) else 20000
        e["steps"], e["warmup"] = min(e.get("steps", 2), 2), 1
        if "parity_sample" in e: e["parity_sample"] = 4 if (e.get("rna_bulges") or e.get("dna_bulges")) else 64
        if "file_e2e" in e: e["file_e2e"]["guides"] = 50000
print(json.dumps(p))
PY
      timeout 600 python tools/gpu_session.py --plan /tmp/smoke_plan.json --tag ${tag}_smoke --genome-mb 120 2> gpurun_out/${tag}_smoke.err || { tail -20 gpurun_out/${tag}_smoke.err; echo "SMOKE FAILED"; exit 1; }
      grep -c '"error"' gpurun_out/${tag}_smoke.jsonl; grep '"error"' gpurun_out/${tag}_smoke.jsonl | cut -c1-600
      shift 2;;
    plan)
      timeout 2400 python tools/gpu_session.py --plan $2 --tag ${tag} 2> gpurun_out/${tag}.err; tail -3 gpurun_out/${tag}.err
      cut -c1-420 gpurun_out/${tag}.jsonl; shift 2;;
    pytestk)       # a subset of the gpu tests: pytestk "<-k expression>"
      timeout 1500 python -m pytest tests -m gpu -q --maxfail 8 --timeout 900 -k "$2" > gpurun_out/${tag}_pytest_gpu_k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest_gpu_k.log
      tail -8 gpurun_out/${tag}_pytest_gpu_k.log; shift 2;;
    ncu)           # ncu <kernel regex> <launch count> <plan>: one --set full capture of the matching launches of a session plan
      timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:$2" -c $3 -o gpurun_out/${tag}_ncu -f python tools/gpu_session.py --plan $4 --tag ${tag}_ncu_run > gpurun_out/${tag}_ncu.log 2>&1
      tail -3 gpurun_out/${tag}_ncu.log
      ncu -i gpurun_out/${tag}_ncu.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_raw.csv 2>/dev/null; wc -c gpurun_out/${tag}_ncu_raw.csv
      shift 4;;
    launches)      # launches <plan>: per-launch durations of the enumerate-path kernels of a session plan (ncu launch list)
      timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:sweep_lean|sweep_kernel|sweep_guides|search_fast|search_kernel|scan_u32|scatter_matches|order_matches|expand_hits|locate_score|specificity|publish|total_u32|variant_|threshold|order_|DeviceRadixSort|DeviceScan" -c 3000 --csv --log-file gpurun_out/${tag}_launches.csv python tools/gpu_session.py --plan $2 --tag ${tag}_launches_run > gpurun_out/${tag}_launches.log 2>&1
      tail -2 gpurun_out/${tag}_launches.log | cut -c1-300; shift 2;;
    torchbench)    # torchbench <N> <steps> <warmup>: bench.py as the driver launches it on N GPUs
      timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $2 --steps $3 --warmup $4 > gpurun_out/${tag}_bench_n$2.json 2> gpurun_out/${tag}_bench_n$2.err
      tail -2 gpurun_out/${tag}_bench_n$2.err; cut -c1-900 gpurun_out/${tag}_bench_n$2.json; shift 4;;
    cfg5)          # cfg5 <N gpus> <kmers Mb> <csv guides>: tools/config5_run.py   (CFG5_FILE_BATCH: guides per device and batch, default 50000; 0 = library default)
      timeout 2400 python tools/config5_run.py --gpus $2 --kmers-mb $3 --csv-guides $4 --file-batch ${CFG5_FILE_BATCH:-50000} --out gpurun_out/${tag}_config5_$2gpu_b${CFG5_FILE_BATCH:-50000}.json > gpurun_out/${tag}_config5.log 2> gpurun_out/${tag}_config5.err
      tail -4 gpurun_out/${tag}_config5.err | cut -c1-600; shift 4;;
    entrysmoke)    # __graft_entry__.smoke() as the driver runs it
      timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log; tail -2 gpurun_out/${tag}_smoke.log; shift;;
    sanitize)      # sanitize <tool> "<-k expression>": a few small GPU tests under compute-sanitizer
      timeout 1500 compute-sanitizer --tool $2 --error-exitcode 99 --target-processes all python -m pytest tests -m gpu -q -x --timeout 1400 -k "$3" > gpurun_out/${tag}_sanitizer_$2.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/${tag}_sanitizer_$2.log
      grep -c "ERROR SUMMARY\|=========" gpurun_out/${tag}_sanitizer_$2.log; tail -4 gpurun_out/${tag}_sanitizer_$2.log | cut -c1-300; shift 3;;
    benchlaunches) # the launch list of two bench steps (per-launch durations under ncu; the kernel SHARES are what counts)
      # (only the kernels of the enumerate path are profiled: the index build in front of them is a thousand radix-sort launches)
      timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:sweep_lean|sweep_kernel|sweep_guides|search_fast|search_kernel|scan_u32|scatter_matches|order_matches|expand_hits|locate_score|specificity|publish|total_u32|variant_|threshold" -c 400 --csv --log-file gpurun_out/${tag}_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-file-e2e > gpurun_out/${tag}_bench_launches.log 2>&1
      tail -2 gpurun_out/${tag}_bench_launches.log | cut -c1-300; shift;;
    refformat)     # refformat <genome Mb>: tools/ref_format_roundtrip.py (GPU build -> reference-format files -> product and unmodified reference open them)
      timeout 1200 python tools/ref_format_roundtrip.py --genome-mb $2 --out gpurun_out/${tag}_reference_format_roundtrip_$2mb.json > gpurun_out/${tag}_refformat_$2.log 2> gpurun_out/${tag}_refformat_$2.err || tail -5 gpurun_out/${tag}_refformat_$2.err
      cut -c1-1500 gpurun_out/${tag}_refformat_$2.log; shift 2;;
    bench)
      shift
      timeout 1200 python bench.py "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -3 gpurun_out/${tag}_bench.err; cut -c1-1200 gpurun_out/${tag}_bench.json
      break;;
    *) echo "unknown step $1"; exit 2;;
  esac
done
