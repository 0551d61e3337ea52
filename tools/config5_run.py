#!/usr/bin/env python
"""BASELINE.json configs[4] end to end through the product's own multi-GPU path, on N GPUs of one box:

  genome-wide guide generation (gsx_generate_kmers, reference scripts/generate_kmers.py:55-136) on the 3.1 Gb synthetic genome
  -> `bin/guidescan enumerate --gpus N -m 4 -a NAG --format sam --mode complete` (reference src/guidescan.cxx:181-258,
  include/genomics/printer.hpp:115-170,302-360), one process, index replicated by peer copies, guides sharded inside gsx_enumerate.

The full job is 3.9e8 guides and ~2 TB of SAM text; this runs a BOUNDED sample of it -- the guides of the first --kmers-mb Mb of
chromosome 1 against the WHOLE genome index -- writes the SAM text to /dev/null (and the first --check-guides guides to a file that
is diffed against the CPU oracle), and reports per-stage rates from which the whole job extrapolates.  Also runs configs[2]
(m = 3, CSV) through the same binary on 1 and on N GPUs and checks that the two files are identical.

  python tools/config5_run.py --gpus 8 [--genome-mb 3100] [--kmers-mb 32] --out profiles/r02_config5_8gpu.json
"""
import argparse
import json
import os
import re
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "guidescan-cli_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
EXE = os.path.join(ROOT, "guidescan-cli_b200", "bin", "guidescan")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def run_cli(args, env=None):
    t0 = time.time()
    r = subprocess.run([EXE] + args, capture_output=True, text=True, env={**os.environ, "GSX_FILE_TIMING": "1", **(env or {})})
    dt = time.time() - t0
    if r.returncode:
        raise RuntimeError("guidescan %s failed: %s" % (" ".join(args), r.stderr[-800:]))
    timing = [l for l in r.stderr.splitlines() if l.startswith("gsx file job:")]
    m = re.search(r"Processed (\d+) kmers in \d+ seconds. \(([\d.]+) s, ([\d.]+) kmers/s; device ([\d.]+) ms", r.stdout)
    o = re.search(r"files ([\d.]+) s, device layout ([\d.]+) s, replication ([\d.]+) s", r.stdout)
    return {"wall_s": dt, "guides": int(m.group(1)), "enumerate_s": float(m.group(2)), "guides_per_s": float(m.group(3)), "device_ms_sum": float(m.group(4)),
            "index_open_s": {"files": float(o.group(1)), "device_layout": float(o.group(2)), "replication": float(o.group(3))} if o else None,
            "file_job_timing": timing[-1] if timing else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=8)
    ap.add_argument("--genome-mb", type=float, default=3100)
    ap.add_argument("--kmers-mb", type=float, default=32, help="guides are generated from the first this-many Mb of chromosome 1")
    ap.add_argument("--check-guides", type=int, default=200)
    ap.add_argument("--csv-guides", type=int, default=1000000, help="configs[2] leg: guides through the binary on 1 and on N GPUs (0 = skip)")
    ap.add_argument("--file-batch", type=int, default=50000, help="guides per device and batch of the whole-file driver (0 = the library's default, 200 000)")
    ap.add_argument("--workdir", default="/tmp/gsx_cfg5")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "config5.json"))
    a = ap.parse_args()
    import bench
    import gsx
    import synth
    import oracle as O
    os.makedirs(a.workdir, exist_ok=True)
    res = {"n_gpus": a.gpus, "genome_mb": a.genome_mb, "devices_visible": gsx.device_count(), "host_cores": os.cpu_count()}
    b = bench.parse_args([])
    b.genome_mb = a.genome_mb
    b.n_chr = 24 if a.genome_mb >= 1000 else 8
    g, chroms, pos, kmers = bench.make_workload(b, 1)
    prefix = os.path.join(a.workdir, "genome")
    t0 = time.time()
    ix = gsx.Index.build_from_text(g, chroms, sa_shift=2, devices=list(range(a.gpus)), save_prefix=prefix)
    res["index_build"] = {"wall_s": time.time() - t0, "seconds": ix.open_seconds(), "gb_per_device": ix.device_bytes / 1e9,
                          "replicas_equal": len({ix.device_checksum(s) for s in range(a.gpus)}) == 1}
    log("index:", res["index_build"])
    # ---- guide generation: the first kmers_mb of chromosome 1 as its own FASTA record (a sample of the genome-wide set) ----
    n_k = int(min(a.kmers_mb * 1e6, chroms[0][1]))
    fa = os.path.join(a.workdir, "chr1_head.fa")
    synth.write_fasta(fa, g[:n_k], [(chroms[0][0], n_k)])
    kcsv = os.path.join(a.workdir, "kmers.csv")
    t0 = time.time()
    n_kmers = gsx.generate_kmers(fa, kcsv, pam="NGG", kmer_length=20, device=0)
    dt = time.time() - t0
    res["generate_kmers"] = {"bases": n_k, "guides": n_kmers, "seconds": dt, "guides_per_s": n_kmers / dt, "csv_bytes": os.path.getsize(kcsv),
                             "whole_genome_guides_extrapolated": int(n_kmers * (len(g) / n_k))}
    log("kmers:", res["generate_kmers"])
    # ---- the CPU oracle on the first guides (over the FM-index exported from the GPU builder) ----
    t0 = time.time()
    b0, b1 = ix.export_bwt(0), ix.export_bwt(1)
    (s0, sh0), (s1, sh1) = ix.export_sa_samples(0), ix.export_sa_samples(1)
    oix = O.Index.from_bwt(b0, np.ascontiguousarray(s0[::1 << (6 - sh0)]), b1, np.ascontiguousarray(s1[::1 << (6 - sh1)]), chroms)
    del b0, b1
    ix.close()
    check_csv = os.path.join(a.workdir, "check.csv")
    with open(kcsv) as f, open(check_csv, "w") as o:
        for i, line in enumerate(f):
            if i > a.check_guides:
                break
            o.write(line)
    cpu_sam = os.path.join(a.workdir, "cpu.sam")
    t1 = time.time()
    oix.enumerate_file(O.make_opts(mismatches=4, alt_pams=("NAG",), fmt="sam"), check_csv, cpu_sam, nthreads=os.cpu_count())
    res["cpu_oracle"] = {"import_s": t1 - t0, "guides": a.check_guides, "seconds": time.time() - t1, "guides_per_s": a.check_guides / (time.time() - t1),
                         "threads": os.cpu_count(), "kind": "port (oracle/gs_oracle.c)"}
    # ---- configs[4]: the binary, all GPUs, SAM complete ----
    common = ["enumerate", prefix, "-m", "4", "--format", "sam", "--mode", "complete", "--gpus", str(a.gpus)]
    head_csv = os.path.join(a.workdir, "head.csv")                      # a batch large enough for the sweep on every device
    n_head = min(n_kmers, 25000 * a.gpus)
    with open(kcsv) as f, open(head_csv, "w") as o:
        for i, line in enumerate(f):
            if i > n_head:
                break
            o.write(line)
    gpu_sam = os.path.join(a.workdir, "gpu.sam")
    env5 = {"GSX_FILE_BATCH": str(a.file_batch)} if a.file_batch else {}          # per device: ~300 hits per guide at m = 4 with two PAMs
    res["file_batch_per_device"] = a.file_batch or 200000
    r_head = run_cli(common + ["-f", head_csv, "-o", gpu_sam, "-a", "NAG"], env=env5)
    want = open(cpu_sam, "rb").read()
    got = open(gpu_sam, "rb").read()
    res["config5_check"] = {**r_head, "sam_bytes": len(got), "parity_on_cpu_sample": got.startswith(want), "cpu_sample_bytes": len(want)}
    log("check:", res["config5_check"])
    os.remove(gpu_sam)
    r_all = run_cli(common + ["-f", kcsv, "-o", "/dev/null", "-a", "NAG"], env=env5)
    sam_per_guide = len(got) / max(1, n_head)
    res["config5"] = {**r_all, "output": "/dev/null", "sam_bytes_per_guide_measured_on_the_check_file": sam_per_guide,
                      "sam_gb_per_s_formatted": r_all["guides_per_s"] * sam_per_guide / 1e9,
                      "whole_genome_job_hours_extrapolated": res["generate_kmers"]["whole_genome_guides_extrapolated"] / r_all["guides_per_s"] / 3600.0,
                      "parity_on_cpu_sample": res["config5_check"]["parity_on_cpu_sample"]}
    log("config5:", res["config5"])
    # ---- configs[2]: m = 3, CSV, the same binary on one GPU and on all of them: identical files ----
    if a.csv_guides:
        gcsv = os.path.join(a.workdir, "m3.csv")
        bench.write_sample_csv(gcsv, kmers, min(a.csv_guides, len(kmers)))
        outs = {}
        for n in sorted({1, a.gpus}):
            out = os.path.join(a.workdir, "m3_%d.out" % n)
            r = run_cli(["enumerate", prefix, "-m", "3", "--gpus", str(n), "-f", gcsv, "-o", out])
            outs[n] = out
            res["config3_cli_%dgpu" % n] = {**r, "csv_bytes": os.path.getsize(out)}
            log("config3 cli:", n, res["config3_cli_%dgpu" % n])
        if a.gpus > 1:
            res["config3_cli_outputs_identical"] = open(outs[1], "rb").read() == open(outs[a.gpus], "rb").read()
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
