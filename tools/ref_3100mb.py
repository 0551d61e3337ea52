#!/usr/bin/env python
"""CPU-container job (no GPU): the UNMODIFIED reference at the north-star size.

Builds bench.py's own 3.1 Gb workload (same seed, same planting), runs `oracle/_ref/guidescan index` on it, times
`oracle/_ref/guidescan enumerate -n <cores>` on the first --sample guides, and writes the figures to
profiles/r02_reference_3100mb.json (BASELINE.md 2.1 is filled from it).  The index files stay in --workdir for
tools/ref_index_check (parser of the reference's files vs the genome text) and for the port's timing on the same index.

  python tools/ref_3100mb.py [--workdir /tmp/ref3100] [--sample 8000] [--stage fasta|index|enumerate|all]
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "guidescan-cli_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workdir", default="/tmp/ref3100")
    ap.add_argument("--sample", type=int, default=8000)
    ap.add_argument("--stage", default="all")
    ap.add_argument("--mismatches", type=int, default=3)
    a = ap.parse_args()
    import bench
    import oracle as O
    import synth
    args = bench.parse_args([])                    # bench.py's defaults = the graded workload
    os.makedirs(a.workdir, exist_ok=True)
    tag = "bench_%dmb_s%d" % (int(args.genome_mb), args.seed)
    fa, prefix = os.path.join(a.workdir, tag + ".fa"), os.path.join(a.workdir, tag)
    gcsv = os.path.join(a.workdir, "cpu_sample.csv")
    res_path = os.path.join(ROOT, "profiles", "r02_reference_3100mb.json")
    res = json.load(open(res_path)) if os.path.exists(res_path) else {}
    res.update({"workload": bench.workload_config(args)["workload"], "cores": os.cpu_count(),
                "cpu": [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]})
    if a.stage in ("all", "fasta") and not os.path.exists(fa):
        g, chroms, pos, kmers = bench.make_workload(args, 1)
        synth.write_fasta(fa, g, chroms)
        bench.write_sample_csv(gcsv, kmers, a.sample)
        g.tofile(os.path.join(a.workdir, tag + ".text"))           # the planted genome bytes, for ref_index_check
        del g
    if a.stage in ("all", "index") and not os.path.exists(prefix + ".reverse"):
        t0 = time.time()
        p = subprocess.run([O.REF_BIN, "index", "--index", prefix, fa], cwd=a.workdir,
                           stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        dt = time.time() - t0
        import resource
        res["index_seconds"] = dt
        res["index_peak_rss_kb"] = resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss      # peak resident set of the single-threaded construction
        res["index_bytes"] = {s: os.path.getsize(prefix + "." + s) for s in ("forward", "reverse")}
        json.dump(res, open(res_path, "w"), indent=1)
        if p.returncode:
            sys.exit("reference index failed: " + p.stderr[-400:])
    if a.stage in ("all", "enumerate"):
        out = os.path.join(a.workdir, "ref.out")
        cores = os.cpu_count()
        t0 = time.time()
        p = subprocess.run([O.REF_BIN, "enumerate", prefix, "-f", gcsv, "-o", out, "-m", str(a.mismatches), "-n", str(cores)],
                           stdout=subprocess.PIPE, text=True)
        t1 = time.time()
        # the log line "Successfully loaded genome index" separates index load from the search
        lines = p.stdout.splitlines()
        import datetime
        import re
        stamp = lambda l: datetime.datetime.strptime(re.match(r"\[([^\]]+)\]", l).group(1), "%Y-%m-%d %H:%M:%S.%f")
        loaded = [l for l in lines if "Successfully loaded genome index" in l]
        done = [l for l in lines if "Processed" in l and "seconds" in l]
        res["enumerate"] = {"guides": a.sample, "threads": cores, "mismatches": a.mismatches, "wall_seconds": t1 - t0,
                            "search_seconds": (stamp(done[-1]) - stamp(loaded[0])).total_seconds() if loaded and done else None,
                            "log": [l for l in lines if "Processed:" not in l][-12:], "rows": sum(1 for _ in open(out)) - 1}
        json.dump(res, open(res_path, "w"), indent=1)
    if a.stage in ("all", "port"):
        # the CPU port on the SAME index: BWT and SA samples as the product's parser extracted them from the reference's files
        # (tests/_build/ref_index_check <prefix> <text> --dump <workdir>/dump), same guides, same thread count
        import numpy as np
        d = os.path.join(a.workdir, "dump")
        if os.path.exists(os.path.join(d, "bwt.forward")):
            chroms = synth.chromosome_table(int(args.genome_mb * 1e6), args.n_chr)
            b0, b1 = np.fromfile(os.path.join(d, "bwt.forward"), dtype=np.uint8), np.fromfile(os.path.join(d, "bwt.reverse"), dtype=np.uint8)
            s0, s1 = np.fromfile(os.path.join(d, "sa64.forward"), dtype=np.uint32), np.fromfile(os.path.join(d, "sa64.reverse"), dtype=np.uint32)
            t0 = time.time()
            oix = O.Index.from_bwt(b0, s0, b1, s1, chroms)
            t1 = time.time()
            out = os.path.join(a.workdir, "port.out")
            oix.enumerate_file(O.make_opts(mismatches=a.mismatches), gcsv, out, nthreads=os.cpu_count())
            t2 = time.time()
            ref_lines = sorted(open(os.path.join(a.workdir, "ref.out"), "rb").read().split(b"\n"))
            port_lines = sorted(open(out, "rb").read().split(b"\n"))
            res["port"] = {"import_seconds": t1 - t0, "enumerate_seconds": t2 - t1, "guides_per_s": a.sample / (t2 - t1), "threads": os.cpu_count(),
                           "output_equals_reference_sorted": ref_lines == port_lines,
                           "port_over_reference_speed": res["enumerate"]["search_seconds"] / (t2 - t1) if res.get("enumerate", {}).get("search_seconds") else None}
            json.dump(res, open(res_path, "w"), indent=1)
    if a.stage in ("all", "rewrite"):
        # the reference's own files through the product's parser (load_sdsl_strand) and back out through its writer of the reference
        # format (gsx_sdsl_write.cpp): identical files; their sha256 go to tests/golden (the GPU-built index must hash the same)
        import hashlib
        tool = os.path.join(ROOT, "tests", "_build", "sdsl_write_check")
        rw = {"cores": os.cpu_count()}
        for strand in ("forward", "reverse"):
            src, out = prefix + "." + strand, os.path.join(a.workdir, "rewritten." + strand)
            p = subprocess.run([tool, "rewrite", src, out], capture_output=True, text=True)
            if p.returncode:
                sys.exit("rewrite failed: " + p.stderr[-400:])
            h = [hashlib.sha256(), hashlib.sha256()]
            for k, path in enumerate((src, out)):
                with open(path, "rb") as f:
                    for chunk in iter(lambda: f.read(1 << 24), b""):
                        h[k].update(chunk)
            rw[strand] = {**json.loads(p.stdout), "bytes": os.path.getsize(src), "sha256_reference_file": h[0].hexdigest(),
                          "rewritten_file_identical": h[0].digest() == h[1].digest() and os.path.getsize(out) == os.path.getsize(src)}
            os.remove(out)
        rw["gs_sha256"] = hashlib.sha256(open(prefix + ".gs", "rb").read()).hexdigest()
        out_path = os.path.join(ROOT, "profiles", "r02q_reference_format_3100mb.json")
        json.dump({"what": "profiles/r02_reference_3100mb.json's index files (unmodified reference, 4 091 s) parsed by load_sdsl_strand and written back "
                           "by save_sdsl_strand (tests/sdsl_write_check rewrite), one strand at a time, all cores", **rw}, open(out_path, "w"), indent=1)
        res["rewrite"] = rw
    print(json.dumps(res))


if __name__ == "__main__":
    main()
