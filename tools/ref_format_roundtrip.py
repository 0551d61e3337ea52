#!/usr/bin/env python
"""Interop at the north-star size, on a GPU box: the 3.1 Gb bench genome indexed on the GPU, exported in the reference's own format
(gsx_index_save_reference_format -> <prefix>.forward / .reverse / .gs, reference src/guidescan.cxx:152-175), then

  * opened again by the product from those files (gsx_index_open, the reader of the reference's format): phases of the open, and the
    digest of the device arrays against the digest of the index that was built -- the same index;
  * opened and searched by the UNMODIFIED reference (oracle/_ref/guidescan enumerate, test infrastructure): its index load time, and
    its output for a few guides against the product's on the re-opened index.

  python tools/ref_format_roundtrip.py [--genome-mb 3100] [--guides 64] --out gpurun_out/r02r_reference_format_roundtrip.json
"""
import argparse
import datetime
import json
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "guidescan-cli_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome-mb", type=float, default=3100)
    ap.add_argument("--guides", type=int, default=64)
    ap.add_argument("--workdir", default="/tmp/gsx_ref_format")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "reference_format_roundtrip.json"))
    a = ap.parse_args()
    import bench
    import gsx
    import oracle as O
    os.makedirs(a.workdir, exist_ok=True)
    b = bench.parse_args([])
    b.genome_mb = a.genome_mb
    b.n_chr = 24 if a.genome_mb >= 1000 else 8
    g, chroms, pos, kmers = bench.make_workload(b, 1)
    res = {"genome_mb": a.genome_mb, "host_cores": os.cpu_count()}
    prefix = os.path.join(a.workdir, "ix")
    t0 = time.time()
    ix = gsx.Index.build_from_text(g, chroms, sa_shift=6, devices=[0])
    res["gpu_build"] = {"wall_s": time.time() - t0, "phases_s": ix.open_seconds()}
    digest = ix.device_checksum(0)
    t0 = time.time()
    ix.save_reference_format(prefix)
    res["export"] = {"seconds": time.time() - t0, "bytes": {e: os.path.getsize(prefix + "." + e) for e in ("forward", "reverse", "gs")}}
    ix.close()
    del g
    t0 = time.time()
    ix = gsx.Index.open(prefix)
    res["open_from_reference_files"] = {"wall_s": time.time() - t0, "phases_s": ix.open_seconds(), "device_arrays_equal_the_built_index": ix.device_checksum(0) == digest}
    gcsv = os.path.join(a.workdir, "guides.csv")
    bench.write_sample_csv(gcsv, kmers, a.guides)
    mine = os.path.join(a.workdir, "gpu.out")
    n, _ = ix.enumerate_file(gcsv, mine, gsx.make_params(mismatches=3))
    ix.close()
    if O.have_ref():
        theirs = os.path.join(a.workdir, "ref.out")
        t0 = time.time()
        p = subprocess.run([O.REF_BIN, "enumerate", prefix, "-f", gcsv, "-o", theirs, "-m", "3", "-n", str(os.cpu_count())], stdout=subprocess.PIPE, text=True)
        wall = time.time() - t0
        lines = p.stdout.splitlines()
        stamp = lambda l: datetime.datetime.strptime(re.match(r"\[([^\]]+)\]", l).group(1), "%Y-%m-%d %H:%M:%S.%f")
        first = [l for l in lines if re.match(r"\[", l)]
        loaded = [l for l in lines if "Successfully loaded genome index" in l]
        blocks = lambda path: sorted(open(path, "rb").read().split(b"\n"))
        res["unmodified_reference_on_the_exported_index"] = {
            "returncode": p.returncode, "guides": n, "wall_s": wall,
            "index_load_s": (stamp(loaded[0]) - stamp(first[0])).total_seconds() if loaded and first else None,
            "output_equals_the_product_s_sorted": p.returncode == 0 and blocks(theirs) == blocks(mine)}
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
