#!/usr/bin/env python
"""One GPU process, one index build, many measurements: runs the experiments of a plan file against a shared 3.1 Gb (or
smaller) index and appends one JSON line per experiment to gpurun_out/<tag>.jsonl.  An index build costs ~40 s of box
time, a bench.py process ~2 min with its imports and genome: sessions that compare options go through here instead.

  python tools/gpu_session.py --plan tools/plans/r2a.json --tag r2a [--genome-mb 3100]

Plan: {"genomes": [{"n_runs": 0, "devices": [0], "experiments": [ {...}, ... ]}, ...]}
Experiment keys (all optional but name): name, env {VAR: value}, mismatches, alt_pams [..], rna_bulges, dna_bulges, guides
(per step), steps, warmup, pipeline (batches in flight through gsx_enumerate_start / _wait; 0 = plain gsx_enumerate calls),
parity_sample (guides diffed byte for byte against the CPU oracle over the exported FM-index), file_e2e {"guides": n, "fmt":
"csv"|"sam"} (whole-file driver instead of the array path), open_again (re-open timing from a saved .gsx on "devices").
Genome keys: n_runs, skew, devices, sa_shift, save_prefix (also write <workdir>/<prefix>.gsx), open_prefix (open that file instead of building).
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "guidescan-cli_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class Session:
    def __init__(self, args, genome_spec):
        import bench
        import gsx
        import synth
        self.bench, self.gsx, self.synth = bench, gsx, synth
        self.args = args
        b = bench.parse_args([])
        b.genome_mb = args.genome_mb; b.n_runs = genome_spec.get("n_runs", 0); b.seed = args.seed
        b.skew = bool(genome_spec.get("skew", False)); b.skew_lowc_mb = genome_spec.get("skew_lowc_mb", b.skew_lowc_mb)
        b.n_chr = 24 if args.genome_mb >= 1000 else 8
        b.guides_per_step = 200000; b.steps = 5; b.warmup = 3
        self.bargs = b
        self.devices = genome_spec.get("devices", [0])
        t0 = time.time()
        self.g, self.chroms, self.pos, self.kmers = bench.make_workload(b, 1)
        log("genome: %.1f s" % (time.time() - t0))
        t0 = time.time()
        self.save_prefix = genome_spec.get("save_prefix")
        if genome_spec.get("open_prefix"):       # an index saved by an earlier process (no suffix sorting in this one: profiles stay clean)
            self.ix = gsx.Index.open(genome_spec["open_prefix"], devices=self.devices)
        else:
            self.ix = gsx.Index.build_from_text(self.g, self.chroms, sa_shift=genome_spec.get("sa_shift", 2), devices=self.devices,
                                                save_prefix=self.save_prefix)
        self.index_s = time.time() - t0
        self.open_s = self.ix.open_seconds()
        log("index: %.1f s %s, %.2f GB/device on %s" % (self.index_s, self.open_s, self.ix.device_bytes / 1e9, self.devices))
        self.oix = None

    def oracle_index(self):
        if self.oix is None:
            import oracle as O
            t0 = time.time()
            b0, b1 = self.ix.export_bwt(0), self.ix.export_bwt(1)
            (s0, sh0), (s1, sh1) = self.ix.export_sa_samples(0), self.ix.export_sa_samples(1)
            s0, s1 = np.ascontiguousarray(s0[::1 << (6 - sh0)]), np.ascontiguousarray(s1[::1 << (6 - sh1)])
            self.oix = O.Index.from_bwt(b0, s0, b1, s1, self.chroms)
            log("oracle index imported in %.1f s" % (time.time() - t0))
        return self.oix

    def guide_array(self, lo, n):
        gsx = self.gsx
        seqs = [self.kmers[i, :20].tobytes() for i in range(lo, lo + n)]
        arr = (gsx.Guide * n)()
        for i, sq in enumerate(seqs):
            arr[i] = gsx.Guide(sq, b"NGG")
        return arr, seqs

    def run(self, e):
        gsx, bench = self.gsx, self.bench
        saved = {k: os.environ.get(k) for k in e.get("env", {})}
        os.environ.update({k: str(v) for k, v in e.get("env", {}).items()})
        try:
            return self._run(e)
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v

    def _run(self, e):
        if e.get("via_devices"):                 # the same experiment through an index re-opened from the saved .gsx on these devices
            keep = self.ix
            self.ix = self.gsx.Index.open(self.save_prefix, devices=e["via_devices"])
            try:
                e2 = {k: v for k, v in e.items() if k != "via_devices"}
                out = self._run(e2)
                out["via_devices"] = e["via_devices"]
                return out
            finally:
                self.ix.close()
                self.ix = keep
        gsx, bench = self.gsx, self.bench
        m = e.get("mismatches", 3)
        params = gsx.make_params(mismatches=m, alt_pams=tuple(e.get("alt_pams", ())), rna_bulges=e.get("rna_bulges", 0),
                                 dna_bulges=e.get("dna_bulges", 0))
        out = {"name": e["name"], "env": e.get("env", {}), "mismatches": m, "alt_pams": e.get("alt_pams", []),
               "rna_bulges": e.get("rna_bulges", 0), "dna_bulges": e.get("dna_bulges", 0), "devices": self.devices,
               "n_runs": self.bargs.n_runs, "skew": self.bargs.skew, "genome_mb": self.bargs.genome_mb}
        if e.get("open_again"):
            devs = e.get("devices", self.devices)
            t0 = time.time()
            ix2 = gsx.Index.open(self.save_prefix, devices=devs)
            out.update({"open_wall_s": time.time() - t0, "open_seconds": ix2.open_seconds(), "open_devices": devs,
                        "checksums_equal": len({ix2.device_checksum(s) for s in range(len(devs))}) == 1,
                        "same_as_built": ix2.device_checksum(0) == self.ix.device_checksum(0)})
            ix2.close()
            return out
        if e.get("file_e2e"):
            fe = e["file_e2e"]
            n = fe["guides"]
            wd = self.args.workdir
            gcsv, fo = os.path.join(wd, "s_file.csv"), fe.get("out", os.path.join(wd, "s_file.out"))
            bench.write_sample_csv(gcsv, self.kmers, n)
            self.ix.enumerate_file(gcsv, fo, params, fmt=fe.get("fmt", "csv"))
            t0 = time.perf_counter()
            _, ctr = self.ix.enumerate_file(gcsv, fo, params, fmt=fe.get("fmt", "csv"))
            dt = time.perf_counter() - t0
            size = os.path.getsize(fo) if os.path.isfile(fo) else 0
            out.update({"file_guides": n, "file_seconds": dt, "file_guides_per_s": n / dt, "file_bytes": size, "file_mb_per_s": size / dt / 1e6,
                        "device_ms": ctr["ms_total_device"], "fmt": fe.get("fmt", "csv")})
            if os.path.isfile(fo):
                os.remove(fo)
            return out
        per, steps, warm = e.get("guides", 200000), e.get("steps", 3), e.get("warmup", 2)
        arrs = [self.guide_array(s * per, per) for s in range(steps + warm)]
        for s in range(warm):
            self.ix.enumerate_raw(arrs[s][0], per, params).close()
        dev_ms = 0.0
        tot = {}
        spec = 0.0

        def consume(r):
            nonlocal dev_ms, spec
            c = r.counters()
            dev_ms += c["ms_total_device"]
            for k, v in c.items():
                tot[k] = tot.get(k, 0) + v
            spec += float(r.guide_arrays()["specificity"].sum())
            r.close()

        depth = e.get("pipeline", 0)
        t0 = time.perf_counter()
        if depth <= 0:
            for s in range(warm, warm + steps):
                consume(self.ix.enumerate_raw(arrs[s][0], per, params))
        else:
            pend = []
            for s in range(warm, warm + steps):
                pend.append(self.ix.enumerate_start(arrs[s][0], per, params))
                if len(pend) >= depth:
                    consume(self.ix.enumerate_wait(pend.pop(0)))
            while pend:
                consume(self.ix.enumerate_wait(pend.pop(0)))
        wall = time.perf_counter() - t0
        n = per * steps
        out.update({"guides_per_step": per, "steps": steps, "pipeline": depth, "guides_per_s_device": n / (dev_ms * 1e-3),
                    "guides_per_s_e2e": n / wall, "ms_per_step_device": dev_ms / steps, "ms_per_step_e2e": wall * 1e3 / steps,
                    "specificity_sum": spec,
                    "counters": {k: (v / steps) for k, v in tot.items()}})
        ps = e.get("parity_sample", 0)
        if ps:
            import oracle as O
            wd = self.args.workdir
            gcsv, co, go = os.path.join(wd, "s_par.csv"), os.path.join(wd, "s_par.cpu"), os.path.join(wd, "s_par.gpu")
            bench.write_sample_csv(gcsv, self.kmers, ps)
            oix = self.oracle_index()
            t0 = time.time()
            oix.enumerate_file(O.make_opts(mismatches=m, alt_pams=tuple(e.get("alt_pams", ())), rna_bulges=e.get("rna_bulges", 0),
                                           dna_bulges=e.get("dna_bulges", 0), fmt=e.get("parity_fmt", "csv")), gcsv, co, nthreads=os.cpu_count())
            cpu_s = time.time() - t0
            # the same guides through the GPU path in a batch large enough for the same kernels as the timed run
            big = os.path.join(wd, "s_par_big.csv")
            bench.write_sample_csv(big, self.kmers, max(ps, min(per, e.get("parity_batch", per))))
            self.ix.enumerate_file(big, go, params, fmt=e.get("parity_fmt", "csv"))
            a = open(co, "rb").read()
            b = open(go, "rb").read()
            # the GPU file holds more guides than the sample: compare the sample's prefix (guides are written in input order)
            out.update({"parity_on_cpu_sample": b.startswith(a), "parity_sample": ps, "cpu_guides_per_s": ps / cpu_s, "cpu_cores": os.cpu_count(),
                        "parity_bytes": len(a)})
        return out

    def close(self):
        self.ix.close()
        if self.oix is not None:
            self.oix.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--plan", required=True)
    ap.add_argument("--tag", required=True)
    ap.add_argument("--genome-mb", type=float, default=3100)
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--workdir", default="/tmp/gsx_session")
    args = ap.parse_args()
    os.makedirs(args.workdir, exist_ok=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    plan = json.load(open(args.plan))
    path = os.path.join(ROOT, "gpurun_out", args.tag + ".jsonl")
    with open(path, "a") as f:
        for gspec in plan["genomes"]:
            for key in ("save_prefix", "open_prefix"):
                if gspec.get(key):
                    gspec[key] = os.path.join(args.workdir, gspec[key])
            s = Session(args, gspec)
            f.write(json.dumps({"name": "_index", "n_runs": gspec.get("n_runs", 0), "index_wall_s": s.index_s, "open_seconds": s.open_s,
                                "device_gb": s.ix.device_bytes / 1e9, "devices": s.devices}) + "\n"); f.flush()
            for e in gspec["experiments"]:
                t0 = time.time()
                try:
                    r = s.run(e)
                except Exception as ex:          # one failing experiment must not cost the rest of the box time
                    r = {"name": e.get("name"), "error": repr(ex)[:500]}
                r["wall_s"] = time.time() - t0
                f.write(json.dumps(r) + "\n"); f.flush()
                log(json.dumps({k: v for k, v in r.items() if k != "counters"})[:600])
            s.close()


if __name__ == "__main__":
    main()
