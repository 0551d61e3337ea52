#!/usr/bin/env python
"""bench.py -- guides/s of the off-target enumeration hot path (BASELINE.json metric) on N B200s of one node.

A step = one pass of the hot path (search both strand indexes + locate + coordinates + CFD + specificity) over one
batch of synthetic guides per GPU.  One process per GPU (torchrun), index replicated, guides sharded, no data-path
collective ("scaling": "weak": the per-GPU batch is fixed).  Timing: CUDA events inside the library for the
device-resident number (`value`), wall clock bracketed by barrier + synchronize for the end-to-end number through the
C ABI with host buffers (`e2e`), max over ranks.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl gsx|reference] [--genome-mb MB] [--guides-per-step G]
                  [--mismatches M] [--alt-pam NAG] [--rna-bulges R] [--dna-bulges D] [--n-runs K] [--skew] [--e2e-pipeline P]

Side blocks of the line: `roofline` (algorithmic sector bytes over the search launches' time against the measured streaming peak, with the
DRAM traffic of the committed ncu capture and the commit it was taken at, and the slice-gather ceiling that applies to a slice-major
sweep), `cpu_baseline` + `parity_on_cpu_sample` (the CPU arm on the first guides of the workload, its text diffed against the GPU arm's),
`file_e2e` (guides CSV in, CSV text out through gsx_enumerate_file; also to /dev/null), `clocks` (nvidia-smi during the timed region).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "guidescan-cli_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

METRIC = "guides/sec at k=3 mismatches"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs, streaming copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(args):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture of this same workload."""
    p = os.path.join(ROOT, "profiles", "search_traffic.json")
    key = "%dmb_%d_m%d" % (int(args.genome_mb), args.guides_per_step, args.mismatches)
    if os.path.exists(p):
        return json.load(open(p)).get(key)
    return None


def slice_gather_ceiling():
    """G sectors/s of dependent-free random 32-byte gathers inside 8 MB slices (tools/slice_gather.cu, profiles/r01j_slice_gather.jsonl)"""
    p = os.path.join(ROOT, "profiles", "r01j_slice_gather.jsonl")
    best = None
    if os.path.exists(p):
        for line in open(p):
            try:
                d = json.loads(line)
            except ValueError:
                continue
            if d.get("slice_mb") == 8 and d.get("sectors_per_lookup") == 1:
                best = max(best or 0.0, d["glookups_per_s"])
    return best


def random_gather_peak():
    p = os.path.join(ROOT, "profiles", "random_gather_peak.json")
    if os.path.exists(p):
        return json.load(open(p))
    return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, gpu_index=0):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx = max(mx, float(s[1]))
                for n, v in zip(names, s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_setup(n_gpus):
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        # NCCL carries the barrier and the timing reductions only.  Its INFO lines would land on stdout next to the JSON line (the
        # version banner does even with NCCL_DEBUG_FILE set), so the default level is WARN; a level set by the caller is respected, and
        # GSX_NCCL_DEBUG=INFO sends the INFO lines (rank count, transports) to a side file
        if "NCCL_DEBUG" not in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ.get("GSX_NCCL_DEBUG", "WARN")
            if os.environ["NCCL_DEBUG"] != "WARN":
                os.makedirs(os.environ.get("GSX_BENCH_DIR", "/tmp/gsx_bench"), exist_ok=True)
                os.environ.setdefault("NCCL_DEBUG_FILE", os.path.join(os.environ.get("GSX_BENCH_DIR", "/tmp/gsx_bench"), "nccl_%h_%p.log"))
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        if torch.cuda.is_available():
            dist_mod.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
        else:
            dist_mod.init_process_group(backend="gloo")
        dist = dist_mod
    return rank, world, local, dist


def bind_to_gpu_numa_node(local):
    """One process per GPU: keep the process -- and with it the pinned host buffers the library allocates for results -- on the NUMA node
    the GPU hangs off, so that result copies of 8 ranks do not cross the socket interconnect.  Best effort; returns what it did."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True,
                             timeout=20).stdout.strip().lower()
        if not bus:
            return None
        bus = bus[-12:] if len(bus) > 12 else bus                              # sysfs spells the domain with four digits
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return {"node": node, "bound": False}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if len(cpus) < 2:
            return {"node": node, "bound": False}
        os.sched_setaffinity(0, cpus)
        return {"node": node, "bound": True, "cpus": len(cpus)}
    except Exception as e:
        return {"error": repr(e)[:120]}


def barrier_sync(dist, local):
    import torch
    if torch.cuda.is_available():
        torch.cuda.synchronize(local)
    if dist is not None:
        dist.barrier()


def reduce_max(dist, local, x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=("cuda:%d" % local) if torch.cuda.is_available() else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(dist, local, x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=("cuda:%d" % local) if torch.cuda.is_available() else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def shard_guides(n_total, rank, world):
    """Contiguous shard of the guide list for `rank` (the multi-GPU partition; no collective on the data path)."""
    return n_total * rank // world, n_total * (rank + 1) // world


def make_workload(args, world):
    import synth
    G = int(args.genome_mb * 1e6)
    t0 = time.time()
    g = synth.make_genome(G, args.seed)
    n_total = args.guides_per_step * world * (args.steps + args.warmup)
    pos, kmers = synth.sample_guides(g, n_total, args.seed)
    n_plant = min(n_total, args.plant_guides)
    synth.plant(g, kmers[:n_plant], args.seed)
    if args.skew:        # SURVEY 8(d) skew stressor: repeat families and low-complexity tracts, 1 % of the guides drawn from each
        kmers, fam, low = synth.add_skew(g, kmers, args.seed, frac=args.skew_frac, lowc_mb=args.skew_lowc_mb)
        log("skew: %d repeat-family guides, %d low-complexity guides, %.1f Mb of tracts" % (len(fam), len(low), args.skew_lowc_mb))
    if args.n_runs:      # runs of N (every real assembly has them): the index then holds exception rows beyond the sentinel
        rng = np.random.default_rng(args.seed + 3000003)
        for _ in range(args.n_runs):
            at = int(rng.integers(1000, G - 100000))
            g[at:at + int(rng.integers(1, 50000))] = ord("N")
    chroms = synth.chromosome_table(G, args.n_chr)
    log("workload: genome %.1f Mb, %d guides total (%d with planted copies) in %.1f s" % (G / 1e6, n_total, n_plant, time.time() - t0))
    return g, chroms, pos, kmers


def build_index(gsx, g, chroms, local, args, workdir):
    t0 = time.time()
    try:
        ix = gsx.Index.build_from_text(g, chroms, sa_shift=args.sa_shift, devices=[local])
        how = "gpu-built"
    except gsx.GsxError as e:
        if "not implemented" not in str(e):
            raise
        import oracle as O
        import synth
        tag = "bench_%dmb_s%d" % (int(args.genome_mb), args.seed)
        fa = os.path.join(workdir, tag + ".fa")
        if not os.path.exists(os.path.join(workdir, tag + ".reverse")):
            synth.write_fasta(fa, g, chroms)
            O.ref_index(fa, os.path.join(workdir, tag), cwd=workdir)
        ix = gsx.Index.open(os.path.join(workdir, tag), devices=[local])
        how = "reference-built (guidescan index), converted"
    log("index: %s in %.1f s, %.2f GB on device" % (how, time.time() - t0, ix.device_bytes / 1e9))
    return ix, how


def run_gsx(args):
    import gsx
    rank, world, local, dist = dist_setup(args.gpus)
    numa = bind_to_gpu_numa_node(local) if (world > 1 and not args.no_numa_bind) else None
    workdir = args.workdir
    os.makedirs(workdir, exist_ok=True)
    g, chroms, pos, kmers = make_workload(args, world)
    ix, how = build_index(gsx, g, chroms, local, args, workdir)
    params = gsx.make_params(mismatches=args.mismatches, alt_pams=tuple(args.alt_pam), rna_bulges=args.rna_bulges, dna_bulges=args.dna_bulges)
    per = args.guides_per_step
    # host buffers of every step's guides for this rank (pinned memory is allocated inside the library for results)
    steps = []
    for s in range(args.steps + args.warmup):
        lo = (s * world + rank) * per
        seqs = [kmers[i, :20].tobytes() for i in range(lo, lo + per)]
        arr = (gsx.Guide * per)()
        for i, sq in enumerate(seqs):
            arr[i] = gsx.Guide(sq, b"NGG")
        steps.append((arr, seqs))
    h2d_bytes = per * 80
    if args.sweep_variants and rank == 0:
        # kernel-variant sweep on the first step's guides (diagnostic lines on stderr; not the bench value)
        # items: "f<k>" = specialised kernel variant k, "g<k>" = general kernel variant k
        for item in args.sweep_variants.split(","):
            name, _, pin = item.partition("@")          # "f1@64": specialised kernel variant 1 with a 64 MB L2 residency budget
            if name[0] == "s":                          # "s5v2": slice-major front end, 5 slice characters, sweep kernel variant 2; "s0": off
                sb, _, var = name[1:].partition("v")
                var, _, lmode = var.partition("l")          # "s5v2l1": ... summary loads with cache policy 1 (GSX_SWEEP_LOAD)
                os.environ["GSX_SWEEP_LOAD"] = lmode or "2"
                var, _, parts = var.partition("p")          # "s5v2p4": ... each (slice, 32 guides) unit cut into 4 work units
                env = {"GSX_FORCE_GENERAL": "0", "GSX_SWEEP": "0" if sb == "0" else "1", "GSX_SWEEP_SB": sb, "GSX_SWEEP_VARIANT": var or "0",
                       "GSX_SWEEP_PARTS": parts or "0"}
            else:
                env = {"GSX_FORCE_GENERAL": "1", "GSX_SEARCH_VARIANT": name[1:]} if name[0] == "g" else {"GSX_FORCE_GENERAL": "0", "GSX_FAST_VARIANT": name[1:]}
            if pin:
                env["GSX_L2_PIN_MB"] = pin
            else:
                os.environ.pop("GSX_L2_PIN_MB", None)
            os.environ.update(env)
            best = None
            for rep in range(3):
                r = ix.enumerate_raw(steps[0][0], per, params); c = r.counters(); r.close()
                best = c if best is None or c["ms_search"] < best["ms_search"] else best
            log(json.dumps({"variant": item, "ms_search": best["ms_search"], "guides_per_s_search": per / best["ms_search"] * 1e3,
                            "glookups_per_s": best["lookups"] / best["ms_search"] / 1e6, "spills": best["spills"], "nodes": best["nodes"],
                            "ms_sweep": best["ms_sweep"], "seeds": best["seeds"], "lookups": best["lookups"]}))
        for k in ("GSX_FORCE_GENERAL", "GSX_SEARCH_VARIANT", "GSX_FAST_VARIANT", "GSX_L2_PIN_MB", "GSX_SWEEP", "GSX_SWEEP_SB", "GSX_SWEEP_VARIANT", "GSX_SWEEP_PARTS", "GSX_SWEEP_LOAD"):
            os.environ.pop(k, None)
        apply_variant(args)
    for s in range(args.warmup):
        ix.enumerate_raw(steps[s][0], per, params).close()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier_sync(dist, local)
    t0 = time.perf_counter()
    dev_ms = search_ms = 0.0
    ctr_tot = {}
    d2h_bytes = 0
    timed = list(range(args.warmup, args.warmup + args.steps))
    # the public call in its two-slot form (gsx_enumerate_start / gsx_enumerate_wait): up to --e2e-pipeline batches in flight, so
    # that batch k's copies to the host and the read of its result run under batch k+1's kernels (the device section of the
    # calls takes turns inside the library); 1 = one plain call after the other
    pend = []

    def finish(h):
        nonlocal dev_ms, search_ms, d2h_bytes
        r = ix.enumerate_wait(h)
        c = r.counters()
        float(r.guide_arrays()["specificity"].sum())                      # the step's result is read on the host
        dev_ms += c["ms_total_device"]; search_ms += c["ms_search"]
        for k, v in c.items():
            ctr_tot[k] = ctr_tot.get(k, 0) + v
        # per guide: dropped, hit offset, hit count, specificity, perfect-match flag, counts per distance; per hit: abs_pos, sa_row, chr, pos1,
        # strand / distance / rna / dna / index_id / counted, cfd, match-string key (two words with bulges) and length
        d2h_bytes = r.n_guides * (1 + 4 + 4 + 4 + 1 + 4 * (args.mismatches + 1)) + r.n_hits * (8 + 4 + 4 + 4 + 6 + 4 + 8 + 1 + (8 if (args.rna_bulges or args.dna_bulges) else 0))
        r.close()

    for s in timed:
        pend.append(ix.enumerate_start(steps[s][0], per, params))
        if len(pend) >= max(1, args.e2e_pipeline):
            finish(pend.pop(0))
    while pend:
        finish(pend.pop(0))
    barrier_sync(dist, local)
    e2e_s = time.perf_counter() - t0
    clocks = sampler.finish() if sampler else None
    e2e_s = reduce_max(dist, local, e2e_s)
    dev_ms = reduce_max(dist, local, dev_ms)
    search_ms_max = reduce_max(dist, local, search_ms)
    lookups = reduce_sum(dist, local, ctr_tot["lookups"])
    nodes = reduce_sum(dist, local, ctr_tot["nodes"])
    hits = reduce_sum(dist, local, ctr_tot["hits"])
    total_guides = per * world * args.steps
    if rank == 0:
        peak, peak_src = peaks()
        # algorithmic bytes = the 32-byte index sectors the search kernels actually request (pattern summaries in the sweep,
        # occurrence blocks / look-ahead lines in the tree search); the same work in the reference's unit (occurrence
        # lookups of its own traversal) is reported beside it
        alg_bytes_per_launch = ctr_tot["sectors"] * 32.0 / args.steps
        launch_ms = search_ms / args.steps
        achieved = alg_bytes_per_launch / (launch_ms * 1e-3) / 1e9
        rg = random_gather_peak()
        tr = measured_traffic(args)
        sg = slice_gather_ceiling()
        line = {
            "metric": METRIC, "value": total_guides / (dev_ms * 1e-3), "unit": "guides/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {**workload_config(args),
                       "index": how, "sa_sample_rows": 1 << args.sa_shift, "parallelism": "guides sharded x%d, index replicated" % world, "e2e_batches_in_flight": max(1, args.e2e_pipeline), "numa_binding_rank0": numa,
                       "plant_guides": min(per * world * (args.steps + args.warmup), args.plant_guides), "index_open_seconds": list(ix.open_seconds()),
                       "l2": "index (%.2f GB) is far larger than L2; every step uses new guides" % (ix.device_bytes / 1e9)},
            "e2e": {"value": total_guides / e2e_s, "unit": "guides/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(d2h_bytes),
                    "ms_per_step": e2e_s * 1e3 / args.steps},
            "gpu_launches": int(ctr_tot["launches"]),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": tr["dram_bytes_per_launch"] if tr else None,
                         "traffic_source": tr["source"] if tr else None,
                         "traffic_capture_commit": tr.get("capture_commit") if tr else None,
                         "kernel": (tr["kernel"] if tr and ctr_tot["seeds"] else ("sweep_lean_kernel + search_fast_kernel" if ctr_tot["seeds"] else "search_fast_kernel")) if not os.environ.get("GSX_FORCE_GENERAL", "0") == "1" else "search_kernel", "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes_per_launch, "sectors_per_guide": ctr_tot["sectors"] / (per * args.steps),
                         "reference_unit_lookups_per_guide": lookups / total_guides,
                         "nodes_per_guide": nodes / total_guides, "launch_ms": launch_ms,
                         "random_sector_peak_gbs": rg["gb_per_s"] if rg else None,
                         "frac_of_random_sector_peak": (achieved / rg["gb_per_s"]) if rg else None,
                         # the ceiling that applies to a slice-major sweep: independent 32-byte gathers confined to an L2-sized slice
                         "sweep_gsectors_per_s": ctr_tot["sectors"] / args.steps / (ctr_tot["ms_sweep"] / args.steps * 1e-3) / 1e9 if ctr_tot["ms_sweep"] else None,
                         "slice_gather_ceiling_gsectors_per_s": sg,
                         "frac_of_slice_gather_ceiling": (ctr_tot["sectors"] / args.steps / (ctr_tot["ms_sweep"] / args.steps * 1e-3) / 1e9 / sg) if (sg and ctr_tot["ms_sweep"]) else None},
            "counters": {"hits_per_guide": hits / total_guides, "spills": ctr_tot["spills"], "lf_steps": ctr_tot["lf_steps"],
                         "ms_search": ctr_tot["ms_search"] / args.steps, "ms_sweep": ctr_tot["ms_sweep"] / args.steps,
                         "seeds_per_guide": ctr_tot["seeds"] / (per * args.steps), "edited_guides_per_guide": ctr_tot.get("edited_guides", 0) / (per * args.steps), "ms_arrange": ctr_tot["ms_arrange"] / args.steps,
                         "ms_locate": ctr_tot["ms_locate"] / args.steps, "ms_score": ctr_tot["ms_score"] / args.steps,
                         "ms_d2h": ctr_tot["ms_d2h"] / args.steps, "ms_h2d": ctr_tot["ms_h2d"] / args.steps,
                         "ms_host_prepare": ctr_tot["ms_prepare"] / args.steps, "ms_call_wall": ctr_tot["ms_wall"] / args.steps},
            "clocks": clocks,
        }
        if world == 1 and not args.no_file_e2e:
            try:
                line["file_e2e"] = file_e2e(ix, gsx, params, args, kmers, workdir, per * args.steps, hits / total_guides)
            except Exception as e:      # e.g. no room for the text in the work directory: the side leg must not cost the bench line
                line["file_e2e"] = {"value": None, "unit": "guides/s", "error": str(e)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(args, g, chroms, kmers, workdir, ix=ix)
            line["parity_on_cpu_sample"] = parity_on_sample(ix, gsx, cb, args)
            cb.pop("out"); cb.pop("csv")
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    ix.close()
    if dist is not None:
        dist.destroy_process_group()


def file_e2e(ix, gsx, params, args, kmers, workdir, n_max, hits_per_guide):
    """SURVEY 8(d) secondary metric: the same guides through the whole-file driver the CLI uses (gsx_enumerate_file) -- guides
    CSV ingest, enumerate in batches, CSV text of every hit written to a file -- wall clock, index resident."""
    n = int(max(64, min(n_max, 1.2e9 / (hits_per_guide * 90.0 + 60.0))))        # about 1 GB of text at most
    gcsv, out = os.path.join(workdir, "file_e2e.csv"), os.path.join(workdir, "file_e2e.out")
    write_sample_csv(gcsv, kmers, n)
    ix.enumerate_file(gcsv, out, params)                                          # warm-up: arenas, page cache of the output file
    t0 = time.perf_counter()
    _, ctr = ix.enumerate_file(gcsv, out, params)
    dt = time.perf_counter() - t0
    size = os.path.getsize(out)
    os.remove(out)
    t0 = time.perf_counter()
    ix.enumerate_file(gcsv, "/dev/null", params)                                   # the same job with nothing to wait for but itself
    dt0 = time.perf_counter() - t0
    return {"value": n / dt, "unit": "guides/s", "guides": n, "seconds": dt, "output_bytes": size, "output_mb_per_s": size / dt / 1e6,
            "device_ms": ctr["ms_total_device"], "to_dev_null_guides_per_s": n / dt0, "output_dir": workdir,
            "what": "gsx_enumerate_file: guides CSV in, CSV text out (complete mode); two batches in flight on the GPU, a formatter thread on all host "
                    "cores, a writer thread; to_dev_null = the same without the file system"}


def workload_config(args):
    """the workload both arms are quoted on (BASELINE.json configs[2] shape on one GPU unless options say otherwise)"""
    return {"workload": "%.0f Mb uniform-random synthetic genome (seed %d, %d chr; 1-4 mismatch copies planted for the first %d guides, "
                        "the other guides' ~12 hits each are the chance hits of a random genome), "
                        "%d NGG 20-mer guides per GPU per step, mismatches=%d, both strand indexes, locate + CFD + specificity"
                        % (args.genome_mb, args.seed, args.n_chr, args.plant_guides, args.guides_per_step, args.mismatches),
            "genome_mb": args.genome_mb, "guides_per_gpu_per_step": args.guides_per_step, "mismatches": args.mismatches,
            "alt_pams": list(args.alt_pam), "rna_bulges": args.rna_bulges, "dna_bulges": args.dna_bulges,
            "skew": ("repeat families (Zipf 10..10^4 near-copies) for %.1f %% of the guides + %.1f Mb of low-complexity tracts with %.1f %% of the guides drawn from them"
                     % (100 * args.skew_frac, args.skew_lowc_mb, 100 * args.skew_frac)) if args.skew else None}


def write_sample_csv(path, kmers, n):
    with open(path, "w") as f:
        f.write("id,sequence,pam,chromosome,position,sense\n")
        for i in range(n):
            f.write("g%d,%s,NGG,chr1,1,+\n" % (i, kmers[i, :20].tobytes().decode()))


def cpu_baseline(args, g, chroms, kmers, workdir, steps=1, sample=None, ix=None):
    """The reference's CPU implementation of the path on a bounded sample of the workload, all host cores.
    kind "reference": the unmodified reference binary (oracle/_ref/guidescan index + enumerate) -- used when its own
    single-threaded index build fits the time budget (genomes up to --ref-max-mb).
    kind "port": the CPU oracle (oracle/gs_oracle.c, pinned byte-for-byte to the reference) over the FM-index the GPU
    builder exported -- the reference's `guidescan index` needs about an hour for a 3.1 Gb genome."""
    import oracle as O
    import synth
    cores = os.cpu_count() or 1
    n = sample or args.cpu_sample
    gcsv = os.path.join(workdir, "cpu_sample.csv")
    write_sample_csv(gcsv, kmers, n)
    out = os.path.join(workdir, "cpu.out")
    if args.genome_mb <= args.ref_max_mb and O.have_ref():
        tag = "bench_%dmb_s%d" % (int(args.genome_mb), args.seed)
        fa, prefix = os.path.join(workdir, tag + ".fa"), os.path.join(workdir, tag)
        if not os.path.exists(prefix + ".reverse"):
            t0 = time.time()
            synth.write_fasta(fa, g, chroms)
            O.ref_index(fa, prefix, cwd=workdir)
            log("cpu_baseline: reference index built in %.1f s" % (time.time() - t0))
        best = None
        for _ in range(steps):
            t0 = time.time()
            O.ref_enumerate(prefix, gcsv, out, mismatches=args.mismatches, alt_pams=tuple(args.alt_pam), rna_bulges=args.rna_bulges,
                            dna_bulges=args.dna_bulges, threads=cores)
            dt = time.time() - t0
            best = dt if best is None else min(best, dt)
        return {"value": n / best, "unit": "guides/s", "cores": cores, "kind": "reference", "out": out, "csv": gcsv,
                "sample": "first %d guides of the workload, unmodified `guidescan enumerate -n %d`, wall clock incl. index load (%.1f s)" % (n, cores, best)}
    # port: oracle over the exported FM-index
    t0 = time.time()
    own = ix is None
    if own:
        import gsx
        os.environ.update({"GSX_FTAB": "0", "GSX_LOOKAHEAD": "0"})        # only the BWT and the SA samples are exported
        ix = gsx.Index.build_from_text(g, chroms, sa_shift=6, devices=[int(os.environ.get("LOCAL_RANK", 0))])
    b0, b1 = ix.export_bwt(0), ix.export_bwt(1)
    (s0, sh0), (s1, sh1) = ix.export_sa_samples(0), ix.export_sa_samples(1)
    if sh0 > 6 or sh1 != sh0:
        raise RuntimeError("cpu_baseline port needs SA samples at least every 64 rows")
    s0, s1 = np.ascontiguousarray(s0[::1 << (6 - sh0)]), np.ascontiguousarray(s1[::1 << (6 - sh0)])        # the oracle keeps the reference's density (every 64th row)
    oix = O.Index.from_bwt(b0, s0, b1, s1, chroms)
    del b0, b1
    log("cpu_baseline: oracle index imported in %.1f s" % (time.time() - t0))
    best = None
    for _ in range(steps):
        t0 = time.time()
        oix.enumerate_file(O.make_opts(mismatches=args.mismatches, alt_pams=tuple(args.alt_pam), rna_bulges=args.rna_bulges, dna_bulges=args.dna_bulges), gcsv, out, nthreads=cores)
        dt = time.time() - t0
        best = dt if best is None else min(best, dt)
    oix.close()
    if own:
        ix.close()
    return {"value": n / best, "unit": "guides/s", "cores": cores, "kind": "port", "out": out, "csv": gcsv,
            "sample": "first %d guides of the workload, oracle/gs_oracle.c (CPU port pinned to the reference) on %d threads over the "
                      "exported FM-index, wall clock of enumerate (%.1f s)" % (n, cores, best)}


def parity_on_sample(ix, gsx, cb, args):
    """the CPU arm's CSV for the sample must equal the GPU arm's, byte for byte"""
    out = cb["out"] + ".gpu"
    ix.enumerate_file(cb["csv"], out, gsx.make_params(mismatches=args.mismatches, alt_pams=tuple(args.alt_pam), rna_bulges=args.rna_bulges, dna_bulges=args.dna_bulges))
    a, b = open(out, "rb").read(), open(cb["out"], "rb").read()
    if cb["kind"] == "reference":      # the reference interleaves per-guide blocks by thread timing: compare sorted lines
        a, b = b"\n".join(sorted(a.split(b"\n"))), b"\n".join(sorted(b.split(b"\n")))
    return a == b


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import synth
    os.makedirs(args.workdir, exist_ok=True)
    g, chroms, pos, kmers = make_workload(args, 1)
    cb = cpu_baseline(args, g, chroms, kmers, args.workdir, steps=max(1, args.steps), sample=args.cpu_sample)
    cb.pop("out"); cb.pop("csv")
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "guides/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {**workload_config(args), "sample": "the reference arm runs a bounded sample of this workload (cpu_baseline.sample)"},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "guides/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def apply_variant(args):
    if args.variant:
        if args.variant[0] == "g":
            os.environ.update({"GSX_FORCE_GENERAL": "1", "GSX_SEARCH_VARIANT": args.variant[1:]})
        else:
            os.environ.update({"GSX_FORCE_GENERAL": "0", "GSX_FAST_VARIANT": args.variant[1:]})


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gsx", choices=["gsx", "reference"])
    ap.add_argument("--genome-mb", type=float, default=float(os.environ.get("GSX_BENCH_GENOME_MB", 3100)))
    ap.add_argument("--n-chr", type=int, default=24)
    ap.add_argument("--guides-per-step", type=int, default=int(os.environ.get("GSX_BENCH_GUIDES", 200000)))
    ap.add_argument("--ref-max-mb", type=float, default=200.0)
    ap.add_argument("--plant-guides", type=int, default=2000)
    ap.add_argument("--mismatches", type=int, default=3)
    ap.add_argument("--alt-pam", action="append", default=[], help="alternative PAM (repeatable), e.g. --alt-pam NAG (BASELINE configs[4])")
    ap.add_argument("--rna-bulges", type=int, default=0)
    ap.add_argument("--dna-bulges", type=int, default=0)
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--sa-shift", type=int, default=2, help="SA sample density 2^k rows (the reference samples every 64th row; the index keeps every 4th)")
    ap.add_argument("--cpu-sample", type=int, default=8000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-file-e2e", action="store_true")
    ap.add_argument("--e2e-pipeline", type=int, default=1, help="batches in flight in the end-to-end loop (gsx_enumerate_start / _wait); 1 = one call after the other")
    ap.add_argument("--n-runs", type=int, default=0, help="insert this many runs of N (1..50000 bases) into the genome after the guides were sampled and planted")
    ap.add_argument("--no-numa-bind", action="store_true", help="multi-GPU runs: do not bind each rank to its GPU's NUMA node")
    ap.add_argument("--skew", action="store_true", help="SURVEY 8(d) skew stressor: Zipf repeat families for 1 %% of the guides, low-complexity tracts with another 1 %% of the guides drawn from them")
    ap.add_argument("--skew-frac", type=float, default=0.01)
    ap.add_argument("--skew-lowc-mb", type=float, default=2.0)
    ap.add_argument("--sweep-variants", default="")
    ap.add_argument("--variant", default=None, help="f<k> specialised kernel variant k, g<k> general kernel variant k")
    ap.add_argument("--workdir", default=os.environ.get("GSX_BENCH_DIR", "/tmp/gsx_bench"))
    return ap.parse_args(argv)


def main():
    args = parse_args()
    apply_variant(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gsx(args)


if __name__ == "__main__":
    main()
