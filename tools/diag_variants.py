"""GPU diagnostic: a bulge batch at full genome size through the edited-guide path under several settings, each diffed
against the general kernel's text for the same guides (first differing lines are printed)."""
import argparse
import collections
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "guidescan-cli_b200"))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome-mb", type=float, default=3100.0)
    ap.add_argument("--n", type=int, default=32)
    ap.add_argument("--mismatches", type=int, default=3)
    a = ap.parse_args()
    import gsx
    args = bench.parse_args(["--genome-mb", str(a.genome_mb), "--guides-per-step", str(max(a.n, 64)), "--steps", "1", "--warmup", "0"])
    g, chroms, pos, kmers = bench.make_workload(args, 1)
    ix, how = bench.build_index(gsx, g, chroms, 0, args, args.workdir)
    os.makedirs(args.workdir, exist_ok=True)
    gcsv = os.path.join(args.workdir, "diag.csv")
    bench.write_sample_csv(gcsv, kmers, a.n)
    p = gsx.make_params(mismatches=a.mismatches, rna_bulges=1, dna_bulges=1)
    outs = {}
    configs = [("general", {"GSX_VARIANTS": "0"}), ("edited", {}), ("edited_nosweep", {"GSX_SWEEP": "0"}), ("edited_order1", {"GSX_ORDER": "1"}),
               ("edited_chunk20k", {"GSX_VARIANT_CHUNK": "20000"}), ("general_order2", {"GSX_VARIANTS": "0", "GSX_ORDER": "2"})]
    for tag, env in configs:
        for k, v in env.items():
            os.environ[k] = v
        t0 = time.time()
        out = os.path.join(args.workdir, "diag_%s.csv" % tag)
        _, ctr = ix.enumerate_file(gcsv, out, p)
        for k in env:
            os.environ.pop(k)
        outs[tag] = open(out, "rb").read().split(b"\n")
        same = outs[tag] == outs["general"]
        print("%-16s lines %8d  edited %8d  matches %9d  %.2f s  %s" % (tag, len(outs[tag]), ctr["edited_guides"], ctr["matches"], time.time() - t0,
                                                                   "== general" if same else "DIFFERS"), flush=True)
        if not same:
            A, B = collections.Counter(outs["general"]), collections.Counter(outs[tag])
            only_a, only_b = list((A - B).elements()), list((B - A).elements())
            print("   lines only in general: %d, only in %s: %d" % (len(only_a), tag, len(only_b)))
            for l in only_a[:6]:
                print("   - " + l.decode()[:200])
            for l in only_b[:6]:
                print("   + " + l.decode()[:200])
            if not only_a and not only_b:
                for i, (x, y) in enumerate(zip(outs["general"], outs[tag])):
                    if x != y:
                        print("   first order difference at line %d:\n   - %s\n   + %s" % (i, x.decode()[:200], y.decode()[:200]))
                        break
    ix.close()


if __name__ == "__main__":
    main()
