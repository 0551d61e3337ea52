#!/usr/bin/env python
"""Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference binary
(oracle/_ref/guidescan, built by oracle/Makefile.ref from /root/reference) on small seeded inputs.

Run in the build container (needs /root/reference once, to build oracle/_ref):
    make -C oracle ref && python tests/golden/make_golden.py
Outputs (committed): <case>.fa.gz, <case>.guides.csv, <case>.<variant>.out.gz, manifest.json.
The reference ships no tests of this path, so these outputs are the pin.
"""
import gzip, hashlib, json, os, shutil, subprocess, sys, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "guidescan-cli_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import synth  # noqa: E402
import oracle as O  # noqa: E402

VARIANTS = {
    "m3_csv": dict(mismatches=3),
    "m3_csv_succinct": dict(mismatches=3, mode="succinct"),
    "m2_sam": dict(mismatches=2, fmt="sam"),
    "m3_sam_succinct": dict(mismatches=3, fmt="sam", mode="succinct"),
    "m0_csv": dict(mismatches=0),
    "m4_max2_csv": dict(mismatches=4, max_off_targets=2),
    "m4_max1_sam": dict(mismatches=4, max_off_targets=1, fmt="sam"),
    "m3_altNAG_sam": dict(mismatches=3, alt_pams=("NAG",), fmt="sam"),
    "m3_altNAG_NGA_csv": dict(mismatches=3, alt_pams=("NAG", "NGA")),
    "m2_start_csv": dict(mismatches=2, start=True),
    "m3_thr1_csv": dict(mismatches=3, threshold=1),
    "m3_thr2_csv": dict(mismatches=3, threshold=2),
    "m1_r1_d1_csv": dict(mismatches=1, rna_bulges=1, dna_bulges=1),
    "m2_r1_csv": dict(mismatches=2, rna_bulges=1),
    "m2_d1_csv": dict(mismatches=2, dna_bulges=1),
    "m1_r1_d1_sam_succinct": dict(mismatches=1, rna_bulges=1, dna_bulges=1, fmt="sam", mode="succinct"),
    "m0_r2_d2_csv": dict(mismatches=0, rna_bulges=2, dna_bulges=2),
    "m1_r1_d1_altNAG_max3_csv": dict(mismatches=1, rna_bulges=1, dna_bulges=1, alt_pams=("NAG",), max_off_targets=3),
    "m0_r1_d1_csv": dict(mismatches=0, rna_bulges=1, dna_bulges=1),
}
CASES = {
    # name: (G, n_chr, n_guides, seed, with_N)
    "g200k": (200_000, 4, 40, 11, False),
    "g150kN": (150_000, 3, 24, 12, True),
}


def build_case(name, G, n_chr, n_guides, seed, with_N, out):
    rng = np.random.default_rng(seed + 77)
    g = synth.make_genome(G, seed)
    pos, kmers = synth.sample_guides(g, n_guides, seed)
    synth.plant(g, kmers, seed)
    chroms = synth.chromosome_table(G, n_chr)
    starts = np.cumsum([0] + [c[1] for c in chroms])
    # edge plants: '-' strand hit at genome offset 0, a site straddling a
    # chromosome boundary, a '+' site starting one base before a chromosome start
    g[0:23] = synth.revcomp_bytes(kmers[2])
    b = int(starts[1]); g[b - 10:b + 13] = kmers[3]
    b = int(starts[2]); g[b - 1:b + 22] = kmers[4]
    g[G - 23:G] = kmers[5]
    g[G - 40:G - 17] = synth.revcomp_bytes(kmers[6])
    # a tandem repeat family so that one SA interval holds several rows
    for r in range(6):
        at = int(rng.integers(2000, G - 2000)); g[at:at + 23] = kmers[7]
    # near-copies carrying alternative PAMs (xAG / xGA) so that -a NAG / NGA finds something
    for i in range(16, min(len(kmers), 24)):
        c = kmers[i].copy()
        for w in rng.choice(20, i % 3, replace=False):
            c[w] = [b for b in b"ACGT" if b != c[w]][int(rng.integers(0, 3))]
        c[20:23] = np.frombuffer((b"AAG", b"TGA", b"CAG", b"GGA")[i % 4], dtype=np.uint8)
        if i % 2:
            c = synth.revcomp_bytes(c)
        at = int(rng.integers(2000, G - 2000)); g[at:at + 23] = c
    if with_N:
        for _ in range(12):
            at = int(rng.integers(1000, G - 2000)); g[at:at + int(rng.integers(1, 40))] = ord("N")
        at = int(rng.integers(1000, G - 2000)); c = kmers[0].copy(); c[20] = ord("N"); g[at:at + 23] = c
        at = int(rng.integers(1000, G - 2000)); c = kmers[1].copy(); c[7] = ord("N"); g[at:at + 23] = c
        at = int(rng.integers(1000, G - 2000)); c = synth.revcomp_bytes(kmers[8]); c[2] = ord("N"); g[at:at + 23] = c
    fa = os.path.join(out, name + ".fa")
    synth.write_fasta(fa, g, chroms)
    gcsv = os.path.join(out, name + ".guides.csv")
    synth.write_guides_csv(gcsv, pos, kmers, chroms)
    # extra hand-made guides: '-' sense, empty PAM, N inside the protospacer, a 19-mer and a 21-mer, a non-NGG PAM,
    # a guide that occurs nowhere, a low-complexity guide
    with open(gcsv, "a") as f:
        k9 = kmers[9][:20].tobytes().decode()
        f.write("neg_sense,%s,NGG,chr1,100,-\n" % kmers[10][:20].tobytes().decode())
        f.write("no_pam,%s,,chr1,100,+\n" % kmers[11][:20].tobytes().decode())
        f.write("n_in_guide,%s,NGG,chr1,100,+\n" % (k9[:5] + "N" + k9[6:]))
        f.write("short19,%s,NGG,chr1,100,+\n" % kmers[12][1:20].tobytes().decode())
        f.write("long21,%s,NGG,chr1,100,+\n" % g[int(pos[13]) - 1:int(pos[13]) + 20].tobytes().decode())
        f.write("pam_nag,%s,NAG,chr1,100,+\n" % kmers[14][:20].tobytes().decode())
        f.write("pam_agg,%s,AGG,chr1,100,+\n" % kmers[15][:20].tobytes().decode())
        f.write("nowhere,ACGTACGTACGTACGTACGT,NGG,chr1,100,+\n")
        f.write("polyA,AAAAAAAAAAAAAAAAAAAA,NGG,chr1,100,+\n")
    return fa, gcsv


def main():
    if not O.have_ref():
        sys.exit("oracle/_ref/guidescan missing: run `make -C oracle ref` first")
    manifest = {"reference": "pritykinlab/guidescan-cli 2.0.0 (unmodified, oracle/Makefile.ref)", "cases": {}}
    for name, (G, n_chr, n_guides, seed, with_N) in CASES.items():
        tmp = tempfile.mkdtemp(prefix="gsgold_")
        fa, gcsv = build_case(name, G, n_chr, n_guides, seed, with_N, tmp)
        prefix = os.path.join(tmp, name)
        O.ref_index(fa, prefix, cwd=tmp)
        entry = {"G": G, "n_chr": n_chr, "seed": seed, "variants": {}}
        for vname, kw in VARIANTS.items():
            outp = os.path.join(tmp, vname + ".out")
            O.ref_enumerate(prefix, gcsv, outp, threads=1, **kw)
            data = open(outp, "rb").read()
            with gzip.GzipFile(os.path.join(HERE, "%s.%s.out.gz" % (name, vname)), "wb", mtime=0) as f:
                f.write(data)
            entry["variants"][vname] = {"opts": {k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()},
                                        "sha256": hashlib.sha256(data).hexdigest(), "bytes": len(data),
                                        "lines": data.count(b"\n")}
        with gzip.GzipFile(os.path.join(HERE, name + ".fa.gz"), "wb", mtime=0) as f:
            f.write(open(fa, "rb").read())
        shutil.copy(gcsv, os.path.join(HERE, name + ".guides.csv"))
        entry["fasta_sha256"] = hashlib.sha256(open(fa, "rb").read()).hexdigest()
        manifest["cases"][name] = entry
        shutil.rmtree(tmp)
    json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)
    print("golden written:", {k: len(v["variants"]) for k, v in manifest["cases"].items()})


if __name__ == "__main__":
    main()
