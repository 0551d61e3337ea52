#!/usr/bin/env python
"""Golden vectors for genome-wide guide generation: runs the UNMODIFIED reference script
/root/reference/scripts/generate_kmers.py (with oracle/bio_stub standing in for Biopython's FASTA reader) on a small FASTA
with the awkward cases in it, and stores its stdout.  Run in the build container:
    python tests/golden/make_kmers_golden.py
Outputs (committed): kmers.fa.gz, kmers.<variant>.csv.gz, kmers_manifest.json."""
import gzip, json, os, subprocess, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SCRIPT = "/root/reference/scripts/generate_kmers.py"

VARIANTS = {
    "default": [],
    "nag": ["--pam", "NAG"],
    "k18_prefix": ["--kmer-length", "18", "--prefix", "lib_"],
    "minlen": ["--min-chr-length", "1000"],
    "cpf1_start": ["--pam", "TTTN", "--start", "--kmer-length", "23"],
    "two_n": ["--pam", "NGN", "--kmer-length", "19"],
    "gg_start": ["--pam", "GG", "--start"],
}


def fasta_text():
    rng = np.random.default_rng(2024)
    def rnd(n): return "".join("ACGT"[i] for i in rng.integers(0, 4, n))
    recs = []
    recs.append(("chrA the first one", rnd(30_000)))
    s = list(rnd(20_000))
    for a, b in ((500, 560), (7000, 7003), (15000, 15400)):
        s[a:b] = "N" * (b - a)
    for a, b in ((1000, 1800), (9000, 9050)):                       # soft-masked stretches
        s[a:b] = "".join(s[a:b]).lower()
    s[3000:3005] = "RYKMS"                                          # IUPAC codes inside candidate k-mers
    recs.append(("chrB", "".join(s)))
    recs.append(("tiny", "ACGTTGGCCAGG"))                             # shorter than a k-mer
    recs.append(("edge", "CCAACC" + rnd(40) + "TGGAGG"))              # PAMs at both chromosome ends
    recs.append(("chrC\tdescr", "GG" * 40 + rnd(500) + "CC" * 40))    # overlapping PAM occurrences
    out = []
    for name, seq in recs:
        out.append(">" + name + "\n")
        for i in range(0, len(seq), 60):
            out.append(seq[i:i + 60] + "\n")
        out.append("\n")                                              # blank line between records
    return "".join(out)


def main():
    fa = os.path.join(HERE, "kmers.fa")
    open(fa, "w").write(fasta_text())
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "oracle", "bio_stub"))
    manifest = {}
    for name, extra in VARIANTS.items():
        out = subprocess.run([sys.executable, SCRIPT, fa] + extra, env=env, capture_output=True, check=True).stdout
        with gzip.open(os.path.join(HERE, "kmers.%s.csv.gz" % name), "wb", compresslevel=9) as f:
            f.write(out)
        manifest[name] = {"args": extra, "rows": out.count(b"\n") - 1}
        print(name, manifest[name])
    with gzip.open(fa + ".gz", "wb", compresslevel=9) as f:
        f.write(open(fa, "rb").read())
    os.remove(fa)
    json.dump(manifest, open(os.path.join(HERE, "kmers_manifest.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
