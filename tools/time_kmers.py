"""GPU timing of genome-wide guide generation (gsx_generate_kmers) on a synthetic genome: FASTA in, guides CSV out."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "guidescan-cli_b200"))
import gsx  # noqa: E402
import synth  # noqa: E402

mb = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
G = int(mb * 1e6)
d = "/tmp/gsx_kmers"
os.makedirs(d, exist_ok=True)
g = synth.make_genome(G, 2)
chroms = synth.chromosome_table(G, 8)
fa, out = os.path.join(d, "g.fa"), os.path.join(d, "kmers.csv")
synth.write_fasta(fa, g, chroms)
res = {}
for rep in range(2):
    t0 = time.perf_counter()
    n = gsx.generate_kmers(fa, out, pam="NGG", kmer_length=20)
    res["seconds_run%d" % rep] = time.perf_counter() - t0
size = os.path.getsize(out)
dt = res["seconds_run1"]
res.update({"genome_mb": mb, "guides": n, "guides_per_s": n / dt, "genome_mb_per_s": mb / dt, "csv_bytes": size, "csv_mb_per_s": size / dt / 1e6,
            "what": "gsx_generate_kmers: FASTA read + PAM scan on the GPU (both strands, 4 concrete PAMs) + CSV text written by the host"})
print(json.dumps(res))
