"""Synthetic genomes + guide sets for the off-target enumeration hot path.

Recipe = SURVEY.md Appendix D / §8(d) (restated, not copied): an iid uniform ACGT genome, guides sampled
from + strand sites that carry an NGG PAM, and for every guide one planted copy at each substitution
distance d = 1..4 (random PAM first base, 50 % reverse-complemented) so that hits exist at every distance.
The FASTA (60 columns, equal-length chromosomes chr1..chrK) is what `guidescan index` consumes
(reference src/genomics/seq_io.cxx:57-122); the guides CSV has the six columns the reference's reader
requires (reference src/genomics/kmer.cxx:9-25).
"""
from __future__ import annotations

import os
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTNacgtn", b"TGCANtgcan"):
    _COMP[_a] = _b


def revcomp_bytes(a: np.ndarray) -> np.ndarray:
    return _COMP[a[::-1]]


def make_genome(G: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    out = np.empty(G, dtype=np.uint8)
    step = 1 << 26
    for s in range(0, G, step):
        e = min(G, s + step)
        out[s:e] = _ACGT[rng.integers(0, 4, e - s, dtype=np.uint8)]
    return out


def sample_guides(g: np.ndarray, n_guides: int, seed: int, margin: int = 1000):
    """Distinct + strand positions p with g[p+21:p+23] == 'GG'; returns (positions, 23-mers incl. PAM)."""
    rng = np.random.default_rng(seed + 1000003)
    G = len(g)
    pos = np.empty(0, dtype=np.int64)
    need = n_guides
    # rejection sampling keeps this O(n_guides) even for a 3.1 Gb genome
    while len(pos) < n_guides:
        cand = rng.integers(margin, G - margin - 23, int(need * 20) + 64)
        ok = (g[cand + 21] == ord("G")) & (g[cand + 22] == ord("G"))
        pos = np.unique(np.concatenate([pos, cand[ok]]))
        need = n_guides - len(pos)
    pos = np.sort(rng.choice(pos, n_guides, replace=False))
    kmers = g[pos[:, None] + np.arange(23)[None, :]].copy()
    return pos, kmers


def plant(g: np.ndarray, kmers: np.ndarray, seed: int, dists=(1, 2, 3, 4), margin: int = 1000):
    """For each guide and each d: a copy with exactly d substitutions in the protospacer, in place."""
    rng = np.random.default_rng(seed + 2000003)
    G = len(g)
    placed = []
    for i in range(len(kmers)):
        for d in dists:
            c = kmers[i].copy()
            where = rng.choice(20, d, replace=False)
            for w in where:
                alt = [b for b in b"ACGT" if b != c[w]]
                c[w] = alt[rng.integers(0, 3)]
            c[20] = _ACGT[rng.integers(0, 4)]
            rc = bool(rng.random() < 0.5)
            if rc:
                c = revcomp_bytes(c)
            at = int(rng.integers(margin, G - margin - 23))
            g[at:at + 23] = c
            placed.append((i, d, at, rc))
    return placed


def chromosome_table(G: int, n_chr: int):
    base = G // n_chr
    lens = [base] * n_chr
    lens[-1] += G - base * n_chr
    return [("chr%d" % (i + 1), lens[i]) for i in range(n_chr)]


def write_fasta(path: str, g: np.ndarray, chroms) -> None:
    off = 0
    with open(path, "wb") as f:
        for name, ln in chroms:
            f.write(b">" + name.encode() + b"\n")
            seq = g[off:off + ln]
            full = (ln // 60) * 60
            if full:
                body = np.empty((full // 60, 61), dtype=np.uint8)
                body[:, :60] = seq[:full].reshape(-1, 60)
                body[:, 60] = 10
                f.write(body.tobytes())
            if ln > full:
                f.write(seq[full:].tobytes() + b"\n")
            off += ln


def write_guides_csv(path: str, pos: np.ndarray, kmers: np.ndarray, chroms, pam: str = "NGG",
                     ids=None) -> None:
    starts = np.cumsum([0] + [c[1] for c in chroms])
    with open(path, "w") as f:
        f.write("id,sequence,pam,chromosome,position,sense\n")
        for i in range(len(pos)):
            ci = int(np.searchsorted(starts, pos[i], side="right") - 1)
            gid = ids[i] if ids is not None else "g%d" % i
            f.write("%s,%s,%s,%s,%d,+\n" % (gid, kmers[i, :20].tobytes().decode(), pam,
                                           chroms[ci][0], pos[i] - starts[ci] + 1))


def make_dataset(outdir: str, G: int, n_chr: int, n_guides: int, seed: int, name: str = "synth",
                 n_runs_of_N: int = 0, plant_dists=(1, 2, 3, 4)):
    """Writes <outdir>/<name>.fa and <outdir>/<name>.guides.csv; returns (genome bytes, chroms, pos, kmers)."""
    os.makedirs(outdir, exist_ok=True)
    g = make_genome(G, seed)
    pos, kmers = sample_guides(g, n_guides, seed)
    plant(g, kmers, seed, dists=plant_dists)
    if n_runs_of_N:
        rng = np.random.default_rng(seed + 3000003)
        for _ in range(n_runs_of_N):
            at = int(rng.integers(1000, G - 2000))
            g[at:at + int(rng.integers(1, 40))] = ord("N")
        # a literal genome N in PAM position 0 of a planted exact copy
        at = int(rng.integers(1000, G - 2000))
        c = kmers[0].copy(); c[20] = ord("N"); g[at:at + 23] = c
        at = int(rng.integers(1000, G - 2000))
        c = kmers[min(1, len(kmers) - 1)].copy(); c[7] = ord("N"); g[at:at + 23] = c
    chroms = chromosome_table(G, n_chr)
    write_fasta(os.path.join(outdir, name + ".fa"), g, chroms)
    write_guides_csv(os.path.join(outdir, name + ".guides.csv"), pos, kmers, chroms)
    return g, chroms, pos, kmers


def add_skew(g: np.ndarray, kmers: np.ndarray, seed: int, frac: float = 0.01, lowc_mb: float = 2.0, max_copies: int = 10000,
             margin: int = 1000):
    """SURVEY.md 8(d) skew stressor, in place (the plain recipe has no inter-guide skew at all: CV 0.2 %).
    (1) Repeat families: `frac` of the guides get a Zipf-distributed number (10 .. max_copies) of extra near-copies of their site --
        0-3 substitutions in the protospacer, random PAM first base, half of them reverse-complemented -- at random positions.
    (2) Low complexity: `lowc_mb` Mb of tracts (10 kb each: homopolymer, di- and trinucleotide repeats that contain GG, 1-5 % noise),
        and another `frac` of the guides REPLACED by 23-mers drawn from NGG sites inside the tracts.
    Returns (kmers with the replaced rows, indices of the family guides, indices of the low-complexity guides)."""
    rng = np.random.default_rng(seed + 4000003)
    G, n = len(g), len(kmers)
    k = max(1, int(n * frac))
    fam = rng.choice(n, k, replace=False)
    # Zipf(1.3) scaled to 10 .. max_copies: a few families of thousands, most of tens
    copies = np.minimum(max_copies, 10 * rng.zipf(1.3, k)).astype(np.int64)
    code = np.zeros(256, dtype=np.uint8)
    code[_ACGT] = np.arange(4, dtype=np.uint8)
    cols = np.arange(23)[None, :]
    for gi, nc in zip(fam, copies):
        nc = int(nc)
        c = np.tile(kmers[gi], (nc, 1))
        nsub = rng.integers(0, 4, nc)
        where = rng.integers(0, 20, (nc, 3))
        shift = rng.integers(1, 4, (nc, 3))
        rows = np.arange(nc)
        for j in range(3):                                   # (positions may coincide: "up to" three substitutions)
            sel = rows[nsub > j]
            w = where[sel, j]
            c[sel, w] = _ACGT[(code[c[sel, w]] + shift[sel, j]) % 4]
        c[:, 20] = _ACGT[rng.integers(0, 4, nc)]
        rc = rng.random(nc) < 0.5
        c[rc] = _COMP[c[rc][:, ::-1]]
        at = rng.integers(margin, G - margin - 23, nc)
        g[at[:, None] + cols] = c
    units = [b"G", b"AG", b"CGG", b"AGG", b"TGG", b"GGA", b"GGC", b"GGT", b"AGGG", b"GGAA", b"CCGG", b"GAGG", b"GGTA", b"TCGG", b"AAGG", b"GGCA"]
    tract = 10000
    n_tracts = max(1, int(lowc_mb * 1e6) // tract)
    starts = np.sort(rng.choice((G - 2 * margin) // tract - 1, n_tracts, replace=False)) * tract + margin
    sites = []
    for ti, s in enumerate(starts):
        u = np.frombuffer(units[ti % len(units)], dtype=np.uint8)
        body = np.tile(u, tract // len(u) + 1)[:tract].copy()
        noise = rng.random(tract) < rng.uniform(0.01, 0.05)
        body[noise] = _ACGT[rng.integers(0, 4, int(noise.sum()))]
        g[s:s + tract] = body
        p = np.arange(s, s + tract - 23)
        ok = (g[p + 21] == ord("G")) & (g[p + 22] == ord("G"))
        sites.append(p[ok])
    sites = np.concatenate(sites)
    rest = np.setdiff1d(np.arange(n), fam)
    low = rng.choice(rest, min(k, len(rest)), replace=False)
    at = rng.choice(sites, len(low), replace=len(sites) < len(low))
    kmers = kmers.copy()
    kmers[low] = g[at[:, None] + np.arange(23)[None, :]]
    return kmers, fam, low
