"""Genome-wide guide generation.  Golden vectors: stdout of the UNMODIFIED reference script
scripts/generate_kmers.py (tests/golden/make_kmers_golden.py).  CPU: the oracle restatement against them; GPU:
gsx_generate_kmers (PAM scan kernel + ordered compaction) against them and against the oracle on a fresh input."""
import gzip
import json
import os
import sys

import pytest

from conftest import ROOT

GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def variants():
    return json.load(open(os.path.join(GOLD, "kmers_manifest.json")))


def opts_of(args):
    kw, it = {}, iter(args)
    for a in it:
        if a == "--pam":
            kw["pam"] = next(it)
        elif a == "--kmer-length":
            kw["kmer_length"] = int(next(it))
        elif a == "--min-chr-length":
            kw["min_chr_length"] = int(next(it))
        elif a == "--prefix":
            kw["prefix"] = next(it)
        elif a == "--start":
            kw["start"] = True
    return kw


@pytest.fixture(scope="module")
def gsx():
    import gsx as g
    if g.device_count() < 1:
        pytest.fail("no CUDA device: the product has no CPU path")
    return g


@pytest.fixture(scope="module")
def fasta(tmp_path_factory):
    p = str(tmp_path_factory.mktemp("kmers") / "kmers.fa")
    open(p, "wb").write(gzip.open(os.path.join(GOLD, "kmers.fa.gz"), "rb").read())
    return p


@pytest.mark.parametrize("name", sorted(variants()))
def test_oracle_restatement_matches_reference_script(fasta, name):
    import kmers_oracle as K
    kw = opts_of(variants()[name]["args"])
    got = K.generate(fasta, pam=kw.get("pam", "NGG"), k=kw.get("kmer_length", 20), min_chr_length=kw.get("min_chr_length", 0),
                     prefix=kw.get("prefix", ""), start=kw.get("start", False))
    want = gzip.open(os.path.join(GOLD, "kmers.%s.csv.gz" % name), "rb").read().decode()
    assert got == want
    assert got.count("\n") - 1 == variants()[name]["rows"]


@pytest.mark.gpu
@pytest.mark.parametrize("rows_per_thread", [0, 3])
@pytest.mark.parametrize("name", sorted(variants()))
def test_gpu_kmer_generation_matches_reference_script(gsx, fasta, tmp_path, monkeypatch, name, rows_per_thread):
    if rows_per_thread:        # the threaded row writer (chromosomes with more than 65536 sites per PAM), reached on the small goldens
        monkeypatch.setenv("GSX_KMERS_ROWS_PER_THREAD", str(rows_per_thread))
    kw = opts_of(variants()[name]["args"])
    out = str(tmp_path / "k.csv")
    n = gsx.generate_kmers(fasta, out, **kw)
    want = gzip.open(os.path.join(GOLD, "kmers.%s.csv.gz" % name), "rb").read()
    assert open(out, "rb").read() == want
    assert n == variants()[name]["rows"]


@pytest.mark.gpu
def test_gpu_kmer_generation_feeds_enumerate(gsx, tmp_path):
    """fresh seeded genome: generator == oracle restatement, and its CSV is accepted by enumerate_file as is"""
    import kmers_oracle as K
    import oracle as O
    import synth
    d = str(tmp_path)
    synth.make_dataset(d, 300_000, 3, 10, seed=77, name="km")
    fa = os.path.join(d, "km.fa")
    out = os.path.join(d, "km.kmers.csv")
    n = gsx.generate_kmers(fa, out, pam="NGG", kmer_length=20, min_chr_length=0, prefix="x_")
    assert open(out).read() == K.generate(fa, prefix="x_")
    assert n > 30_000
    # the first 300 generated guides through the search, against the oracle
    head = os.path.join(d, "head.csv")
    open(head, "w").writelines(open(out).readlines()[:301])
    ix = gsx.Index.build(fa, devices=[0])
    got = os.path.join(d, "g.out")
    ix.enumerate_file(head, got, gsx.make_params(mismatches=2))
    want = os.path.join(d, "o.out")
    O.Index(fa).enumerate_file(O.make_opts(mismatches=2), head, want, nthreads=4)
    assert open(got, "rb").read() == open(want, "rb").read()
    ix.close()
