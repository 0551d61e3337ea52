import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "guidescan-cli_b200"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def golden_manifest():
    return json.load(open(os.path.join(GOLDEN, "manifest.json")))


def golden_cases():
    m = golden_manifest()
    return [(c, v) for c, e in sorted(m["cases"].items()) for v in sorted(e["variants"])]


def golden_output(case, variant) -> bytes:
    return gzip.open(os.path.join(GOLDEN, "%s.%s.out.gz" % (case, variant)), "rb").read()


@pytest.fixture(scope="session")
def golden_dir(tmp_path_factory):
    """Unpacked golden FASTAs + guides in a scratch directory: {case: (fasta, guides_csv)}."""
    d = tmp_path_factory.mktemp("golden")
    out = {}
    for case in golden_manifest()["cases"]:
        fa = os.path.join(d, case + ".fa")
        with open(fa, "wb") as f:
            f.write(gzip.open(os.path.join(GOLDEN, case + ".fa.gz"), "rb").read())
        out[case] = (fa, os.path.join(GOLDEN, case + ".guides.csv"))
    return out


def canonical_blocks(text: str, fmt: str):
    """Per-guide row blocks (order inside a block kept, blocks keyed by id) -- the reference interleaves
    blocks by thread timing (manual.tex:203-206) but each block is written atomically."""
    blocks = {}
    header = []
    for line in text.splitlines():
        if fmt == "sam" and line.startswith("@"):
            header.append(line)
            continue
        if fmt == "csv" and not header:
            header.append(line)
            continue
        key = line.split("\t" if fmt == "sam" else ",", 1)[0]
        blocks.setdefault(key, []).append(line)
    return header, blocks


@pytest.fixture(scope="session")
def golden_index(golden_dir, tmp_path_factory):
    """{case: index prefix} built by the unmodified reference `guidescan index` (oracle/_ref)."""
    import oracle as O
    if not O.have_ref():
        pytest.skip("oracle/_ref/guidescan not built (needs /root/reference in the build container)")
    out = {}
    for case, (fa, _) in golden_dir.items():
        d = tmp_path_factory.mktemp("idx_" + case)
        prefix = os.path.join(d, case)
        O.ref_index(fa, prefix, cwd=str(d))
        out[case] = prefix
    return out


def variant_cli_args(kw):
    """golden variant options -> argument list shared by tests/host_core_check and bin/guidescan-style tools"""
    a = ["-m", str(kw.get("mismatches", 3))]
    if kw.get("rna_bulges"):
        a += ["--rna", str(kw["rna_bulges"])]
    if kw.get("dna_bulges"):
        a += ["--dna", str(kw["dna_bulges"])]
    if kw.get("threshold") is not None:
        a += ["-t", str(kw["threshold"])]
    if kw.get("start"):
        a += ["--start"]
    if kw.get("max_off_targets") is not None:
        a += ["--max", str(kw["max_off_targets"])]
    if kw.get("fmt") == "sam":
        a += ["--sam"]
    if kw.get("mode") == "succinct":
        a += ["--succinct"]
    for p in kw.get("alt_pams", ()):
        a += ["-a", p]
    return a
