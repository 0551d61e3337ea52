"""CPU test: the product's parser of the reference's index files (gsx_index.cpp load_sdsl_strand, the host half of gsx_index_open)
against the genome text itself -- full inversion of the BWT it extracts plus every SA sample (tests/ref_index_check.cpp).  The same
tool checks the 3.1 Gb reference index built by tools/ref_3100mb.py (profiles/r02_reference_3100mb.json)."""
import gzip
import json
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT

TOOL = os.path.join(ROOT, "tests", "_build", "ref_index_check")
LIBDIR = os.path.join(ROOT, "guidescan-cli_b200")


@pytest.fixture(scope="module")
def tool():
    lib = os.path.join(LIBDIR, "libgsx.so")
    if not os.path.exists(lib):
        pytest.skip("libgsx.so not built (run __graft_entry__.build())")
    src = os.path.join(ROOT, "tests", "ref_index_check.cpp")
    if not os.path.exists(TOOL) or os.path.getmtime(TOOL) < max(os.path.getmtime(src), os.path.getmtime(lib)):
        os.makedirs(os.path.dirname(TOOL), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I/usr/local/cuda/include", "-o", TOOL, src, "-L" + LIBDIR, "-lgsx",
                               "-Wl,-rpath," + LIBDIR, "-lpthread"])
    return TOOL


def _text_of(fasta_gz, out):
    seq = []
    for line in gzip.open(fasta_gz, "rt"):
        if not line.startswith(">"):
            seq.append(line.strip().upper())
    open(out, "w").write("".join(seq))


@pytest.mark.parametrize("case", ["g200k", "g150kN"])
def test_parser_of_reference_index_files_inverts_to_the_genome(tool, golden_index, tmp_path, case):
    text = os.path.join(tmp_path, case + ".text")
    _text_of(os.path.join(GOLDEN, case + ".fa.gz"), text)
    r = subprocess.run([tool, golden_index[case], text], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    reps = [json.loads(l) for l in r.stdout.splitlines()]
    assert [x["strand"] for x in reps] == ["forward", "reverse"]
    for x in reps:
        assert x["ok"] and x["lf_steps_checked"] == x["genome_length"] == os.path.getsize(text)
        assert x["sa_samples_checked"] == x["genome_length"] // 64 + 1
    if case == "g150kN":
        assert all(x["exception_rows_met"] > 1 for x in reps)          # genome N rows besides the sentinel


def test_a_wrong_text_is_detected(tool, golden_index, tmp_path):
    text = os.path.join(tmp_path, "x.text")
    _text_of(os.path.join(GOLDEN, "g200k.fa.gz"), text)
    s = bytearray(open(text, "rb").read())
    s[12345] = ord("A") if s[12345] != ord("A") else ord("C")
    open(text, "wb").write(bytes(s))
    r = subprocess.run([tool, golden_index["g200k"], text], capture_output=True, text=True)
    assert r.returncode == 1 and '"ok": false' in r.stdout
