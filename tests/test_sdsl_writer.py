"""The index written in the reference's own on-disk format (guidescan-cli_b200/csrc/gsx_sdsl_write.cpp, gsx_index_save_reference_format).

CPU tests (no GPU needed; they need oracle/_ref, i.e. the build container or a box the snapshot travelled to):
  * index files written by the unmodified `guidescan index` are parsed by the product (load_sdsl_strand) and written back:
    the files must be IDENTICAL -- bit vector, rank and select directories, tree, SA and ISA samples, alphabet;
  * the rank / select directories and the wavelet tree against sdsl's own (oracle/_ref/sdsl_probe, built from the reference's
    vendored sdsl) on inputs a whole index run cannot be steered to: the corner cases of select_support_mcl (zero counts around
    the 4033rd argument of the last superblock with padding bits behind, sparse vectors with long blocks, the switch of
    construction at 100 000 bits) and byte alphabets with ties, deep trees and all 256 bytes.
GPU tests: an index built by the GPU builder and saved in this format is byte for byte the file the reference builds from the same
FASTA, and the unmodified reference enumerates over the files `bin/guidescan index --reference-format` writes.
The 3.1 Gb reference index goes through the same tool (profiles/r02q_reference_format_3100mb.json)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden_output

TOOL = os.path.join(ROOT, "tests", "_build", "sdsl_write_check")
PROBE = os.path.join(ROOT, "oracle", "_ref", "sdsl_probe")
LIBDIR = os.path.join(ROOT, "guidescan-cli_b200")


@pytest.fixture(scope="module")
def tool():
    lib = os.path.join(LIBDIR, "libgsx.so")
    if not os.path.exists(lib):
        pytest.skip("libgsx.so not built (run __graft_entry__.build())")
    src = os.path.join(ROOT, "tests", "sdsl_write_check.cpp")
    if not os.path.exists(TOOL) or os.path.getmtime(TOOL) < max(os.path.getmtime(src), os.path.getmtime(lib)):
        os.makedirs(os.path.dirname(TOOL), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I/usr/local/cuda/include", "-o", TOOL, src, "-L" + LIBDIR, "-lgsx",
                               "-Wl,-rpath," + LIBDIR, "-lpthread"])
    return TOOL


@pytest.fixture(scope="module")
def probe():
    if not os.path.exists(PROBE):
        pytest.skip("oracle/_ref/sdsl_probe not built (needs /root/reference in the build container)")
    return PROBE


def _same(a, b):
    return open(a, "rb").read() == open(b, "rb").read()


@pytest.mark.parametrize("case", ["g200k", "g150kN"])
@pytest.mark.parametrize("isa_from_samples", ["1", "0"])
def test_reference_index_files_survive_parse_and_rewrite_unchanged(tool, golden_index, tmp_path, case, isa_from_samples):
    for strand in ("forward", "reverse"):
        src = golden_index[case] + "." + strand
        out = os.path.join(tmp_path, strand)
        subprocess.check_call([tool, "rewrite", src, out, "3"], env={**os.environ, "GSX_ISA_FROM_SAMPLES": isa_from_samples},
                              stdout=subprocess.DEVNULL)
        assert _same(src, out), (case, strand)


def test_truncated_or_scrambled_reference_files_are_refused(tool, golden_index, tmp_path):
    src = golden_index["g200k"] + ".forward"
    data = open(src, "rb").read()
    bad = os.path.join(tmp_path, "bad")
    for cut in (8, 100, 4096, int(len(data) * 0.3), int(len(data) * 0.6), int(len(data) * 0.8)):
        open(bad, "wb").write(data[:cut])
        r = subprocess.run([tool, "rewrite", bad, os.path.join(tmp_path, "out")], capture_output=True, text=True)
        assert r.returncode == 1 and ("malformed index file" in r.stderr or "too few SA samples" in r.stderr), (cut, r.stderr)
    # a size field that asks for more than the file holds
    broken = bytearray(data)
    broken[16:24] = (1 << 50).to_bytes(8, "little")
    open(bad, "wb").write(bytes(broken))
    r = subprocess.run([tool, "rewrite", bad, os.path.join(tmp_path, "out")], capture_output=True, text=True)
    assert r.returncode == 1 and "malformed index file" in r.stderr


def _bit_vector_cases():
    rng = np.random.default_rng(11)
    cases = []
    for n, dens in [(64, 1.0), (4097, 0.0), (4096, 0.5), (70000, 0.5), (99999, 0.5), (100000, 0.5), (100001, 0.5), (131072, 0.5),
                    (262145, 0.3), (300000, 0.01), (300000, 0.99), (500000, 0.001), (524288, 0.5)]:
        cases.append(("random n=%d p=%g" % (n, dens), (rng.random(n) < dens).astype(np.uint8)))
    # zeros of the last superblock around 4033 with up to 63 padding bits behind the vector: sdsl counts the padding as zeros
    # when it works on words
    for r in (3965, 3970, 4000, 4031, 4032, 4033, 4034, 4060, 4095, 0, 1, 30):
        n = 64 * int(rng.integers(1600, 4000)) + int(rng.integers(1, 8))
        bits = (rng.random(n) < 0.5).astype(np.uint8)
        z = int((bits == 0).sum())
        ones = np.flatnonzero(bits == 1)
        bits[rng.choice(ones, (r - z) % 4096, replace=False)] = 0
        assert int((bits == 0).sum()) % 4096 == r
        cases.append(("zeros mod 4096 = %d, padding %d" % (r, (-n) % 64), bits))
    for n, k in [(421730, 850), (1300000, 13821), (2600000, 9000)]:
        bits = np.zeros(n, dtype=np.uint8)
        bits[rng.choice(n, k, replace=False)] = 1
        cases.append(("sparse n=%d k=%d" % (n, k), bits))
        cases.append(("dense n=%d k=%d" % (n, k), 1 - bits))
    return cases


def test_rank_and_select_directories_equal_sdsls(tool, probe, tmp_path):
    words, ref, mine = (os.path.join(tmp_path, x) for x in ("words", "ref", "mine"))
    for name, bits in _bit_vector_cases():
        n = len(bits)
        padded = np.zeros((n + 63) // 64 * 64, dtype=np.uint8)
        padded[:n] = bits
        np.packbits(padded, bitorder="little").tofile(words)
        subprocess.check_call([probe, "bv", words, str(n), ref])
        subprocess.check_call([tool, "bv", words, str(n), mine])
        assert _same(ref, mine), name


def _bwt_cases():
    rng = np.random.default_rng(12)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    cases = []
    for n in (2, 5, 63, 64, 65, 1000, 44000, 50000, 200000):
        b = acgt[rng.integers(0, 4, n)].copy()
        b[rng.integers(0, n)] = 0
        cases.append(("ACGT$ n=%d" % n, b))
        b = b.copy()
        b[rng.choice(n, max(1, n // 50), replace=False)] = ord("N")
        b[rng.integers(0, n)] = 0
        cases.append(("ACGTN$ n=%d" % n, b))
    cases.append(("equal counts", np.tile(np.frombuffer(b"\x00ACGT", dtype=np.uint8), 5000)))
    b = np.tile(np.frombuffer(b"ACGTNRYK", dtype=np.uint8), 3000)
    b[7] = 0
    cases.append(("IUPAC, equal counts", b))
    b = acgt[rng.choice(4, 300000, p=[0.7, 0.2, 0.09, 0.01])].copy()
    b[5] = 0
    cases.append(("skewed", b))
    cases.append(("all 256 bytes", rng.integers(0, 256, 100000).astype(np.uint8)))
    cases.append(("few of many bytes", rng.integers(0, 256, 400).astype(np.uint8)))
    b = np.full(200000, ord("T"), dtype=np.uint8)
    b[17] = 0
    cases.append(("one letter", b))
    p = np.array([2.0 ** -i for i in range(1, 25)])
    b = (rng.choice(24, 500000, p=p / p.sum()) + 65).astype(np.uint8)
    b[9] = 0
    cases.append(("deep tree", b))
    return cases


def test_wavelet_tree_equals_sdsls(tool, probe, tmp_path):
    seq, ref, mine = (os.path.join(tmp_path, x) for x in ("bytes", "ref", "mine"))
    for name, b in _bwt_cases():
        b.tofile(seq)
        subprocess.check_call([probe, "wt", seq, ref], cwd=tmp_path)
        for threads in ("1", "5"):
            subprocess.check_call([tool, "wt", seq, mine, threads])
            assert _same(ref, mine), (name, threads)


def test_c_abi_entry_is_exported_and_rejects_null():
    import ctypes as C
    lib = C.CDLL(os.path.join(LIBDIR, "libgsx.so"))
    assert hasattr(lib, "gsx_index_save_reference_format")
    assert lib.gsx_index_save_reference_format(None, b"x") != 0


# ---- GPU ------------------------------------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("case", ["g200k", "g150kN"])
def test_gpu_built_index_saved_in_reference_format_is_the_reference_s_file(golden_dir, golden_index, tmp_path, case):
    import gsx
    fasta, _ = golden_dir[case]
    ix = gsx.Index.build(fasta)
    prefix = os.path.join(tmp_path, case)
    ix.save_reference_format(prefix)
    ix.close()
    for ext in (".forward", ".reverse", ".gs"):
        assert _same(prefix + ext, golden_index[case] + ext), ext


@pytest.mark.gpu
def test_unmodified_reference_enumerates_over_an_index_written_by_the_cli(golden_dir, tmp_path):
    import oracle as O
    if not O.have_ref():
        pytest.skip("oracle/_ref/guidescan not built")
    fasta, guides = golden_dir["g150kN"]
    prefix = os.path.join(tmp_path, "ix")
    exe = os.path.join(LIBDIR, "bin", "guidescan")
    r = subprocess.run([exe, "index", "--index", prefix, "--reference-format", fasta], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Wrote" in r.stdout and os.path.exists(prefix + ".forward") and os.path.exists(prefix + ".gsx")
    out = os.path.join(tmp_path, "ref.out")
    O.ref_enumerate(prefix, guides, out, mismatches=3)
    assert open(out, "rb").read() == golden_output("g150kN", "m3_csv")
    # both formats at the prefix: gsx_index_open takes the more recently written one (the CLI leaves that to be <prefix>.gsx)
    import gsx
    import shutil
    good = open(prefix + ".forward", "rb").read()
    open(prefix + ".forward", "wb").write(good[:4096])                      # not an index any more ...
    os.utime(prefix + ".forward", ns=(0, os.stat(prefix + ".gsx").st_mtime_ns - 10**9))   # ... but older than <prefix>.gsx: not looked at
    gsx.Index.open(prefix).close()
    os.utime(prefix + ".forward")                                          # newer: looked at, and refused
    with pytest.raises(gsx.GsxError, match="malformed index file"):
        gsx.Index.open(prefix)
    open(prefix + ".forward", "wb").write(good)
    # and the product opens its own reference-format files like the reference's
    os.remove(prefix + ".gsx")
    ix = gsx.Index.open(prefix)
    res = ix.enumerate_file(guides, os.path.join(tmp_path, "gpu.out"), gsx.make_params(mismatches=3))
    ix.close()
    assert res[0] > 0 and open(os.path.join(tmp_path, "gpu.out"), "rb").read() == golden_output("g150kN", "m3_csv")
