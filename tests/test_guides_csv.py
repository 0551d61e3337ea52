"""Guides CSV ingest (gsx_guides_csv_open, csrc/gsx_format.cpp::read_guides_csv) on the host: same rows, trimming and error
behaviour as the reference's kmer_producer (src/genomics/kmer.cxx:9-25 over fast-cpp-csv-parser with trim_chars<' ','\\t'>,
ignore_no_column, no quoting).  No GPU."""
import os
import random

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def gsx():
    import gsx as g
    return g


def _restate(text):
    """what the reference's reader yields: header names any order, fields trimmed of spaces/tabs, blank lines skipped"""
    lines = text.split("\n")
    hdr = [h.strip(" \t\r") for h in lines[0].rstrip("\r").split(",")]
    rows = []
    for l in lines[1:]:
        l = l.rstrip("\r\n")
        if not l:
            continue
        f = [x.strip(" \t\r") for x in l.split(",")]
        assert len(f) == len(hdr)
        d = dict(zip(hdr, f))
        rows.append((d["id"], d["sequence"], d["pam"], d["sense"] == "+"))
    return rows


@pytest.mark.parametrize("case", ["g200k", "g150kN"])
def test_golden_guides_files(gsx, case):
    path = os.path.join(ROOT, "tests", "golden", case + ".guides.csv")
    assert gsx.read_guides_csv(path) == _restate(open(path).read())


def test_column_order_trimming_crlf_blank_lines_no_final_newline(gsx, tmp_path):
    text = ("sense , pam,position,\tid ,chromosome,sequence\r\n"
            "+,NGG,12, g0 ,chr1, ACGTACGTACGTACGTACGT \r\n"
            "\r\n"
            "-,\tNAG ,,g1,,TTTTACGTACGTACGTACGA\n"
            "\n"
            " + ,,7,g2,chr2,ACGT")
    p = os.path.join(tmp_path, "a.csv")
    open(p, "w", newline="").write(text)
    assert gsx.read_guides_csv(p) == [("g0", "ACGTACGTACGTACGTACGT", "NGG", True), ("g1", "TTTTACGTACGTACGTACGA", "NAG", False),
                                     ("g2", "ACGT", "", True)]
    assert gsx.read_guides_csv(p) == _restate(text)


@pytest.mark.parametrize("text,msg", [
    ("id,sequence,pam,chromosome,position\ng,ACGT,NGG,chr1,1\n", 'Missing column "sense"'),
    ("id,sequence,pam,chromosome,position,sense,score\n", 'Extra column "score"'),
    ("id,sequence,pam,chromosome,position,sense\ng0,ACGT,NGG,chr1,1,+\ng1,ACGT,NGG,chr1,+\n", "wrong number of columns in kmers file line: g1,ACGT,NGG,chr1,+"),
    ("id,sequence,pam,chromosome,position,sense\ng0,ACGT,NGG,chr1,1,+,9\n", "wrong number of columns"),
    ("", "empty kmers file"),
])
def test_error_behaviour(gsx, tmp_path, text, msg):
    p = os.path.join(tmp_path, "bad.csv")
    open(p, "w").write(text)
    with pytest.raises(gsx.GsxError) as e:
        gsx.read_guides_csv(p)
    assert msg in str(e.value)
    with pytest.raises(gsx.GsxError):
        gsx.read_guides_csv(os.path.join(tmp_path, "missing.csv"))


def test_large_file_parsed_by_several_threads(gsx, tmp_path):
    """~16 MB = 15 pieces, so every piece boundary lands in the middle of some line; a bad line late in the file is still reported"""
    rnd = random.Random(5)
    rows = ["id,sequence,pam,chromosome,position,sense"]
    for i in range(300_000):
        rows.append("chr%d:%d:%s,%s,NGG,chr%d,%d,%s" % (i % 24, i * 7, "+-"[i & 1], "".join(rnd.choice("ACGT") for _ in range(20)), i % 24, i * 7, "+-"[i & 1]))
    text = "\n".join(rows) + "\n"
    assert len(text) > 15 << 20
    p = os.path.join(tmp_path, "big.csv")
    open(p, "w").write(text)
    got = gsx.read_guides_csv(p)
    assert len(got) == 300_000 and got == _restate(text)
    rows[250_001] = "oops,ACGT"
    open(p, "w").write("\n".join(rows) + "\n")
    with pytest.raises(gsx.GsxError) as e:
        gsx.read_guides_csv(p)
    assert "wrong number of columns in kmers file line: oops,ACGT" in str(e.value)


def test_output_file_is_written_by_several_threads_in_order(tmp_path):
    """the whole-file driver's writer (gsx_format.cpp OutFile): the formatting workers' slices reach the file at their own offsets on their
    own threads (pwrite), through a mapped extent, or as one sequential stream: same bytes in every mode, for slices that start and end
    anywhere in a page; sequential writes for outputs that are not regular files"""
    import ctypes as C
    import random
    import gsx
    lib = C.CDLL(gsx.LIB_PATH)
    lib.gsx_internal_write_parts.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_size_t, C.c_size_t]
    rnd = random.Random(5)
    for sizes in ([1], [0, 5, 0], [4095, 1, 4096, 8193], [100000, 0, 3, 777777, 12], [rnd.randrange(0, 300000) for _ in range(16)]):
        parts = [bytes(rnd.randrange(33, 127) for _ in range(min(n, 997))) * (n // 997 + 1) for n in sizes]
        parts = [p[:n] for p, n in zip(parts, sizes)]
        arr = (C.c_char_p * len(parts))(*parts)
        lens = (C.c_size_t * len(parts))(*sizes)
        for env in (None, "write", "mmap", "pwrite"):
            if env:
                os.environ["GSX_OUT_MODE"] = env
            else:
                os.environ.pop("GSX_OUT_MODE", None)
            out = os.path.join(tmp_path, "w.out")
            assert lib.gsx_internal_write_parts(out.encode(), arr, lens, len(parts), 3) == 0
            assert open(out, "rb").read() == b"".join(parts) * 3
        os.environ.pop("GSX_OUT_MODE", None)
    assert lib.gsx_internal_write_parts(b"/dev/null", arr, lens, len(parts), 2) == 0
    assert lib.gsx_internal_write_parts(os.path.join(tmp_path, "no", "such", "dir", "x").encode(), arr, lens, len(parts), 1) != 0
