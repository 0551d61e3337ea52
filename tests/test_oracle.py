"""CPU tests: the oracle restatement (oracle/gs_oracle.c) against the golden vectors produced by the
unmodified reference binary, and -- when oracle/_ref/guidescan is present -- against fresh reference runs."""
import os
import random

import pytest

import oracle as O
from conftest import golden_cases, golden_manifest, golden_output

pytestmark = pytest.mark.timeout(600)

_IDX = {}


def _index(golden_dir, case):
    if case not in _IDX:
        _IDX[case] = O.Index(golden_dir[case][0])
    return _IDX[case]


@pytest.mark.parametrize("case,variant", golden_cases())
def test_oracle_matches_reference_golden(golden_dir, tmp_path, case, variant):
    kw = dict(golden_manifest()["cases"][case]["variants"][variant]["opts"])
    if "alt_pams" in kw:
        kw["alt_pams"] = tuple(kw["alt_pams"])
    out = os.path.join(tmp_path, "o.out")
    _index(golden_dir, case).enumerate_file(O.make_opts(**kw), golden_dir[case][1], out, nthreads=4)
    assert open(out, "rb").read() == golden_output(case, variant)


def test_oracle_fm_index_self_consistency(golden_dir):
    """rank_bwt against a brute-force count, LF-walk locate against the full suffix array."""
    ix = _index(golden_dir, "g150kN")
    n = ix.n
    rnd = random.Random(5)
    L = O.lib()
    for strand in (0, 1):
        bwt = bytes(L.gso_bwt(ix.h, strand, i) for i in range(n))
        assert bwt.count(b"\0") == 1
        for _ in range(300):
            i = rnd.randrange(0, n + 1)
            c = rnd.choice(b"ACGTN")
            assert ix.rank_bwt(strand, i, c) == bwt[:i].count(bytes([c]))
        for _ in range(300):
            r = rnd.randrange(0, n)
            assert ix.sa(strand, r) == L.gso_sa_direct(ix.h, strand, r)
        assert ix.rank_bwt(strand, n, ord("X")) == 0


def test_cfd_known_values():
    L = O.lib()
    g = b"ACGTACGTACGTACGTACGT"
    assert L.gso_calculate_cfd(g, g + b"AGG", b"AGG") == 1.0
    assert abs(L.gso_calculate_cfd(g, g + b"AAG", b"AAG") - 0.259259259) < 1e-6
    assert L.gso_calculate_cfd(g[:19], g, b"AGG") == 1.0          # CFD undefined off 20/3: 1.0
    assert L.gso_calculate_cfd(g, b"." + g[1:], b"AGG") == 0.0     # key absent from the table => 0.0


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/guidescan not built (needs /root/reference)")
@pytest.mark.parametrize("kw", [dict(mismatches=3), dict(mismatches=2, fmt="sam"),
                                dict(mismatches=1, rna_bulges=1, dna_bulges=1)])
def test_oracle_matches_live_reference(tmp_path, kw):
    import synth
    d = str(tmp_path)
    synth.make_dataset(d, 300_000, 5, 30, seed=99, name="live", n_runs_of_N=3)
    fa, gcsv = os.path.join(d, "live.fa"), os.path.join(d, "live.guides.csv")
    O.ref_index(fa, os.path.join(d, "live"), cwd=d)
    O.ref_enumerate(os.path.join(d, "live"), gcsv, os.path.join(d, "r.out"), **kw)
    ix = O.Index(fa)
    ix.enumerate_file(O.make_opts(**kw), gcsv, os.path.join(d, "o.out"), nthreads=2)
    assert open(os.path.join(d, "o.out"), "rb").read() == open(os.path.join(d, "r.out"), "rb").read()
