"""CPU tests of the device arithmetic (gsx_core.h) through tests/host_core_check.cpp: the product's index loader,
child generation, keys, locate, coordinates, CFD, specificity and text formatter, run sequentially on the host and
diffed byte-for-byte against the reference's golden output.  No GPU, no compute call into libgsx.so kernels."""
import os
import subprocess

import pytest

from conftest import ROOT, golden_cases, golden_manifest, golden_output, variant_cli_args

HARNESS = os.path.join(ROOT, "tests", "_build", "host_core_check")
LIBDIR = os.path.join(ROOT, "guidescan-cli_b200")


@pytest.fixture(scope="session")
def harness():
    lib = os.path.join(LIBDIR, "libgsx.so")
    if not os.path.exists(lib):
        pytest.skip("libgsx.so not built (run __graft_entry__.build())")
    src = os.path.join(ROOT, "tests", "host_core_check.cpp")
    if not os.path.exists(HARNESS) or os.path.getmtime(HARNESS) < max(os.path.getmtime(src), os.path.getmtime(lib)):
        os.makedirs(os.path.dirname(HARNESS), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I/usr/local/cuda/include", "-o", HARNESS, src,
                               "-L" + LIBDIR, "-lgsx", "-Wl,-rpath," + LIBDIR])
    return HARNESS


@pytest.mark.timeout(600)
@pytest.mark.parametrize("case,variant", golden_cases())
def test_device_arithmetic_on_host_matches_golden(harness, golden_dir, golden_index, tmp_path, case, variant):
    kw = golden_manifest()["cases"][case]["variants"][variant]["opts"]
    out = os.path.join(tmp_path, "h.out")
    subprocess.check_call([harness, golden_index[case], golden_dir[case][1], out] + variant_cli_args(kw),
                          stderr=subprocess.DEVNULL)
    assert open(out, "rb").read() == golden_output(case, variant)


@pytest.mark.timeout(600)
@pytest.mark.parametrize("m", [2, 3, 4])
def test_lookahead_pruning_keeps_output_and_cuts_nodes(harness, tmp_path, m):
    """viable_children (gsx_core.h), the pruning step of the specialised search kernel, on host-built look-ahead planes:
    identical text, fewer expanded nodes."""
    import oracle as O
    import synth
    if not O.have_ref():
        pytest.skip("needs oracle/_ref/guidescan for the index files")
    d = str(tmp_path)
    synth.make_dataset(d, 400_000, 3, 60, seed=31, name="la")
    fa, gcsv = os.path.join(d, "la.fa"), os.path.join(d, "la.guides.csv")
    O.ref_index(fa, os.path.join(d, "la"), cwd=d)
    O.Index(fa).enumerate_file(O.make_opts(mismatches=m), gcsv, os.path.join(d, "o.out"), nthreads=4)
    nodes = {}
    for tag, extra in (("plain", []), ("look", ["--lookahead"]), ("ftab", ["--ftab", "7"]), ("both", ["--lookahead", "--ftab", "8"]),
                       ("sweep1", ["--lookahead", "--ftab", "8", "--sweep", "1"]), ("sweep3", ["--lookahead", "--ftab", "8", "--sweep", "3"]),
                       ("sweep5", ["--lookahead", "--ftab", "9", "--sweep", "5"]),
                       # shallow tables: intervals of ~24 and ~98 rows, i.e. summaries with and without the "wide" pass-through
                       ("sweepw", ["--lookahead", "--ftab", "7", "--sweep", "2"]), ("sweepww", ["--lookahead", "--ftab", "6", "--sweep", "1"])):
        out = os.path.join(d, tag + ".out")
        r = subprocess.run([harness, os.path.join(d, "la"), gcsv, out, "-m", str(m)] + extra, capture_output=True, text=True, check=True)
        assert open(out, "rb").read() == open(os.path.join(d, "o.out"), "rb").read()
        nodes[tag] = int(r.stderr.split(" guides, ")[1].split(" nodes")[0])
    assert nodes["look"] < 0.8 * nodes["plain"], nodes
    # on a genome this small the pruned walk already dies near the root, so the table need not beat it; it must beat
    # the plain walk, and pruning must help behind the table as well
    assert nodes["ftab"] < nodes["plain"] and nodes["both"] < nodes["ftab"], nodes
    # the slice-major front end filters level-L nodes before the walk expands them
    assert nodes["sweep1"] < nodes["both"] and nodes["sweep3"] == nodes["sweep1"], nodes


def _golden_subset(case, tmp_path, keep):
    """guides file restricted to rows satisfying keep(fields) + the matching slice of a golden output"""
    lines = open(os.path.join(ROOT, "tests", "golden", case + ".guides.csv")).read().splitlines()
    rows = [l for l in lines[1:] if keep(l.split(","))]
    path = os.path.join(tmp_path, "subset.csv")
    with open(path, "w") as f:
        f.write("\n".join([lines[0]] + rows) + "\n")
    ids = {l.split(",")[0] for l in rows}

    def slice_of(text: str, sam: bool) -> str:
        out = []
        for i, l in enumerate(text.splitlines(True)):
            if (sam and l.startswith("@")) or (not sam and i == 0) or l.split("\t" if sam else ",", 1)[0] in ids:
                out.append(l)
        return "".join(out)
    return path, slice_of


BULGE_VARIANTS = [v for c, v in golden_cases() if c == "g200k" and ("_r" in v or "_d" in v)]


@pytest.mark.timeout(900)
@pytest.mark.parametrize("variant", BULGE_VARIANTS)
@pytest.mark.parametrize("mirror", ["root", "table"])
def test_bulges_as_edited_guides_match_golden(harness, golden_dir, golden_index, tmp_path, monkeypatch, variant, mirror):
    """gsx_core.h variant_pack / variant_rewrite + gsx_host.h bulge_variants: the bulge search restated as mismatch-only
    searches of edited guides (what gsx_enumerate runs on the specialised kernels when Prepared::variant_ok) gives the
    reference's text byte for byte -- including two bulges of each kind, alternative PAMs and guides of 19-21 nt."""
    kw = golden_manifest()["cases"]["g200k"]["variants"][variant]["opts"]
    if mirror == "table" and (kw.get("alt_pams") or kw.get("rna_bulges", 0) + kw.get("dna_bulges", 0) > 2):
        pytest.skip("the table / look-ahead mirrors run single-PAM passes; the 600k-variant case runs once")
    monkeypatch.setenv("GSX_VARIANTS_MAX", "100000000")
    gcsv, slice_of = _golden_subset("g200k", str(tmp_path), lambda f: f[2] == "NGG" and set(f[1]) <= set("ACGT"))
    out = os.path.join(tmp_path, "h.out")
    extra = ["--variants"] + (["--lookahead", "--ftab", "7"] if mirror == "table" else [])
    subprocess.check_call([harness, golden_index["g200k"], gcsv, out] + variant_cli_args(kw) + extra, stderr=subprocess.DEVNULL)
    assert open(out).read() == slice_of(golden_output("g200k", variant).decode(), kw.get("fmt") == "sam")


@pytest.mark.timeout(600)
@pytest.mark.parametrize("threads", [2, 3, 7, 64])
@pytest.mark.parametrize("case,variant", [("g200k", "m3_csv"), ("g200k", "m1_r1_d1_csv"), ("g150kN", "m3_altNAG_sam"), ("g200k", "m3_thr1_csv"),
                                          ("g150kN", "m4_max2_csv")])
def test_formatter_slices_by_rows_keep_the_text(harness, golden_dir, golden_index, tmp_path, monkeypatch, case, variant, threads):
    """gsx_format_rows cuts the guides into slices of about equal numbers of rows, one host thread each (bulge batches have
    thousands of rows per guide); forced here on small golden cases, incl. more threads than guides with hits"""
    monkeypatch.setenv("GSX_FORMAT_THREADS", str(threads))
    kw = golden_manifest()["cases"][case]["variants"][variant]["opts"]
    out = os.path.join(tmp_path, "h.out")
    subprocess.check_call([harness, golden_index[case], golden_dir[case][1], out] + variant_cli_args(kw), stderr=subprocess.DEVNULL)
    assert open(out, "rb").read() == golden_output(case, variant)


ALT_PAM_VARIANTS = [v for c, v in golden_cases() if c == "g200k" and "alt" in v]


@pytest.mark.timeout(600)
@pytest.mark.parametrize("variant", ALT_PAM_VARIANTS)
@pytest.mark.parametrize("mirror", ["root", "table", "sweep"])
def test_alternative_pams_in_one_pass_match_golden(harness, golden_dir, golden_index, tmp_path, variant, mirror):
    """gsx_core.h fused_filter_pampack / fused_pam_ok: one search with the filter PAM (wildcard where the PAMs differ) + a check
    of the consumed PAM characters gives the union of the per-PAM searches -- two and three PAMs (with cross combinations
    that spell no real PAM), also together with bulges through edited guides.  (The GPU path behind GSX_FUSED_PAMS=1.)"""
    kw = golden_manifest()["cases"]["g200k"]["variants"][variant]["opts"]
    bulges = kw.get("rna_bulges") or kw.get("dna_bulges")
    gcsv, slice_of = _golden_subset("g200k", str(tmp_path), lambda f: f[2] == "NGG" and set(f[1]) <= set("ACGT"))
    out = os.path.join(tmp_path, "h.out")
    extra = ["--fused"] + (["--variants"] if bulges else []) + {"root": [], "table": ["--lookahead", "--ftab", "7"],
                                                                "sweep": ["--lookahead", "--ftab", "8", "--sweep", "2"]}[mirror]
    subprocess.check_call([harness, golden_index["g200k"], gcsv, out] + variant_cli_args(kw) + extra, stderr=subprocess.DEVNULL)
    assert open(out).read() == slice_of(golden_output("g200k", variant).decode(), kw.get("fmt") == "sam")


@pytest.mark.timeout(600)
@pytest.mark.parametrize("variant", [v for v in BULGE_VARIANTS if "_d" in v and "r2" not in v and "alt" not in v])
def test_sweep_skips_patterns_that_substitute_an_inserted_position(harness, golden_dir, golden_index, tmp_path, variant):
    """gsx_core.h variant_forced_mask / forced_kept (sweep_kernel<..., FORCED>, GSX_FORCED_SWEEP=1): an inserted position of an
    edited guide must match exactly, so the sweep need not visit patterns that substitute it -- same text, fewer nodes"""
    kw = golden_manifest()["cases"]["g200k"]["variants"][variant]["opts"]
    gcsv, slice_of = _golden_subset("g200k", str(tmp_path), lambda f: f[2] == "NGG" and set(f[1]) <= set("ACGT"))
    nodes = {}
    for tag, extra in (("all", []), ("forced", ["--forced"])):
        out = os.path.join(tmp_path, tag + ".out")
        r = subprocess.run([harness, golden_index["g200k"], gcsv, out] + variant_cli_args(kw) + ["--variants", "--lookahead", "--ftab", "8", "--sweep", "2"] + extra,
                           capture_output=True, text=True, check=True)
        assert open(out).read() == slice_of(golden_output("g200k", variant).decode(), kw.get("fmt") == "sam")
        nodes[tag] = int(r.stderr.split(" guides, ")[1].split(" nodes")[0])
    assert nodes["forced"] <= nodes["all"] and (kw.get("mismatches", 3) == 0 or nodes["forced"] < nodes["all"]), nodes


def test_per_layout_row_filters_equal_the_general_ones(harness):
    """gsx_core.h summary_exact_shape / summary_masks_shape (the compiled-per-plane-layout filters of sweep_lean_kernel) against
    summary_eval_exact / summary_eval_masks on 600 k random sectors and guides of every compiled layout; sweep_shape_of classification"""
    r = subprocess.run([harness, "--shape-selftest"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "shape selftest ok" in r.stdout


@pytest.mark.timeout(600)
@pytest.mark.parametrize("case,variant", [("g200k", "m3_csv"), ("g200k", "m3_altNAG_sam"), ("g150kN", "m1_r1_d1_csv"), ("g200k", "m4_max1_sam"), ("g150kN", "m3_thr1_csv")])
@pytest.mark.parametrize("parts", [2, 3, 64])
def test_results_in_several_parts_format_and_merge_like_one(harness, golden_dir, golden_index, tmp_path, case, variant, parts):
    """what gsx_enumerate assembles from several devices, without a GPU: the result cut into contiguous guide shards (64: more parts than
    guides, so some are empty) formats to the golden text -- the formatter reads the parts in place -- and the merged view built by
    gsx_result_view_get, first_hit and gsx_result_match_sequence equal the unsplit arrays"""
    kw = golden_manifest()["cases"][case]["variants"][variant]["opts"]
    out = os.path.join(tmp_path, "h.out")
    r = subprocess.run([harness, golden_index[case], golden_dir[case][1], out] + variant_cli_args(kw) + ["--parts", str(parts)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(out, "rb").read() == golden_output(case, variant)
