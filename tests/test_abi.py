"""CPU tests of the drop-in boundary: libgsx.so loads, exports every symbol include/gsx.h declares, and fails loudly
(no CPU fallback) when there is no device.  No compute calls."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT

LIB = os.path.join(ROOT, "guidescan-cli_b200", "libgsx.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        pytest.skip("libgsx.so not built (run __graft_entry__.build())")
    return C.CDLL(LIB)


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gsx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gsx_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("gsx_index_open", "gsx_index_build", "gsx_enumerate", "gsx_result_view_get", "gsx_format_rows", "gsx_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_lists_the_same_symbols(lib):
    import gsx
    assert sorted(gsx.EXPORTS) == declared_symbols()


def test_version_and_defaults(lib):
    import gsx
    lib.gsx_version.restype = C.c_char_p
    assert lib.gsx_version() == b"2.0.0"           # reference include/version.hpp:2
    p = gsx.Params()
    lib.gsx_params_default(C.byref(p))
    assert (p.mismatches, p.rna_bulges, p.dna_bulges, p.max_bulge_size, p.threshold, p.max_off_targets, p.start) == (3, 0, 0, 1, -1, -1, 0)


def test_errors_are_loud_without_files_or_device(lib, tmp_path):
    import gsx
    with pytest.raises(gsx.GsxError) as e:
        gsx.Index.open(os.path.join(tmp_path, "nothing"))
    assert e.value.code == 2 and "genome structure" in str(e.value)      # GSX_ERR_IO, reference src/guidescan.cxx:193-196
    if gsx.device_count() == 0:
        open(os.path.join(tmp_path, "x.gs"), "w").write("chr1\n10\n")
        for ext in (".forward", ".reverse"):
            open(os.path.join(tmp_path, "x" + ext), "wb").write(b"\0" * 16)
        with pytest.raises(gsx.GsxError) as e:
            gsx.Index.open(os.path.join(tmp_path, "x"))
        assert e.value.code in (2, 3)      # malformed file (IO) is detected before the device is needed
        with pytest.raises(gsx.GsxError) as e:
            gsx.Index.build(os.path.join(ROOT, "tests", "golden", "g200k.guides.csv"))
        assert e.value.code == 3           # GSX_ERR_NO_DEVICE: there is no CPU path


def test_cli_binary_mirrors_reference_surface():
    import subprocess
    exe = os.path.join(ROOT, "guidescan-cli_b200", "bin", "guidescan")
    if not os.path.exists(exe):
        pytest.skip("CLI not built")
    assert subprocess.run([exe, "--version"], capture_output=True, text=True).stdout.strip() == "2.0.0"
    r = subprocess.run([exe, "enumerate", "-f", "x.csv", "-o", "y"], capture_output=True, text=True)
    assert r.returncode != 0 and "index is required" in r.stderr
    r = subprocess.run([exe, "enumerate", "idx", "-f", "x.csv", "-o", "y", "--format", "bam"], capture_output=True, text=True)
    assert r.returncode != 0
