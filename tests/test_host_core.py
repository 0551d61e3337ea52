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
