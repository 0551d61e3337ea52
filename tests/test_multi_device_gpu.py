"""GPU tests of the product's own sharding (run with -m gpu): gsx_index_open on several devices, gsx_enumerate /
gsx_enumerate_file over them, the two-slot start / wait form, and the `guidescan` binary end to end.  The reference's
counterpart is the split of the guide list over worker threads inside one process (src/guidescan.cxx:225-251): whatever
the number of shards, the output is the single-shard output.  A device named twice gives two job slots over one copy of
the index, which exercises the merge of the per-device result parts on a one-GPU box; [0, 1] runs where a second GPU exists."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden_manifest, golden_output

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

CASES = [("g200k", "m3_csv"), ("g200k", "m1_r1_d1_csv"), ("g150kN", "m3_altNAG_sam"), ("g200k", "m3_thr1_csv"), ("g150kN", "m3_csv"),
         ("g200k", "m4_max2_csv")]


@pytest.fixture(scope="module")
def gsx():
    import gsx as g
    if g.device_count() < 1:
        pytest.fail("no CUDA device: the product has no CPU path")
    return g


def _device_lists(g):
    out = [[0, 0], [0, 0, 0]]
    if g.device_count() >= 2:
        out += [[0, 1], [1, 0, 1]]
    return out


def _params(gsx, kw):
    from test_parity_gpu import _params as p
    return p(gsx, kw)


@pytest.fixture(scope="module")
def shard_index(gsx, golden_index):
    cache = {}

    def get(case, devices):
        key = (case, tuple(devices))
        if key not in cache:
            cache[key] = gsx.Index.open(golden_index[case], devices=list(devices))
        return cache[key]
    yield get
    for ix in cache.values():
        ix.close()


@pytest.mark.parametrize("case,variant", CASES)
def test_sharded_enumerate_file_equals_golden(gsx, shard_index, golden_dir, tmp_path, case, variant):
    kw = golden_manifest()["cases"][case]["variants"][variant]["opts"]
    for devices in _device_lists(gsx):
        ix = shard_index(case, devices)
        assert ix.n_devices == len(devices)
        out = os.path.join(tmp_path, "g.out")
        for batch in (0, 7):                      # whole file in one call; batches of 7 guides (shards of 2-4 guides, some of them empty at the tail)
            ix.enumerate_file(golden_dir[case][1], out, _params(gsx, kw), fmt=kw.get("fmt", "csv"), mode=kw.get("mode", "complete"), batch_guides=batch)
            assert open(out, "rb").read() == golden_output(case, variant), (devices, batch)


def test_sharded_result_arrays_equal_single_device(gsx, shard_index, golden_dir):
    rows = gsx.read_guides_csv(golden_dir["g200k"][1])
    guides = [(s, p) for _, s, p, _ in rows if set(s) <= set("ACGT") and p == "NGG"]
    p = gsx.make_params(mismatches=3)
    ref = shard_index("g200k", [0]).enumerate(guides, p)
    ga0, ha0 = ref.guide_arrays(), ref.hit_arrays()
    seqs0 = [ref.match_sequence(h) for h in range(ref.n_hits)]
    for devices in _device_lists(gsx):
        r = shard_index("g200k", devices).enumerate(guides, p)
        ga, ha = r.guide_arrays(), r.hit_arrays()
        for k in ga0:
            assert np.array_equal(ga[k], ga0[k]), (devices, k)
        for k in ha0:
            assert np.array_equal(ha[k], ha0[k]), (devices, k)
        assert [r.match_sequence(h) for h in range(r.n_hits)] == seqs0          # hits of every part, through part_h0 / part_g0
        c = r.counters()
        assert c["hits"] == ref.counters()["hits"] and c["matches"] == ref.counters()["matches"]
        r.close()
    # fewer guides than devices: some shards are empty
    for devices in _device_lists(gsx):
        r = shard_index("g200k", devices).enumerate(guides[:1], p)
        assert r.n_guides == 1 and int(r.guide_arrays()["n_hits"][0]) == int(ga0["n_hits"][0])
        one = shard_index("g200k", [0]).enumerate(guides[:1], p)
        row = [("x", guides[0][0], guides[0][1], True)]
        assert r.format(row, p) == one.format(row, p)
        one.close()
        r.close()
        r = shard_index("g200k", devices).enumerate([], p)
        assert r.n_guides == 0 and r.n_hits == 0
        r.close()
    ref.close()


def test_replicas_are_byte_identical(gsx, shard_index, golden_index):
    for devices in _device_lists(gsx):
        ix = shard_index("g150kN", devices)
        sums = [ix.device_checksum(s) for s in range(len(devices))]
        assert len(set(sums)) == 1, (devices, sums)
        t = ix.open_seconds()
        assert all(x >= 0 for x in t)
    if gsx.device_count() >= 2:
        # an index opened directly on the second device derives its arrays there: same digest as the peer copy
        alone = gsx.Index.open(golden_index["g150kN"], devices=[1])
        assert alone.device_checksum(0) == shard_index("g150kN", [0, 1]).device_checksum(1)
        alone.close()


def test_start_wait_pipelines_batches(gsx, shard_index, golden_dir):
    """gsx_enumerate_start / gsx_enumerate_wait: several batches in flight, results as from gsx_enumerate; errors surface at wait"""
    import ctypes as C
    rows = gsx.read_guides_csv(golden_dir["g200k"][1])
    guides = [(s, p) for _, s, p, _ in rows if set(s) <= set("ACGT") and p == "NGG"]
    p = gsx.make_params(mismatches=3)
    for devices in ([0], [0, 0]):
        ix = shard_index("g200k", devices)
        want = ix.enumerate(guides, p)
        batches = [guides[i::3] for i in range(3)]
        arrs, keep, pend = [], [], []
        for b in batches:
            arr = (gsx.Guide * len(b))()
            for i, (s, pam) in enumerate(b):
                bs = (s.encode(), pam.encode()); keep.append(bs); arr[i] = gsx.Guide(bs[0], bs[1])
            arrs.append(arr)
            pend.append(ix.enumerate_start(arr, len(b), p))
        spec = {}
        for k, (b, h) in enumerate(zip(batches, pend)):
            r = ix.enumerate_wait(h)
            for i, sp in enumerate(r.guide_arrays()["specificity"]):
                spec[k + 3 * i] = (float(sp), int(r.guide_arrays()["n_hits"][i]))
            r.close()
        ws, wn = want.guide_arrays()["specificity"], want.guide_arrays()["n_hits"]
        assert [spec[i] for i in range(len(guides))] == [(float(ws[i]), int(wn[i])) for i in range(len(guides))]
        want.close()
        bad = (gsx.Guide * 1)(gsx.Guide(b"", b"NGG"))
        h = ix.enumerate_start(bad, 1, p)
        with pytest.raises(gsx.GsxError) as e:
            ix.enumerate_wait(h)
        assert "length" in str(e.value)


def test_cli_index_then_enumerate(gsx, golden_dir, tmp_path):
    """bin/guidescan index + enumerate (reference src/guidescan.cxx:109-258) against the reference's golden text, on one device,
    on two job slots, and on two GPUs where present"""
    exe = os.path.join(ROOT, "guidescan-cli_b200", "bin", "guidescan")
    fa, gcsv = golden_dir["g200k"]
    prefix = os.path.join(tmp_path, "cli")
    r = subprocess.run([exe, "index", "--index", prefix, fa], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert os.path.exists(prefix + ".gsx") and os.path.exists(prefix + ".gs")
    runs = [["--gpus", "1"], ["--devices", "0,0"]]
    if gsx.device_count() >= 2:
        runs.append(["--gpus", "2"])
    for extra in runs:
        for variant, args in (("m3_csv", ["-m", "3"]), ("m3_altNAG_sam", ["-m", "3", "--format", "sam", "-a", "NAG"]),
                              ("m1_r1_d1_csv", ["-m", "1", "--rna-bulges", "1", "--dna-bulges", "1"]), ("m3_thr1_csv", ["-m", "3", "-t", "1"])):
            if variant not in golden_manifest()["cases"]["g200k"]["variants"]:
                continue
            out = os.path.join(tmp_path, "cli.out")
            cmd = [exe, "enumerate", prefix, "-f", gcsv, "-o", out, "-n", "3"] + extra + args      # (-a is greedy: last, as the manual writes it)
            r = subprocess.run(cmd, capture_output=True, text=True)
            assert r.returncode == 0, (cmd, r.stderr)
            assert "Processed" in r.stdout
            assert open(out, "rb").read() == golden_output("g200k", variant), cmd
    r = subprocess.run([exe, "enumerate", prefix, "-f", gcsv, "-o", os.path.join(tmp_path, "x"), "--gpus", "99"], capture_output=True, text=True)
    assert r.returncode != 0 and "device" in r.stderr
