"""bench.py's own arm assembled without a GPU: the index and the results are stubs, everything else (workload, timed loop with
one or several batches in flight, JSON line with roofline / e2e / file_e2e / cpu_baseline / clocks blocks) is the real code.  Guards the
driver-facing contract against runtime errors that only a GPU box would otherwise reveal."""
import json
import os

import numpy as np
import pytest


class _Result:
    def __init__(self, n):
        self.n_guides, self.n_hits = n, 5 * n

    def counters(self):
        keys = ("nodes", "lookups", "matches", "hits", "lf_steps", "spills", "ms_search", "ms_arrange", "ms_locate", "ms_score", "ms_total_device",
                "ms_h2d", "ms_d2h", "launches", "ms_sweep", "seeds", "ms_prepare", "ms_wall", "sectors", "edited_guides")
        return {k: 1.0 for k in keys}

    def guide_arrays(self):
        return {"specificity": np.ones(self.n_guides, dtype=np.float32)}

    def close(self):
        pass


class _Index:
    device_bytes = 1e9

    def enumerate_raw(self, arr, per, params):
        return _Result(per)

    def enumerate_start(self, arr, per, params):
        return per

    def enumerate_wait(self, pending):
        return _Result(pending)

    def open_seconds(self):
        return (1.0, 2.0, 0.0)

    def enumerate_file(self, csv, out, params, **kw):
        open(out, "w").write("x" * 1000)
        return 10, _Result(10).counters()

    def close(self):
        pass


class _Clocks:
    def __init__(self, local):
        pass

    def start(self):
        pass

    def finish(self):
        return {"sm_mhz": 1, "sm_max_mhz": 1, "reasons": [], "samples": 1}


@pytest.mark.parametrize("extra", [[], ["--e2e-pipeline", "1", "--n-runs", "2"], ["--e2e-pipeline", "3"], ["--rna-bulges", "1", "--dna-bulges", "1", "--alt-pam", "NAG"]])
def test_own_arm_prints_the_contract_line(monkeypatch, tmp_path, capsys, extra):
    import bench
    monkeypatch.setattr(bench, "build_index", lambda gsx, g, chroms, local, args, workdir: (_Index(), "stub"))
    monkeypatch.setattr(bench, "cpu_baseline", lambda *a, **k: {"value": 1.0, "unit": "guides/s", "cores": 1, "kind": "port", "sample": "stub", "out": "o", "csv": "c"})
    monkeypatch.setattr(bench, "parity_on_sample", lambda *a, **k: True)
    monkeypatch.setattr(bench, "ClockSampler", _Clocks)
    args = bench.parse_args(["--genome-mb", "1", "--n-chr", "2", "--guides-per-step", "50", "--steps", "4", "--warmup", "3", "--plant-guides", "10",
                             "--workdir", str(tmp_path)] + extra)
    bench.apply_variant(args)
    bench.run_gsx(args)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks", "file_e2e"):
        assert key in line, key
    assert line["steps"] == 4 and line["warmup"] == 3 and line["n_gpus"] == 1 and line["unit"] == "guides/s"
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    assert line["config"]["workload"].startswith("1 Mb") and line["value"] == pytest.approx(50 * 4 / 4e-3)
