"""bench.py's host-side pieces that need no GPU: the workload table against BASELINE.json, one `config` for both arms,
the kernel label mirroring the library's dispatch, the traffic file's shape, the committed bench records' contract keys."""
import glob
import importlib.util
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_workloads_cover_baseline_configs(bench):
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert len(base["configs"]) == 5
    for name in ("c1", "c2", "c3", "c4", "c5"):
        assert name in bench.WORKLOADS
    assert bench.WORKLOADS["c2"]["dims"] == (1024, 1024) or list(bench.WORKLOADS["c2"]["dims"]) == [1024, 1024]
    assert list(bench.WORKLOADS["c3"]["dims"]) == [64] * 4 and list(bench.WORKLOADS["c4"]["dims"]) == [256] * 4


def test_config_is_one_function_of_the_workload(bench):
    for name, wl in bench.WORKLOADS.items():
        a = bench.make_config(name, wl, 1, "fast")
        b = bench.make_config(name, wl, 1, "fast")
        assert a == b and a["workload"] == name and "model" not in a
        assert set(a) >= {"workload", "dims", "dtau", "tau_steps_per_step", "potential", "math", "seed", "parallelism"}


def test_kernel_label_mirrors_dispatch(bench, monkeypatch):
    monkeypatch.delenv("SQ_ROWS", raising=False)
    assert bench.kernel_name((1024, 1024), "f32") == ("rowres_kernel", True)
    assert bench.kernel_name((64, 64, 64, 64), "f32") == ("lattice_tile_kernel", False)
    assert bench.kernel_name((64, 64, 64, 64), "f64")[0] == "lattice_step_kernel"
    assert bench.kernel_name((1000, 1000), "f32")[0] == "lattice_step_kernel"  # row length not a multiple of 128
    monkeypatch.setenv("SQ_ROWS", "1")
    assert bench.kernel_name((256, 256, 256, 32), "f32")[0] == "lattice_rows_kernel"


def test_traffic_file_shape(bench):
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    for name, e in t.items():
        assert name in bench.WORKLOADS
        assert e["bytes_per_launch"] > 0 and e["tau_steps_per_launch"] >= 1 and e["kernel"] and e["source"]
        kname, _ = bench.kernel_name(tuple(bench.WORKLOADS[name]["dims"]), bench.WORKLOADS[name]["real"])
        assert e["kernel"].startswith(kname), (name, e["kernel"], kname)


def test_committed_bench_lines_keep_the_contract():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench", "bench_*.json")) +
                   glob.glob(os.path.join(ROOT, "profiles", "r02_bench", "headcheck_bench_*.json")))
    assert len(files) > 20
    for f in files:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
                  "data", "config"):
            assert k in d, (f, k)
        assert d["vs_baseline"] is None and d["config"]["workload"]
        if d.get("impl") == "reference":
            assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] in ("port", "reference")
        if d.get("impl") != "reference" and d["n_gpus"] == 1 and d["config"]["workload"] in ("c2", "c3", "slab", "c5"):
            r = d["roofline"]   # the GPU arm's line: roofline object, clocks, launches of our own kernels
            assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
            assert d["gpu_launches"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if d["n_gpus"] > 1 and d["config"]["workload"] == "c2":
            assert d["ring"]["ring_parity"] == "bit-identical" and d["ring_parity"] == "bit-identical"
