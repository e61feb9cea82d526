"""bench.py on a box without a GPU: the reference arm prints the contract's JSON line (it times
the reference's own CPU Layer on a bounded sample), and the GPU arm fails loudly -- there is no
CPU fallback to fall back to."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=e,
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "20000")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle histories/s (whole box)"
    assert d["unit"] == "histories/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "histories/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert d["config"]["nb_cells"] == 1000 and d["config"]["histories_per_step"] == 20000


def test_reference_arm_other_ranks_print_nothing():
    r = run_bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                  "--cpu-sample", "2000", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_without_a_gpu_fails_loudly(mcb_lib):
    if mcb_lib.mcb200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = run_bench("--steps", "1", "--warmup", "1", "--no-cpu-baseline")
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_nvlink_summary_counts_24_bytes_per_record():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    s = bench.nvlink_summary([0, 100, 200], [300, 400, 0], steps=2, step_ms=1.0)
    assert s["records_per_step"] == 500 and s["bytes_per_step"] == 12000
    assert abs(s["busiest_link_direction_GB_per_s"] - 24 * 400 / 2e-3 / 1e9) < 1e-12


def test_settle_fast_regime_rebuilds_until_fast_or_gives_up():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    times, rebuilds = iter([404.0, 398.0, 341.0, 999.0]), []
    seen = bench.settle_fast_regime(lambda: next(times), lambda: rebuilds.append(1), best_ms=337.0)
    assert seen == [404.0, 398.0, 341.0] and len(rebuilds) == 2        # third build is fast
    times, rebuilds = iter([400.0] * 10), []
    seen = bench.settle_fast_regime(lambda: next(times), lambda: rebuilds.append(1), best_ms=337.0)
    assert seen == [400.0] * 4 and len(rebuilds) == 3                  # gives up after 3 rebuilds
    times, rebuilds = iter([330.0]), []
    assert bench.settle_fast_regime(lambda: next(times), lambda: rebuilds.append(1), 337.0) == [330.0]
    assert rebuilds == []
