"""bench.py's N > 1 arm executed end to end on CPU with stand-ins for the GPU pieces (Worker,
CUDA events / streams, the process group): every line of its host logic -- equal-cuts step,
calibration, settle, warm-up, timed steps, parity, per-rank tables, the JSON line -- runs, so a
typo cannot first show up on the 8-GPU box.  Numbers are made up; the real arm is measured on
the GPU box (profiles/r02_bench_n*_world_*.json)."""
import importlib.util
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeWorker:
    made = 0

    def __init__(self, cfg, *, device=0, cuts=None, **opts):
        self.cfg, self.cuts, self.rank, self.world_size, self.group = cfg, cuts, 0, 4, None
        self.r = types.SimpleNamespace(set_option=lambda k, v: None, stream_ptr=0)
        self.spins = 0
        FakeWorker.made += 1

    def spin(self, n=None, seed=5127801):
        self.spins += 1
        slow = 1.2 if FakeWorker.made in (4,) else 1.0      # one "slow" world, rebuilt by settle
        return {"events": 600 * n // 4, "scatters": n // 8, "n_left": n // 10, "n_right": n // 9, "n_dead": 0,
                "births": n // 4, "sent_left": n // 2, "sent_right": n // 3, "window_crossings": 0,
                "idle_polls": 5, "blocked_passes": 0, "bank_pushes": 0, "bank_pops": 0,
                "lane_slots": 700 * n // 4, "idle_warp_ns": 1000, "w_left": 0.2, "w_right": 0.4, "w_dead": 0.0,
                "kernel_ms": 300.0 * slow, "windows": 1, "ctas": 592, "block": 256, "stripes": 4736,
                "ring_cap": 4096, "error": 0}

    def recut(self, cuts):
        FakeWorker.made += 1
        self.cuts = list(cuts) if cuts is not None else None

    def all_ranks(self, values, op="sum"):
        v = [float(x) for x in values]
        if op == "table":
            return [[x * (1.0 + 0.05 * r) for x in v] for r in range(self.world_size)]
        return v if op == "max" else [x * self.world_size for x in v]

    def parity(self, case, digest):
        return {"checked": True, "case": case, "tally_bit_exact": True, "counts_exact": True,
                "conservation": 1.0, "conservation_ok": True, "kernel_error": 0, "events": 1,
                "migrations_per_history": 4.0, "ranks": self.world_size, "cuts": self.cuts or "equal"}

    def gather_weights_absorbed(self, exact=False):
        return np.ones(self.cfg.nb_cells)

    def close(self):
        pass


def test_world_arm_host_logic_runs_end_to_end(monkeypatch, capsys):
    import torch
    import torch.distributed as dist
    spec = importlib.util.spec_from_file_location("bench_dry", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from mc_mpi_b200 import worker

    class Ev:
        def __init__(self, enable_timing=False):
            pass

        def record(self, stream=None):
            pass

        def elapsed_time(self, other):
            return 900.0

    monkeypatch.setattr(worker, "Worker", FakeWorker)
    monkeypatch.setattr(dist, "init_process_group", lambda *a, **k: None)
    monkeypatch.setattr(dist, "barrier", lambda *a, **k: None)
    monkeypatch.setattr(dist, "destroy_process_group", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "ExternalStream", lambda *a, **k: object())
    monkeypatch.setattr(torch.cuda, "Event", Ev)
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)
    monkeypatch.setattr(bench.ClockSampler, "stop", lambda self: {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": []})
    FakeWorker.made = 0
    args = types.SimpleNamespace(particles=None, retire_batch=0, inflight=0, rng="lcg", balance=True,
                                 calibrations=4, seg_cost=25.0, warmup=3, steps=3, no_e2e=False,
                                 verbose=True, gpus=4)
    bench.run_world_arm(args, 4, 0, 0)
    out = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(out) == 1
    d = json.loads(out[0])
    assert d["n_gpus"] == 4 and d["steps"] == 3 and d["scaling"] == "weak" and d["higher_is_better"]
    assert d["value"] == 4 * 125_000_000 * 3 / 0.9 and d["ms_per_step"] == 300.0
    assert d["parity"]["tally_bit_exact"] and d["config"]["equal_cuts"]["parity"]["tally_bit_exact"]
    assert d["config"]["equal_cuts_value"] > 0 and d["e2e"]["value"] > 0 and d["gpu_launches"] == 12
    assert d["roofline"]["kernel"] == "world_kernel" and d["roofline"]["frac"] > 0
    w = d["world"]
    assert len(w["lane_occupancy_per_rank"]) == 4 and w["nvlink"]["bytes_per_step"] > 0
    settle = [c for c in w["calibration"] if "settle_step_ms" in c]
    assert len(settle) == 1 and settle[0]["settle_step_ms"][-1] <= 1.05 * settle[0]["best_candidate_ms"]

    # Philox mode: parity is reported as not checked, nothing else changes
    args.rng = "philox"
    FakeWorker.made = 0
    bench.run_world_arm(args, 4, 0, 0)
    d = json.loads([l for l in capsys.readouterr().out.splitlines() if l.startswith("{")][0])
    assert d["parity"]["checked"] is False and d["config"]["rng"] == "philox"
