"""bench.py's driver contract: the reference arm runs on CPU (the oracle's restatement of best_multiexp on the host cores) and
prints ONE JSON line with the agreed keys; the product arm needs a GPU and carries the extra roofline / e2e / cpu_baseline
objects. Under torchrun only rank 0 of the reference arm prints."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
             "config", "e2e"}


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return [l for l in out.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_one_json_line_on_cpu():
    lines = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-max-log", "16"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    assert _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-max-log", "16"], env={"RANK": "1", "WORLD_SIZE": "2"}) == []


@pytest.mark.gpu
def test_product_arm_line_has_roofline_e2e_and_cpu_baseline():
    lines = _run(["--log-n", "20", "--ntt-log-n", "20", "--steps", "3", "--warmup", "3", "--prove-k", "12", "--prove-real-k", "10"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and "impl" not in d
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"]
    # frac is EXECUTED multiply-accumulates over the issue-limit peak: a real fraction; the pinned-algorithm reading sits beside it
    assert d["roofline"]["bound"] == "int" and 0 < d["roofline"]["frac"] <= 1.0
    assert d["roofline"]["vs_pinned_algorithm"]["mad32_per_point"] == 21760
    assert d["parity"]["point"]["x"].startswith("0x") and d["parity"]["paths_agree"] and d["e2e_pageable"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == (1 << 20) * 32 and d["e2e"]["d2h_bytes_per_step"] == 80 and d["e2e"]["value"] > 0
    assert d["gpu_launches"] > 0 and d["cpu_baseline"]["gpu_matches_cpu_on_sample"] is True
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert d["ntt"]["value"] > 0 and d["prove_ms"][0]["ms_per_proof"] > 0 and d["prove_ms_resident"][0]["ms_per_proof"] > 0
    assert d["prove_real_ms"][0]["quotient_identity_holds"] is True and d["prove_real_ms"][0]["ms_per_proof"] > 0


def test_executed_work_model_of_the_accumulation():
    """bench.py's MAD32-per-entry model: XYZZ alone at depth 0, the affine share growing with the tree depth"""
    sys.path.insert(0, ROOT) if ROOT not in sys.path else None
    import importlib

    bench = importlib.import_module("bench")
    assert bench.mad32_per_entry(0) == 1232 == 6 * 136 + 2 * 108 + 200
    assert abs(bench.MAD32_PER_AFFINE_ADD - (5 * 136 + 108 - 17 + 25.5)) < 1e-9
    prev = 1232.0
    for levels in range(1, 7):
        cur = bench.mad32_per_entry(levels)
        assert bench.MAD32_PER_AFFINE_ADD < cur < prev
        prev = cur
    assert abs(13 * bench.mad32_per_entry(4) - 10708.34375) < 1e-6
