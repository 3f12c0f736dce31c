"""bench.py prints ONE JSON line with the keys the driver reads (both arms), on a small batch."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_line():
    """--impl reference: the CPU arm (oracle port on the host cores), no GPU needed."""
    d = _run(["--impl", "reference", "--steps", "5", "--warmup", "3", "--envs-per-gpu", "1024", "--long-steps", "50"])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None
    assert d["config"]["envs_per_gpu"] == 1024 and d["cpu_baseline"]["long_window"]["value"] > 0


@pytest.mark.gpu
def test_b200_arm_line(gpu):
    d = _run(["--steps", "30", "--warmup", "5", "--envs-per-gpu", "4096", "--long-steps", "200"])
    assert BASE_KEYS | {"clocks", "gpu_launches", "roofline"} <= set(d)
    assert d["metric"] == "env_steps_per_sec" and d["n_gpus"] == 1 and d["steps"] == 30 and d["scaling"] == "weak"
    assert d["dtype"] == "u8" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["gpu_launches"] >= 30 * 5
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 4096 and 0 < e["d2h_bytes_per_step"] < 4096 * 1969
    assert d["e2e_full_copy"]["d2h_bytes_per_step"] == 4096 * 1969
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["value"] > 0 and c["cores"] >= 1 and c["sample"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert "workload" in d["config"] and "cache" in d["config"]
    # the measured window agrees with the oracle, and so does every extra workload
    assert d["oracle_digest_match"] is True and d["oracle_check_rank0"]["live"] > 1000
    x = d["extra"]
    assert x["mixed_seeds"]["value"] > 0 and x["mixed_seeds"]["oracle_check_rank0"]["match"] is True
    assert x["mini_4096"]["value"] > 0 and x["mini_4096"]["oracle_check_rank0"]["match"] is True
    assert set(x["reset_sweep"]) == {"32x16", "80x24", "160x48"}
    assert all(v["value"] > 0 and v["oracle_check_rank0"]["match"] for v in x["reset_sweep"].values())
