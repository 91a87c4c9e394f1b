"""bench.py's reference arm runs on the host alone (the oracle port of the reference's scorer), so its side of the
driver's contract is checked here on the CPU: one JSON line with the contract's keys, the SAME metric / unit / config
as the GPU arm prints (the driver compares them), bounded run time, and ranks other than 0 exit without work."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _run(extra_env=None, gpus=1):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.pop("WORLD_SIZE", None)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", str(gpus), "--steps", "2",
                           "--warmup", "1"], capture_output=True, text=True, timeout=300, env=env, cwd=str(ROOT))


@pytest.mark.timeout(360)
def test_reference_arm_prints_the_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the same metric, unit and config as the GPU arm's line (bench.py builds both from these)
    sys.path.insert(0, str(ROOT))
    import bench

    assert d["metric"] == bench.METRIC and d["config"] == bench.config_for(1) and d["data"] == bench.DATA
    assert "workload" in d["config"] and not any(k in d["config"] for k in ("model", "global_batch", "seq_len"))


@pytest.mark.timeout(360)
def test_reference_arm_scales_its_rows_with_n_and_only_rank_0_works():
    sys.path.insert(0, str(ROOT))
    import bench

    r = _run({"RANK": "1", "WORLD_SIZE": "2"}, gpus=2)
    assert r.returncode == 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    r = _run({"RANK": "0", "WORLD_SIZE": "2"}, gpus=2)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
    assert d["n_gpus"] == 2 and d["config"] == bench.config_for(2) and d["config"]["rows_total"] == 2 * bench.N_ROWS
