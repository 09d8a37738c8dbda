"""bench.py's reference arm runs the CPU oracle only, so its side of the JSON contract can be checked without a GPU: one line on
stdout, the keys the driver reads, the same `metric` / `config.workload` our arm prints for that --config, a bounded sample."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
            "data", "config", "cpu_baseline", "e2e", "gpu_launches")


@pytest.mark.parametrize("config,cpu_games", [(0, 8), (1, 8), (2, 4), (4, 4)])
def test_reference_arm_prints_one_contract_line(config, cpu_games):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", str(config), "--steps", "1", "--warmup", "1",
                          "--cpu-games", str(cpu_games)], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["gpu_launches"] == 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["unit"] == "explores/s" and d["steps"] == 1 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "configs[%d]" % config in d["config"]["workload"]
    want_explores = {0: 800, 1: 800, 2: 1600}.get(config)
    if want_explores:
        assert str(want_explores) in d["metric"] and d["config"]["explores_per_move"] == want_explores
    else:
        assert "evaluation" in d["metric"]
