"""CPU: the benchmark drivers run end to end through the C-ABI test double (tools/dryrun_bench.py) and print ONE JSON line that
carries every key of the measurement contract.  Guards the driver logic of scripts that otherwise only run on the GPU box; the
numbers are meaningless here and are not looked at."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONTRACT = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
            "config", "clocks", "e2e", "gpu_launches", "roofline")


def _run(which):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dryrun_bench.py"), which], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.strip().split("\n") if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


@pytest.mark.parametrize("which", ["decode", "lora"])
def test_bench_driver_prints_the_contract_line(which):
    d = _run(which)
    for k in CONTRACT:
        assert k in d, k
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert "workload" in d["config"]
    if which == "decode":
        assert d["metric"] == "decode_tokens_per_s" and d["scaling"] == "strong"
        assert set(d["attention"]) >= {"decode", "prefill"} and "error" not in d["attention"]
        assert d["ts_encoder"]["patch_rows"] > 0
        assert d["config"]["variants"] == {"decode_fused": 0, "peer_ll": 0, "native_step": 0, "decode_chain": 0}
    else:
        assert d["metric"] == "lora_finetune_positions_per_s" and d["scaling"] == "weak"
        assert 11.0 < d["loss"] < 13.0          # ln(vocab) at initialisation (LoRA B = 0)
