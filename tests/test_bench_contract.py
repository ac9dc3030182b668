"""CPU: the parts of bench.py's contract that do not need a GPU -- the reference arm (the CPU port of the path timed on
the host cores) prints exactly one JSON line with the agreed keys, only rank 0 runs it under a multi-rank launch, and the
B200 arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CMD = [sys.executable, os.path.join(ROOT, "bench.py")]
FAST = ["--workload", "cfg1", "--steps", "1", "--warmup", "0", "--cpu-sample", "1"]


def _run(extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(CMD + extra, capture_output=True, text=True, env=e, cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = _run(["--impl", "reference"] + FAST)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "stdout must carry the JSON line only"
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "trials/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None


def test_reference_arm_runs_on_rank_zero_only():
    r = _run(["--impl", "reference", "--gpus", "2"] + FAST, env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return                                   # on the GPU box the arm simply runs (covered by the driver)
    r = _run(["--steps", "1", "--warmup", "0"])
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)
