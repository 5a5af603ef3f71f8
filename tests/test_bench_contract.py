"""CPU tier: the driver-facing contract of bench.py that can be checked without a GPU -- the reference arm
(`--impl reference`, the oracle port on the host cores) prints exactly ONE JSON line on stdout carrying the keys the
contract names, and the CUDA arm refuses to run without a device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600,
                          env={**os.environ, **(env or {})})


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    res = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--height", "96", "--width", "128", "--cpu-images", "2")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, f"stdout must hold exactly one line, got {len(lines)}"
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "Mpixel/s" and j["higher_is_better"] is True
    assert j["metric"].startswith("Mpixel/s DML head+OOD score")
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e", "gpu_launches"):
        assert key in j, key
    assert j["vs_baseline"] is None and j["data"] == "synthetic" and j["gpu_launches"] == 0
    cb = j["cpu_baseline"]
    # "reference": the unmodified reference copy under baseline/_ref was executed; "port": the oracle (copy absent)
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == j["value"] and "sample" in cb
    import os
    if os.path.isdir(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "anomaly")):
        assert cb["kind"] == "reference" and cb.get("reference_load_error") is None
    assert j["e2e"] == {"value": j["value"], "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and "model" not in j["config"]
    assert j["value"] > 0 and j["ms_per_step"] > 0


def test_reference_arm_other_ranks_exit_silently():
    res = _run("--impl", "reference", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_cuda_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is visible")
    res = _run("--steps", "1", "--warmup", "0", "--images", "1", "--no-e2e", "--no-cpu-baseline")
    assert res.returncode != 0
    assert res.stdout.strip() == ""                      # no JSON line from a run that measured nothing
    assert "CUDA" in res.stderr or "cuda" in res.stderr
