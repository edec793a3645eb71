"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the keys the driver reads, and
the B200 arm refuses to run without CUDA instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample-videos", "1", "--frames", "12", "--boxes", "4")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "frames/s"
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d, k
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and "workload" in d["config"]


def test_b200_arm_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        return
    p = _run("--steps", "1", "--warmup", "0")
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)
