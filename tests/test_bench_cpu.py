"""CPU-side checks of bench.py: the reference arm runs the unmodified reference (baseline/_ref) through its own public call
and prints the contract's JSON line; the line's bookkeeping helpers are consistent."""

import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, cwd=ROOT, env=e,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout[-2000:]          # exactly ONE JSON line on stdout
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    from oracle import ref_shim
    line = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--frames", "16", "--no-config1"])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["higher_is_better"] is True
    assert line["metric"] == "mel frames/s of 50-step CFG reverse diffusion"
    assert line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 1 and line["value"] > 0
    cb = line["cpu_baseline"]
    assert cb["value"] == line["value"] and cb["cores"] == (os.cpu_count() or 1)
    assert cb["kind"] == ("reference" if ref_shim.installed_reference_available() else "port")
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["config"]["frames"] == 16


def test_reference_arm_other_ranks_exit_quietly():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, cwd=ROOT, env=dict(os.environ, RANK="1", WORLD_SIZE="2"), timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True,
                       text=True, cwd=ROOT, timeout=300)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
