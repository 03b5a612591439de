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


@pytest.mark.parametrize("name", ["r2_bench_driver_style.json", "r2_bench_head_final.json", "r2_bench_8gpu.json", "r2_bench_2gpu.json"])
def test_kept_bench_lines_carry_the_contract_keys(name):
    """The bench lines kept under profiles/ (the ones README / DESIGN quote) are complete contract lines: metric and config
    of BASELINE.json, e2e with its byte counts, roofline against MEASURED_PEAKS-style denominators, clocks without a
    thermal / hardware slowdown, a non-zero launch count."""
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(name + " not kept")
    with open(path) as f:
        line = json.load(f)
    assert line["metric"] == "mel frames/s of 50-step CFG reverse diffusion" and line["unit"] == "frames/s"
    assert line["higher_is_better"] is True and line["scaling"] == "weak" and line["data"] == "synthetic" and line["dtype"] == "f16"
    assert line["n_gpus"] >= 1 and line["steps"] >= 1 and line["warmup"] >= 3 and line["value"] > 0 and line["ms_per_step"] > 0
    cfg = line["config"]
    assert "workload" in cfg and cfg["frames"] == 1000 and cfg["batch_per_gpu"] == 32 and cfg["diffusion_steps"] == 50
    assert cfg["job_utterances"] == 32 * line["n_gpus"]
    # value = frames of the whole job / time of one pass
    assert line["value"] == pytest.approx(cfg["job_utterances"] * cfg["frames"] / (line["ms_per_step"] / 1e3), rel=1e-6)
    e2e = line["e2e"]
    assert e2e["unit"] == "frames/s" and 0 < e2e["value"] <= line["value"] * 1.02
    assert e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0
    assert line["gpu_launches"] > 0
    roof = line["roofline"]
    assert roof["bound"] == "tensor" and roof["unit"] == "TFLOP/s" and roof["frac"] == pytest.approx(roof["achieved"] / roof["peak"], rel=1e-6)
    assert 0.5 < roof["frac"] < 1.0
    clocks = line["clocks"]
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(clocks["reasons"])
    assert line["vs_baseline"] is None      # BASELINE.md holds no published number for this metric
