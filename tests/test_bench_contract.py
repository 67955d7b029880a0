"""bench.py contract checks that need no GPU: the reference arm (the C restatement of the reference's serial CPU matvec,
`--impl reference`) prints ONE JSON line with the keys the driver reads; our own arm refuses to run without a device
(no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=300, cwd=ROOT)


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--levels", "30", "--steps", "4", "--warmup", "3")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fsp_matvec_hbm_gbs" and d["unit"] == "GB/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 4 and d["warmup"] >= 3
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and d["config"]["states"] == 5456          # C(33, 3): simplex i + j + k <= 30
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] > 0 and "sample" in cb
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    # the same bytes-per-matvec figure both arms divide by: SURVEY 8(d) B_mv on the reference's stored structure
    n, R = 5456, 6
    assert d["config"]["algorithmic_bytes_per_step"] > 16 * (n + R)
    assert abs(d["value"] - d["config"]["algorithmic_bytes_per_step"] / (d["ms_per_step"] * 1e-3) / 1e9) <= 1e-6 * d["value"]


def test_own_arm_needs_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = _run("--levels", "20", "--steps", "2", "--warmup", "3", "--no-cpu", "--no-solve")
    assert r.returncode != 0, "bench.py must not produce a number without a GPU (no CPU fallback)"
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
