"""CPU: bench.py's reference arm (`--impl reference`) prints the contract's JSON line without a GPU, and the GPU arm refuses to run on
a CPU box instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--no-full-forward"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "clips/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("grounding clips/sec") and line["value"] > 0 and line["steps"] == 1
    # the reference's own modules where they are importable (source tree or the byte-compiled oracle/_ref), else the numpy port
    from ref_loader import reference_modules_available
    assert line["cpu_baseline"]["kind"] == ("reference" if reference_modules_available() else "port")
    assert line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_full_reference_forward_tool():
    """BASELINE configs[0]: the reference's full VSTGNet.forward runs here through the stand-ins of tools/full_forward_cpu.py
    (a short clip keeps the CPU suite fast: 4 frames at 64 px)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from ref_loader import reference_modules_available
    if not reference_modules_available():
        import pytest
        pytest.skip("neither /root/reference nor oracle/_ref is present")
    out = subprocess.run([sys.executable, "-c",
                          "import sys, json; sys.path.insert(0, 'tools'); import full_forward_cpu as F; "
                          "print(json.dumps(F.full_forward_seconds(frames=4, res=64)))"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["pred_boxes_shape"] == [4, 4] and r["seconds_per_forward"] > 0 and "vid" in r["parts_seconds_last_call"]


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
