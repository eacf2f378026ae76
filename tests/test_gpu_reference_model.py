"""The drop-in against the unmodified reference running a real reference network on the GPU (the checks
are in tests/ref_gpu_check.py, started as a subprocess so the reference's `utils` / `models` packages stay out of this
process).  Needs the reference tree: `baseline/_ref/` on the GPU box (baseline/install_ref.py), `/root/reference` here."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_reference_model_on_gpu_class_swap_monkey_patch_detect_py():
    from baseline import refshim
    if refshim.find_root() is None:
        pytest.skip("no reference tree (baseline/_ref not installed)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_gpu_check.py")], capture_output=True, text=True,
                       timeout=900, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, "ref_gpu_check.py printed no result:\n%s\n%s" % (r.stdout[-2000:], r.stderr[-4000:])
    res = json.loads(lines[-1])
    if res.get("skipped"):
        pytest.skip(res["skipped"])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ref_gpu_check.json"), "w") as f:
        json.dump(res, f, indent=1)
    assert res["checks"] >= 10 and not res["failures"], res
