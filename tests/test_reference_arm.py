"""The reference arm of bench.py (baseline/ref_runner.py, oracle/ref_port.py) on CPU: the harness around the UNMODIFIED
reference and the library-call port both reproduce the oracle's results on the bench's kind of workload, and the
all-cores pool runs.  The reference is imported in child processes only (its `utils` / `models` packages stay out of
pytest)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import json, sys
sys.path.insert(0, %(root)r)
import numpy as np
from workloads import standin, synth
from oracle import oracle as O
from baseline import ref_runner as R
mode = sys.argv[1]
h, w, sizes, B = 270, 480, [160, 256, 320], 3
wl = dict(h=h, w=w, sizes=sizes, faces=25, conf=0.5, iou=0.5, radius=1)
images, heads, shapes = standin.make_batch(h, w, sizes, B, 25, seed=5)
out = {"cases": 0, "bad": []}
if mode == "harness":
    if not R.available():
        print(json.dumps({"skipped": "no reference tree"})); sys.exit(0)
    hs = R.RefHarness(sizes, 0.5, 0.5, synth.P5_STRIDES, synth.P5_ANCHORS)
    run = lambda b: hs.detect(images[b], R.maps_of(heads, shapes, b))[0]
else:
    from oracle import ref_port
    import torch
    torch.set_grad_enabled(False)
    run = lambda b: np.asarray(ref_port.detect_image(images[b], R.maps_of(heads, shapes, b), sizes, synth.P5_STRIDES, synth.P5_ANCHORS, 0.5, 0.5), np.float64)
for b in range(B):
    got = run(b)
    ref, _ = O.detect_from_heads([[l[b:b + 1] for l in heads[s]] for s in range(len(sizes))], shapes, (h, w, 3),
                                 synth.P5_STRIDES, synth.P5_ANCHORS, 0.5, 0.5, O.TV_CPU)
    out["cases"] += 1
    ok = got.shape == ref.shape and len(ref) > 0 and np.array_equal(got[:, [0, 1, 2, 3, 5, 6]], ref[:, [0, 1, 2, 3, 5, 6]]) \
        and float(np.max(np.abs(got[:, 4] - ref[:, 4]))) <= 2e-6  # ATen's vectorised expf vs the correctly rounded one
    if not ok:
        out["bad"].append([b, list(got.shape), list(ref.shape)])
if mode == "harness":
    pool = R.RefPool(wl, synth.P5_STRIDES, synth.P5_ANCHORS, workers=2, per_worker=2)
    try:
        secs, n = pool.step()
        out["pool"] = [pool.images_per_step, n > 0, secs > 0]
    finally:
        pool.close()
print(json.dumps(out))
'''


def _run(mode):
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}, mode], capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, (r.stdout[-1000:], r.stderr[-3000:])
    return json.loads(lines[-1])


def test_library_call_port_matches_oracle():
    res = _run("port")
    assert res["cases"] == 3 and not res["bad"], res


def test_unmodified_reference_harness_matches_oracle_and_pool_runs():
    res = _run("harness")
    if res.get("skipped"):
        pytest.skip(res["skipped"])
    assert res["cases"] == 3 and not res["bad"], res
    assert res["pool"] == [4, True, True], res


def test_bench_config_is_identical_in_both_arms():
    sys.path.insert(0, ROOT)
    import bench
    for name, wl in bench.WORKLOADS.items():
        a, b = bench.make_config(wl, 1), bench.make_config(dict(wl), 1)
        assert a == b and a["workload"].startswith(name)
    assert np.all([k in bench.WORKLOADS for k in ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5")])
