mkdir -p gpurun_out
python -m pytest tests/test_gpu_head_module.py tests/test_gpu_reference_model.py -x -q > gpurun_out/b_tests.log 2>&1; tail -30 gpurun_out/b_tests.log
python - <<PY
import json
d=json.load(open("gpurun_out/ref_gpu_check.json")); r=d["report"]
print(d["checks"], d["failures"][:6])
print(json.dumps(r.get("P1h_half_model_decode_diffs"))[:1500])
print(json.dumps(r.get("C_half_mode"))[:1200])
print(json.dumps(r.get("D_detect_py"))[:800])
PY
