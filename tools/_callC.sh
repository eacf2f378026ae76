mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/c_tests.log 2>&1; tail -6 gpurun_out/c_tests.log
(time python bench.py --steps 20 --warmup 5) > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; tail -4 gpurun_out/c_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/c_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['serial_ms_per_step'], d['stages_ms'], d['roofline']['frac'], d['roofline']['isolated']['frac'], d['e2e']['value'])
print({k:(v['value'], v['stages_ms']) for k,v in d['configs'].items()})
print(json.dumps(d['backbone'])[:700])
PY
