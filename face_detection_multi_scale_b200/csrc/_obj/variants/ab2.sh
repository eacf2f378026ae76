#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
FDMS_MERGE_FUSED=0 python -m pytest tests -m gpu -x -q -k "e2e or postprocess or merge or weighted or api" 2>&1 | tail -1
for rep in 1 2; do
for f in 0 1; do
  echo "== fused=$f"
  FDMS_MERGE_FUSED=$f python bench.py --steps 400 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print(j['value'], j['ms_per_step'], j['serial_ms_per_step'], j['stages_ms'], j['gpu_launches'])"
done; done
echo "== cfg3 fused=0/1"
for f in 0 1; do
FDMS_MERGE_FUSED=$f python bench.py --workload cfg3 --steps 30 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print(j['value'], j['ms_per_step'], j['serial_ms_per_step'], j['stages_ms'], j['gpu_launches'])"
done
