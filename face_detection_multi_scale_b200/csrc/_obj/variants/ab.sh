#!/bin/bash
# A/B of decode_filter builds: numerics (decode tests) then stage timings from bench.py
V=/root/repo/face_detection_multi_scale_b200/csrc/_obj/variants
for v in u4 u2 u1; do
  FDMS_NO_BUILD=1 FDMS_LIB=$V/libfdms_$v.so python -m pytest tests -m gpu -x -q -k "decode or postprocess_fuzz or e2e" 2>&1 | tail -1
done
for rep in 1 2; do
for v in base u1 u2 u4; do
  echo "== $v"
  FDMS_NO_BUILD=1 FDMS_LIB=$V/libfdms_$v.so python bench.py --steps 400 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print(j['value'], j['ms_per_step'], j['serial_ms_per_step'], j['stages_ms'])"
done; done
