python -m pytest tests/test_gpu_nms_binned.py -x -q 2>&1 | tail -15
python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -k "nms or cfg or postprocess or config" 2>&1 | tail -5
B="python bench.py --workload cfg3 --configs none --no-cpu-baseline --no-e2e --no-backbone --steps 20 --warmup 5"
for v in 2 1; do FDMS_NMS_BIN=$v $B 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('BIN=$v', j['value'], j['ms_per_step'], j['serial_ms_per_step'], j['stages_ms'])"; done
B="python bench.py --workload cfg3 --profile --no-graph --steps 2 --warmup 3"
ncu --set full --clock-control none --import-source on -k regex:"nms_binned" -s 8 -c 2 -o gpurun_out/prof_binned2 -f $B > gpurun_out/p_ncu.log 2>&1
