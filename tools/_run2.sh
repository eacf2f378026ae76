B="python bench.py --workload cfg3 --profile --no-graph --steps 2 --warmup 3"
$B > gpurun_out/p_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"nms_binned" -s 8 -c 2 -o gpurun_out/prof_binned -f $B > gpurun_out/p_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
