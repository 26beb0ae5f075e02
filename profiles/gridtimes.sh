CMD="python bench.py --steps 1 --warmup 3 --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,launch__grid_size --clock-control none -k regex:"grid_tma_kernel|grid_stage_kernel|grid_kernel" -s 60 -c 20 --csv --log-file gpurun_out/grid_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['achieved'], d['kernels_ms_per_step'])"
