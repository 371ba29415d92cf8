#!/bin/bash
# one steady-state frame under ncu --set full (all five kernels), plus the cheap launch list; results in gpurun_out/
TAG=${1:-r1}
ncu --set full --clock-control none --import-source on -k regex:"drawcull_kernel|pyramid_kernel|clustercull_kernel" -s 20 -c 5 -f -o gpurun_out/prof_frame_$TAG python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 40 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
