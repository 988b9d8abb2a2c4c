#!/bin/bash
set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_march_kernel -s 8 -c 1 -o gpurun_out/prof_jacobi_push -f python bench.py --steps 5 --warmup 5 --no-cpu-baseline --no-e2e --schedule fused > gpurun_out/ncu_push.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_march_kernel -s 8 -c 1 -o gpurun_out/prof_jacobi_whole -f python bench.py --steps 5 --warmup 5 --no-cpu-baseline --no-e2e --no-overlap > gpurun_out/ncu_whole.log 2>&1
tail -2 gpurun_out/ncu_push.log
