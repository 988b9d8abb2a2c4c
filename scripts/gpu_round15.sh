#!/bin/bash
set -x
mkdir -p gpurun_out
export SKIP_CELL=1 SKIP_ITER=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ac_tile_kernel -s 3 -c 2 -o gpurun_out/prof_ac_tile_f64 -f python scripts/time_astaroth.py 256 f64 1 > gpurun_out/ncu_ac1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ac_tile_kernel -s 3 -c 1 -o gpurun_out/prof_ac_tile_f32 -f python scripts/time_astaroth.py 256 f32 1 > gpurun_out/ncu_ac2.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:solve -s 4 -c 1 -o gpurun_out/prof_ac_ref -f oracle/_ref/ref_astaroth_solve 256 1e-8 - - > gpurun_out/ncu_ac3.log 2>&1
tail -3 gpurun_out/ncu_ac1.log gpurun_out/ncu_ac3.log
