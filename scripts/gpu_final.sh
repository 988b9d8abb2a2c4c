#!/bin/bash
# final single-GPU evidence run for profiles/ (round 1)
set -x
mkdir -p gpurun_out/final
F=gpurun_out/final
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -5 > $F/pytest_gpu.txt; cat $F/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 > $F/smoke.txt; cat $F/smoke.txt
timeout 600 python bench.py --steps 30 --warmup 5 2>$F/bench_err.log > $F/bench_n1.json; cut -c1-600 $F/bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null > $F/bench_reference.json; cut -c1-300 $F/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $F/launches_n1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $F/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_march_kernel -s 4 -c 1 -o $F/prof_jacobi_fused -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $F/ncu_full.log 2>&1
( timeout 600 bin/test_cuda 2>&1 | tail -3 ) > $F/test_cuda.txt; cat $F/test_cuda.txt
( timeout 300 bin/test_cpu 2>&1 | tail -3 ) > $F/test_cpu.txt; cat $F/test_cpu.txt
SKIP_CELL=1 timeout 300 python scripts/time_astaroth.py 256 f64 5 2>&1 | tail -5 > $F/astaroth_f64.txt; cat $F/astaroth_f64.txt
SKIP_CELL=1 timeout 300 python scripts/time_astaroth.py 256 f32 5 2>&1 | tail -5 > $F/astaroth_f32.txt; cat $F/astaroth_f32.txt
( timeout 300 bin/astaroth_b200 5 2>&1 | tail -1 ) > $F/astaroth_b200.txt; cat $F/astaroth_b200.txt
( cd gpurun_out && timeout 300 ../bin/jacobi3d 512 512 512 -n 30 2>&1 | tail -2 ) > $F/ref_driver_our_lib.txt; cat $F/ref_driver_our_lib.txt
rm -f gpurun_out/plan_*.txt gpurun_out/mat_npy_loadtxt.txt plan_*.txt
