#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -5
for cfg in "" "SB_JACOBI_PREFETCH=0" "SB_JACOBI_PREFETCH=4" "SB_JACOBI_ZCHUNK=64" "SB_JACOBI_RY=1" "SB_JACOBI_RY=4"; do
  env $cfg python scripts/time_jacobi.py 512 f64 10 2>&1 | grep -E "interior|whole"
done
python scripts/time_jacobi.py 512 f32 10 2>&1 | grep -E "interior|whole|exterior"
ncu --metrics smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:jacobi_march -s 4 -c 1 python scripts/time_jacobi.py 512 f64 2 2>&1 | grep -E "smsp__|dram__|gpu__time|sm__warps"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench3.json 2> gpurun_out/bench3.err; cat gpurun_out/bench3.json; tail -3 gpurun_out/bench3.err
echo "=== reference drivers against OUR library (bin/) ==="
(cd gpurun_out && timeout 300 ../bin/test_cpu 2>&1 | tail -4)
(cd gpurun_out && timeout 600 ../bin/test_cuda 2>&1 | tail -6)
(cd gpurun_out && timeout 300 ../bin/jacobi3d 512 512 512 -n 30 2>/dev/null | tail -2)
(cd gpurun_out && timeout 300 ../bin/bench_exchange --x 512 --y 512 --z 512 --q 3 --fr 2 --er 2 --cr 2 2>/dev/null | tail -7)
(cd gpurun_out && timeout 300 ../bin/bench_pack 2>/dev/null | tail -4)
echo "=== the reference itself (oracle/_ref) ==="
(cd gpurun_out && timeout 600 ../oracle/_ref/ref_test_cuda 2>&1 | tail -4)
(cd gpurun_out && timeout 300 ../oracle/_ref/ref_jacobi3d 512 512 512 -n 30 2>/dev/null | tail -2)
(cd gpurun_out && timeout 300 ../oracle/_ref/ref_bench_exchange --x 512 --y 512 --z 512 --q 3 --fr 2 --er 2 --cr 2 2>/dev/null | tail -7)
(cd gpurun_out && timeout 300 ../oracle/_ref/ref_bench_pack 2>/dev/null | tail -4)
rm -f gpurun_out/plan_*.txt gpurun_out/mat_npy_loadtxt.txt
