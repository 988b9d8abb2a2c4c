#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python scripts/time_pack.py 512 3 float32 1
python scripts/time_pack.py 512 1 float64 1
python scripts/time_pack.py 512 2 float32 3
python scripts/time_jacobi.py 512 f64 10
python bench.py --steps 30 --warmup 5 > gpurun_out/bench5.json 2> gpurun_out/bench5.err; cat gpurun_out/bench5.json; tail -3 gpurun_out/bench5.err
echo "=== exchange latency: reference library vs ours, same driver source (oracle/ref/ref_exchange_uniform.cu) ==="
cd gpurun_out
for args in "512 512 512 3 2 30" "512 512 512 1 1 30" "128 128 128 1 2 30" "256 256 256 8 3 30"; do
  for how in default cudampi; do
    timeout 300 ../oracle/_ref/ref_exchange_uniform $args $how 2>/dev/null | tail -1
    timeout 300 ../bin/exchange_uniform $args $how 2>/dev/null | tail -1 | sed 's/ref_exchange/our_exchange/'
  done
done
timeout 300 ../bin/bench_pack 2>/dev/null | tail -3
timeout 300 ../bin/jacobi3d_strong 512 512 512 -n 30 2>/dev/null | tail -1
timeout 300 ../bin/exchange_weak 256 256 256 2>&1 | tail -3
cd ..
ncu --set full --clock-control none --import-source on -k regex:jacobi_march -s 4 -c 1 -f -o gpurun_out/prof_jacobi_r1_v2 python scripts/time_jacobi.py 512 f64 2 > gpurun_out/ncu_full2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1_v2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench2.log 2>&1
rm -f gpurun_out/plan_*.txt gpurun_out/mat_npy_loadtxt.txt
