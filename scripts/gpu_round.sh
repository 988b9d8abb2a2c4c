#!/bin/bash
# One GPU visit: parity tests, kernel timings, bench, ncu launch list + full capture of the top kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -8
for cfg in "" "SB_JACOBI_PD=2" "SB_JACOBI_PD=2 SB_JACOBI_PREFETCH=0" "SB_JACOBI_PD=2 SB_JACOBI_PREFETCH=8" "SB_JACOBI_ZCHUNK=16" "SB_JACOBI_ZCHUNK=64" "SB_JACOBI_PREFETCH=2" "SB_JACOBI_PREFETCH=6"; do
  env $cfg python scripts/time_jacobi.py 512 f64 10 2>&1 | grep -E "interior|exterior|exchange"
done
python bench.py --steps 20 --warmup 5 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; cat gpurun_out/bench2.json; tail -5 gpurun_out/bench2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jacobi_march -s 4 -c 1 -f -o gpurun_out/prof_jacobi_r1 python scripts/time_jacobi.py 512 f64 2 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/
