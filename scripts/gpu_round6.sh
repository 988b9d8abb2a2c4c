#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -3
for cfg in "" "SB_JACOBI_MB=5" "SB_JACOBI_MB=6" "SB_JACOBI_RY=2 SB_JACOBI_MB=4" "SB_JACOBI_RY=2 SB_JACOBI_MB=3" "SB_JACOBI_MB=5 SB_JACOBI_PREFETCH=3" "SB_JACOBI_MB=5 SB_JACOBI_PREFETCH=1" "SB_JACOBI_MB=5 SB_JACOBI_ZCHUNK=64"; do
  env $cfg python scripts/time_jacobi.py 512 f64 10 2>&1 | grep -E "interior"
done
for cfg in "" "SB_JACOBI_MB=5" "SB_JACOBI_RY=2 SB_JACOBI_MB=3"; do
  env $cfg python scripts/time_jacobi.py 512 f32 10 2>&1 | grep -E "interior"
done
python scripts/time_pack.py 512 3 float32 1
SB_LIB_PATH=$PWD/stencil_b200/_alt/libstencil_b200_ld1.so python scripts/time_pack.py 512 3 float32 1
SB_LIB_PATH=$PWD/stencil_b200/_alt/libstencil_b200_ld2.so python scripts/time_pack.py 512 3 float32 1
python scripts/time_pack.py 512 1 float64 1
python scripts/time_pack.py 512 2 float32 3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err; cat gpurun_out/bench4.json; tail -3 gpurun_out/bench4.err
SB_JACOBI_MB=5 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null
(cd gpurun_out && timeout 300 ../oracle/_ref/ref_bench_exchange --x 512 --y 512 --z 512 --q 3 --fr 2 --er 2 --cr 2 2>&1 | grep -v cudaDeviceProp | tail -25)
ncu --set full --clock-control none --import-source on -k regex:box_copy -c 2 -f -o gpurun_out/prof_pack_r1 python scripts/time_pack.py 512 3 float32 1 > gpurun_out/ncu_pack.log 2>&1
rm -f gpurun_out/plan_*.txt gpurun_out/mat_npy_loadtxt.txt
