#!/bin/bash
set -x
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
python scripts/time_jacobi.py 512 f32 10 2>&1 | grep -E "interior|whole"
SB_JACOBI_SHIFT=0 python scripts/time_jacobi.py 512 f32 10 2>&1 | grep -E "interior|whole"
SB_JACOBI_MB=5 python scripts/time_jacobi.py 512 f32 10 2>&1 | grep -E "interior|whole"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/mg_err_1.log | tee gpurun_out/bench_mg_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e 2>gpurun_out/mg_err_$N.log | tee gpurun_out/bench_mg_n$N.json
tail -3 gpurun_out/mg_err_$N.log
SB_NO_STAGING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e 2>gpurun_out/mg_err_nostage_$N.log | tee gpurun_out/bench_mg_nostage_n$N.json
timeout 300 python bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/mg_err_sp_$N.log | tee gpurun_out/bench_mg_singleproc_n$N.json
tail -3 gpurun_out/mg_err_sp_$N.log
cd gpurun_out
echo "=== reference vs ours on $N GPUs (1 process x N GPUs) ==="
for args in "512 512 512 3 2 30" "512 512 512 1 1 30"; do
  for how in default; do
    timeout 300 ../oracle/_ref/ref_exchange_uniform $args $how 2>/dev/null | tail -1
    timeout 300 ../bin/exchange_uniform $args $how 2>/dev/null | tail -1 | sed 's/ref_exchange/our_exchange/'
  done
done
timeout 600 ../bin/test_cuda 2>&1 | tail -3
rm -f plan_*.txt mat_npy_loadtxt.txt
