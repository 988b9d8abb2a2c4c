#!/bin/bash
# usage: gpu_multi3.sh N   (N GPUs visible)
set -x
N=${1:-2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
timeout 300 bin/test_exchange_multigpu 2>&1 | tail -4
for a in "512 3 4 float32" "512 1 1 float64" "256 3 8 float64"; do
  timeout 200 python scripts/time_exchange_mg.py $a 2>&1 | tail -1
  SB_NO_STAGING=1 timeout 200 python scripts/time_exchange_mg.py $a 2>&1 | tail -1 | sed 's/^/[nostage] /'
done
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/m3_err_1.log | tee gpurun_out/bench_m3_n1.json
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 30 --warmup 5 --no-cpu-baseline 2>gpurun_out/m3_err_$n.log | tee gpurun_out/bench_m3_n$n.json
    tail -2 gpurun_out/m3_err_$n.log
  fi
done
SB_FORCE_NCCL=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/m3_err_nccl.log | tee gpurun_out/bench_m3_nccl_n$N.json
echo "=== astaroth drop-in (reference driver, our library vs reference library) ==="
( timeout 300 bin/astaroth 5 2>&1 | tail -6 )
( timeout 300 oracle/_ref/ref_astaroth 5 2>&1 | tail -6 )
cd gpurun_out
echo "=== reference vs ours on $N GPUs (1 process x N GPUs) ==="
for args in "512 512 512 3 2 30" "512 512 512 1 1 30" "256 256 256 8 3 30"; do
  timeout 300 ../oracle/_ref/ref_exchange_uniform $args default 2>/dev/null | tail -1
  timeout 300 ../bin/exchange_uniform $args default 2>/dev/null | tail -1 | sed 's/ref_exchange/our_exchange/'
done
rm -f plan_*.txt mat_npy_loadtxt.txt
