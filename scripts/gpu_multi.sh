#!/bin/bash
# multi-GPU visit: run with gpurun --gpus N
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi topo -m 2>/dev/null | head -12
timeout 900 python -m pytest tests -q -m gpu -x -k "multi_gpu or one_process_per_gpu" 2>&1 | tail -15
for n in 1 $N; do
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/mg_err_1.log | tee gpurun_out/bench_mg_n1.json
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $n --steps 20 --warmup 5 2>gpurun_out/mg_err_$n.log | tee gpurun_out/bench_mg_n$n.json
    tail -5 gpurun_out/mg_err_$n.log
    SB_FORCE_NCCL=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $n --steps 20 --warmup 5 --no-e2e 2>gpurun_out/mg_err_nccl_$n.log | tee gpurun_out/bench_mg_nccl_n$n.json
    tail -3 gpurun_out/mg_err_nccl_$n.log
    timeout 300 python bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/mg_err_sp_$n.log | tee gpurun_out/bench_mg_singleproc_n$n.json
    tail -3 gpurun_out/mg_err_sp_$n.log
  fi
done
cd gpurun_out
echo "=== reference vs ours on $N GPUs (1 process x N GPUs) ==="
timeout 300 ../oracle/_ref/ref_jacobi3d 512 512 512 -n 30 2>/dev/null | tail -1
timeout 300 ../bin/jacobi3d 512 512 512 -n 30 2>/dev/null | tail -1
for args in "512 512 512 3 2 30" "512 512 512 1 1 30"; do
  for how in default cudampi; do
    timeout 300 ../oracle/_ref/ref_exchange_uniform $args $how 2>/dev/null | tail -1
    timeout 300 ../bin/exchange_uniform $args $how 2>/dev/null | tail -1 | sed 's/ref_exchange/our_exchange/'
  done
done
rm -f plan_*.txt mat_npy_loadtxt.txt
