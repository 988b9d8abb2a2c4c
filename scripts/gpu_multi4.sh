#!/bin/bash
# step_async validation + timing.  usage: gpu_multi4.sh N
set -x
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_jacobi.py tests/test_gpu_exchange.py -q -m gpu -x 2>&1 | tail -4
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/m4_err_1.log | tee gpurun_out/bench_m4_n1.json | cut -c1-330
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --host-sync 2>>gpurun_out/m4_err_1.log | tee gpurun_out/bench_m4_n1_hostsync.json | cut -c1-330
for n in 2 4 8; do
  if [ $n -le $N ]; then
    for mode in "" "--host-sync"; do
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 30 --warmup 5 --no-cpu-baseline --no-e2e $mode 2>gpurun_out/m4_err_$n.log | tee gpurun_out/bench_m4_n$n$mode.json | cut -c1-330
      tail -2 gpurun_out/m4_err_$n.log | cut -c1-300
    done
  fi
done
timeout 300 python bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/m4_err_sp.log | tee gpurun_out/bench_m4_singleproc_n$N.json | cut -c1-330
tail -2 gpurun_out/m4_err_sp.log
echo "=== astaroth (reference driver) on ONE GPU: reference library, then ours ==="
( CUDA_VISIBLE_DEVICES=0 timeout 300 oracle/_ref/ref_astaroth 5 2>&1 | tail -3 )
( CUDA_VISIBLE_DEVICES=0 timeout 300 bin/astaroth 5 2>&1 | tail -3 )
