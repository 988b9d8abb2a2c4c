#!/bin/bash
set -x
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_jacobi.py tests/test_gpu_exchange.py -q -m gpu -x 2>&1 | tail -5
pick() { grep "^{" | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$1', 'n', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'per_gpu', '%.4g' % d['per_gpu'], 'launches', d['gpu_launches'], d['clocks']['sm_mhz'])"; }
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --schedule fused 2>/dev/null | tee gpurun_out/bench_m8_n1_fused.json | pick n1_fused
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29820+n)) bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --schedule fused 2>gpurun_out/m8_err_$n.log | tee gpurun_out/bench_m8_n${n}_fused.json | pick torchrun_fused
    tail -3 gpurun_out/m8_err_$n.log | cut -c1-200
  fi
done
SB_FUSED_DENSE_X=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29831 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --schedule fused 2>/dev/null | pick torchrun_fused_direct
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29832 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --schedule queued 2>/dev/null | pick torchrun_queued
timeout 300 python bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --schedule fused 2>gpurun_out/m8_err_sp.log | tee gpurun_out/bench_m8_singleproc_n$N.json | pick singleproc_fused
tail -2 gpurun_out/m8_err_sp.log | cut -c1-200
