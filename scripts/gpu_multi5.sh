#!/bin/bash
# fused schedule: validation + timing.  usage: gpu_multi5.sh N
set -x
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_jacobi.py tests/test_gpu_exchange.py -q -m gpu -x 2>&1 | tail -4
pick() { python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$1', 'n', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'step_frac', round(d['roofline']['step_frac_of_roofline'],3), 'launches', d['gpu_launches'])"; }
for sch in fused queued host-sync; do
  timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --schedule $sch 2>gpurun_out/m5_err_1.log | tee gpurun_out/bench_m5_n1_$sch.json | pick $sch
done
for n in 2 4 8; do
  if [ $n -le $N ]; then
    for sch in fused queued; do
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --schedule $sch 2>gpurun_out/m5_err_$n.log | tee gpurun_out/bench_m5_n${n}_$sch.json | pick torchrun_$sch
      tail -2 gpurun_out/m5_err_$n.log | cut -c1-300
    done
  fi
done
timeout 300 python bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/m5_err_sp.log | tee gpurun_out/bench_m5_singleproc_n$N.json | pick singleproc_fused
tail -2 gpurun_out/m5_err_sp.log
timeout 300 python bench.py --steps 30 --warmup 5 2>/dev/null | tee gpurun_out/bench_m5_full_n1.json | cut -c1-400
