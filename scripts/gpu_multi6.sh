#!/bin/bash
# lean multi-GPU run: validation + scaling.  usage: gpu_multi6.sh N
set -x
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_exchange.py tests/test_gpu_jacobi.py -q -m gpu -x -k "multi_gpu or one_process or step_async" 2>&1 | tail -3
pick() { grep "^{" | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$1', 'n', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'per_gpu', '%.4g' % d['per_gpu'], 'launches', d['gpu_launches'])"; }
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --schedule fused 2>/dev/null | tee gpurun_out/bench_m6_n1_fused.json | pick n1_fused
for n in 2 4 8; do
  if [ $n -le $N ]; then
    for sch in fused queued; do
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800+n)) bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --schedule $sch 2>gpurun_out/m6_err_${n}_$sch.log | tee gpurun_out/bench_m6_n${n}_$sch.json | pick torchrun_$sch
    done
  fi
done
timeout 300 python bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --schedule fused 2>gpurun_out/m6_err_sp.log | tee gpurun_out/bench_m6_singleproc_n$N.json | pick singleproc_fused
timeout 200 bin/test_exchange_multigpu 2>&1 | tail -2
cd gpurun_out
for args in "512 512 512 3 2 30" "512 512 512 1 1 30"; do
  timeout 200 ../oracle/_ref/ref_exchange_uniform $args default 2>/dev/null | tail -1
  timeout 200 ../bin/exchange_uniform $args default 2>/dev/null | tail -1 | sed 's/ref_exchange/our_exchange/'
done
rm -f plan_*.txt mat_npy_loadtxt.txt
