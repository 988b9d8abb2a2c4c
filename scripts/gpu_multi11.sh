#!/bin/bash
set -x
N=${1:-8}
pick() { grep "^{" | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$1', 'n', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'per_gpu', '%.4g' % d['per_gpu'], 'launches', d['gpu_launches'], d['config']['schedule'], d['clocks']['sm_mhz'])"; }
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29871 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/m11_err.log | tee gpurun_out/bench_m11_n$N.json | pick torchrun_default
tail -2 gpurun_out/m11_err.log | cut -c1-200
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tee gpurun_out/bench_m11_n1.json | pick n1_default
timeout 300 python -m pytest tests/test_gpu_exchange.py -q -m gpu -x -k "one_process" 2>&1 | tail -2
