#!/bin/bash
set -x
N=${1:-4}
SB_FUSED_IPC=1 timeout 300 python -m pytest tests/test_gpu_exchange.py -q -m gpu -x -k "one_process and ipc" 2>&1 | tail -3
pick() { grep "^{" | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$1', 'n', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'per_gpu', '%.4g' % d['per_gpu'], 'launches', d['gpu_launches'], d['config']['schedule'])"; }
SB_FUSED_IPC=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29891 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/m13_err.log | tee gpurun_out/bench_m13_n${N}_fused_ipc.json | pick torchrun_fused_ipc
tail -2 gpurun_out/m13_err.log | cut -c1-250
