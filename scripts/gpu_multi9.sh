#!/bin/bash
set -x
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_jacobi.py tests/test_gpu_exchange.py -q -m gpu -x -k "step_async or one_process" 2>&1 | tail -3
pick() { grep "^{" | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$1', 'n', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'per_gpu', '%.4g' % d['per_gpu'], 'launches', d['gpu_launches'], d['clocks']['sm_mhz'])"; }
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --schedule fused 2>/dev/null | pick n1_fused
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29841 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --schedule fused 2>gpurun_out/m9_err.log | pick torchrun_fused
tail -3 gpurun_out/m9_err.log | cut -c1-200
timeout 300 python bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --schedule fused 2>/dev/null | pick singleproc_fused
SB_FUSED_DENSE_X=0 timeout 300 python bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --schedule fused 2>/dev/null | pick singleproc_fused_direct
