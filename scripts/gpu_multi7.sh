#!/bin/bash
set -x
pick() { grep "^{" | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$1', 'n', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'per_gpu', '%.4g' % d['per_gpu'], 'launches', d['gpu_launches'], d['clocks']['sm_mhz'])"; }
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$T --master-port 29811 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --schedule fused 2>/dev/null | pick torchrun_fused
SB_DEBUG_NOPUSH=1 $T --master-port 29812 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --schedule fused 2>/dev/null | pick torchrun_fused_nopush
SB_DEBUG_NOXPUSH=1 $T --master-port 29813 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --schedule fused 2>/dev/null | pick torchrun_fused_noxpush
timeout 300 python bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --schedule fused 2>/dev/null | pick singleproc_fused
SB_DEBUG_NOPUSH=1 timeout 300 python bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --schedule fused 2>/dev/null | pick singleproc_fused_nopush
