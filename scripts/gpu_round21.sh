#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_jacobi.py -q -m gpu -x 2>&1 | tail -3
pick() { grep "^{" | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$1', 'n', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'step_frac', round(d['roofline']['step_frac_of_roofline'],3), 'launches', d['gpu_launches'])"; }
B="timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e"
$B --schedule fused 2>/dev/null | pick fused_wrap
$B --schedule fused --dtype f32 2>/dev/null | pick fused_wrap_f32
$B --schedule queued --dtype f32 2>/dev/null | pick queued_f32
