#!/bin/bash
set -x
pick() { grep "^{" | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$1', 'n', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'step_frac', round(d['roofline']['step_frac_of_roofline'],3), 'launches', d['gpu_launches'])"; }
B="timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e"
SB_JACOBI_PUSH_MB=3 $B --schedule fused 2>/dev/null | pick fused_mb3
SB_JACOBI_PUSH_MB=3 SB_DEBUG_NOPUSH=1 $B --schedule fused 2>/dev/null | pick fused_mb3_nopush
SB_JACOBI_MB=3 python scripts/time_jacobi.py 512 f64 10 2>&1 | grep -E "interior|whole"
$B --schedule fused 2>/dev/null | pick fused_mb4
