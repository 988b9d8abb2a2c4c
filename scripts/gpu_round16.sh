#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_jacobi.py tests/test_gpu_exchange.py tests/test_gpu_copy.py -q -m gpu -x 2>&1 | tail -4
pick() { grep "^{" | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$1', 'n', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'step_frac', round(d['roofline']['step_frac_of_roofline'],3), 'launches', d['gpu_launches'])"; }
B="timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e"
$B --schedule fused 2>/dev/null | pick fused_aligned
SB_DEBUG_NOPUSH=1 $B --schedule fused 2>/dev/null | pick fused_aligned_nopush
SB_ALLOC_ALIGN=0 $B --schedule fused 2>/dev/null | pick fused_refalloc
$B --schedule queued 2>/dev/null | pick queued_aligned
SB_ALLOC_ALIGN=0 $B --schedule queued 2>/dev/null | pick queued_refalloc
$B --schedule host-sync 2>/dev/null | pick hostsync_aligned
python scripts/time_jacobi.py 512 f64 10 2>&1 | grep -E "interior|whole"
echo "=== astaroth after weight folding / early prev loads ==="
SKIP_CELL=1 SKIP_ITER=1 timeout 300 python scripts/time_astaroth.py 256 f64 5
SKIP_CELL=1 SKIP_ITER=1 timeout 300 python scripts/time_astaroth.py 256 f32 5
timeout 600 python -m pytest tests/test_gpu_astaroth.py -q -m gpu -x 2>&1 | tail -3
