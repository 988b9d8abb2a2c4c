#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python oracle/ref/make_astaroth_golden.py 2>&1 | tail -2
mkdir -p tests/golden && cp gpurun_out/astaroth_solve_ref.npz tests/golden/ 2>/dev/null
timeout 300 python -m pytest tests/test_oracle_astaroth.py -q -x 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_astaroth.py -q -m gpu -x 2>&1 | tail -15
timeout 300 oracle/_ref/ref_astaroth_solve 256 1e-8 - - 2>&1 | tail -3
timeout 300 python scripts/time_astaroth.py 256 f64 5
SB_AC_SHAPE=1 SKIP_CELL=1 SKIP_ITER=1 timeout 300 python scripts/time_astaroth.py 256 f64 5
timeout 300 python scripts/time_astaroth.py 256 f32 5
SB_AC_SHAPE=1 SKIP_CELL=1 SKIP_ITER=1 timeout 300 python scripts/time_astaroth.py 256 f32 5
( CUDA_VISIBLE_DEVICES=0 timeout 300 oracle/_ref/ref_astaroth 5 2>&1 | tail -2 )
( CUDA_VISIBLE_DEVICES=0 timeout 300 bin/astaroth 5 2>&1 | tail -2 )
( CUDA_VISIBLE_DEVICES=0 timeout 300 bin/astaroth_b200 5 2>&1 | tail -2 )
echo "=== jacobi step contention diagnostics (queued iterations) ==="
B="timeout 200 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e"
pick() { python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$1', 'ms/step', round(d['ms_per_step'],4), 'interior', round(d['roofline']['kernel_ms'],4))"; }
$B 2>/dev/null | pick default
SB_DEBUG_SKIP=ext $B 2>/dev/null | pick skip_exterior
SB_DEBUG_SKIP=xchg $B 2>/dev/null | pick skip_exchange
SB_DEBUG_SKIP=both $B 2>/dev/null | pick skip_both
SB_COPY_CTAS_PER_SM=1 $B 2>/dev/null | pick copy_ctas_1
SB_COPY_CTAS_PER_SM=2 $B 2>/dev/null | pick copy_ctas_2
SB_JACOBI_EXT_CTAS_PER_SM=1 $B 2>/dev/null | pick ext_ctas_1
SB_JACOBI_EXT_CTAS_PER_SM=4 $B 2>/dev/null | pick ext_ctas_4
SB_JACOBI_EXT_CTAS_PER_SM=2 SB_COPY_CTAS_PER_SM=1 $B 2>/dev/null | pick ext2_copy1
