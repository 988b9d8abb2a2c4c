#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -12
python scripts/time_pack.py 510 1 float64 1
SB_TMA=0 python scripts/time_pack.py 510 1 float64 1
python scripts/time_pack.py 512 1 float64 1
SB_TMA=0 python scripts/time_pack.py 512 1 float64 1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null
SB_TMA=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null
compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_copy.py -q -m gpu -k "tma" 2>&1 | tail -8
