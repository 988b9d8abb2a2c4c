#!/bin/bash
set -x
timeout 300 python -m pytest tests/test_gpu_copy.py -q -m gpu -k "tma" 2>&1 | tail -12
timeout 300 compute-sanitizer --tool memcheck --print-limit 2 --show-backtrace no python -m pytest tests/test_gpu_copy.py -q -m gpu -x -k "tma" 2>&1 | grep -A8 "=========" | head -40
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 120 python scripts/time_pack.py 512 1 float64 1
SB_TMA=0 timeout 120 python scripts/time_pack.py 512 1 float64 1
timeout 120 python scripts/time_pack.py 508 2 float64 1
SB_TMA=0 timeout 120 python scripts/time_pack.py 508 2 float64 1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | cut -c1-400
SB_TMA=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | cut -c1-400
