#!/bin/bash
set -x
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
timeout 120 python scripts/time_pack.py 508 2 float64 1
SB_TMA=0 timeout 120 python scripts/time_pack.py 508 2 float64 1
timeout 120 python scripts/time_pack.py 504 4 float32 3
SB_TMA=0 timeout 120 python scripts/time_pack.py 504 4 float32 3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null
