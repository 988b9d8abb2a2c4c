#!/bin/bash
set -x
timeout 300 python -m pytest tests/test_gpu_jacobi.py tests/test_gpu_astaroth.py -q -m gpu -x 2>&1 | tail -2
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | cut -c1-330
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-330
