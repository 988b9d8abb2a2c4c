#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_copy.py -q -m gpu -x -k "tma" 2>&1 | tail -6
timeout 300 compute-sanitizer --tool memcheck --print-limit 1 --show-backtrace no python -m pytest tests/test_gpu_copy.py -q -m gpu -x -k "tma and 64" 2>&1 | grep -A12 "=========" | head -40
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
