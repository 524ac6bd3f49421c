#!/bin/bash
for B in 30 50 80 250; do echo "== B=$B SOLO"; SOLO=1 timeout 200 python scratch/perf_c3.py $B 2>&1 | grep -E "eval 2"; done
timeout 900 python -m pytest tests/test_eval_parity_gpu.py tests/test_fit_gpu.py -x -q 2>&1 | tail -2
