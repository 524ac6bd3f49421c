#!/bin/bash
for t in 0 1; do echo "== WV_CHOL_ALL=$t"; WV_CHOL_ALL=$t timeout 200 python scratch/perf_c3.py 2000 2>&1 | grep -E "per-class|eval 2"; done
for lag in 320 1280; do echo "== lag $lag"; WV_CHOL_LAG=$lag timeout 200 python scratch/perf_c3.py 2000 2>&1 | grep -E "per-class|eval 2"; done
timeout 600 python -m pytest tests/test_eval_parity_gpu.py tests/test_fit_gpu.py tests/test_c3_parity_gpu.py -x -q 2>&1 | tail -3
