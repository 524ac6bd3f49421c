#!/bin/bash
for t in 0 1; do echo "== WV_KINV_TMA=$t"; WV_KINV_TMA=$t python scratch/perf_c3.py 2000 2>&1 | grep -E "per-class|eval 2"; done
WV_KINV_TMA=1 python -m pytest tests/test_eval_parity_gpu.py tests/test_fit_gpu.py -x -q 2>&1 | tail -3
