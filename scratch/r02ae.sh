#!/bin/bash
for t in 0 1; do echo "== WV_PREFETCH_C=$t"; WV_PREFETCH_C=$t timeout 200 python scratch/perf_c3.py 2000 2>&1 | grep -E "per-class|eval 2"; done
WV_PREFETCH_C=1 timeout 200 python scratch/perf_large.py 2>&1 | tail -4
WV_PREFETCH_C=0 timeout 200 python scratch/perf_large.py 2>&1 | tail -4
timeout 600 python -m pytest tests/test_eval_parity_gpu.py tests/test_fit_gpu.py tests/test_large_n_gpu.py -x -q 2>&1 | tail -3
