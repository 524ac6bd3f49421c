#!/bin/bash
timeout 200 python scratch/perf_c3.py 2000 2>&1 | grep -E "per-class|eval 2"
timeout 200 python scratch/perf_large.py 512 16 1 2>&1 | tail -4
timeout 900 python -m pytest tests/test_eval_parity_gpu.py tests/test_fit_gpu.py tests/test_large_n_gpu.py tests/test_c3_parity_gpu.py -x -q 2>&1 | tail -3
