#!/bin/bash
timeout 600 python -m pytest tests/test_large_n_gpu.py -x -q 2>&1 | tail -4
for c in 74 148 296 444; do echo "== WV_PANEL_CTAS=$c"; WV_PANEL_CTAS=$c timeout 300 python scratch/perf_large.py 512 16 1 2>&1 | grep -E "per-class|cholesky"; done
