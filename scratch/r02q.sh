#!/bin/bash
timeout 600 python -m pytest tests/test_large_n_gpu.py -x -q 2>&1 | tail -4
for f in 1 0; do echo "== WV_PANEL_FUSED=$f"; WV_PANEL_FUSED=$f timeout 300 python scratch/perf_large.py 512 16 1 2>&1 | grep -E "eval 2|per-class|cholesky"; done
for pt in 3 5 6; do echo "== fused PT=$pt"; WV_PANEL_TILES=$pt timeout 300 python scratch/perf_large.py 512 16 1 2>&1 | grep -E "cholesky"; done
