#!/bin/bash
for B in 100 250 500; do for t in 0 1; do echo "== B=$B WV_CHOL_ALL=$t"; WV_CHOL_ALL=$t timeout 200 python scratch/perf_c3.py $B 2>&1 | grep -E "per-class|eval 2"; done; done
for B in 100 250; do echo "== B=$B WV_TRTRI_ROWS=1"; WV_TRTRI_ROWS=1 timeout 200 python scratch/perf_c3.py $B 2>&1 | grep -E "per-class|eval 2"; done
