#!/bin/bash
for g in 1 2 3 4; do echo "== WV_TRTRI_GROUP=$g"; WV_TRTRI_GROUP=$g timeout 200 python scratch/perf_c3.py 2000 2>&1 | grep -E "per-class|eval 2"; done
