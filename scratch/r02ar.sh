#!/bin/bash
for B in 100 250 1000 2000; do for t in 0 1; do echo "== B=$B SOLO WV_TRTRI_ALL=$t"; SOLO=1 WV_TRTRI_ALL=$t timeout 200 python scratch/perf_c3.py $B 2>&1 | grep -E "eval 2|per-class"; done; done
for c in 444 740; do echo "== B=2000 SOLO trtri ctas $c"; SOLO=1 WV_TRTRI_CTAS=$c timeout 200 python scratch/perf_c3.py 2000 2>&1 | grep -E "per-class"; done
