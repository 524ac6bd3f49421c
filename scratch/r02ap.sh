#!/bin/bash
for B in 12 25 50 80; do for mn in 888 0; do echo "== B=$B SOLO WV_CHOL_ALL_MIN=$mn"; SOLO=1 WV_CHOL_ALL_MIN=$mn timeout 200 python scratch/perf_c3.py $B 2>&1 | grep -E "eval 2|per-class"; done; done
