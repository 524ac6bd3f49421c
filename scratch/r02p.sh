#!/bin/bash
for pt in 2 3 4 5 6 8; do
  echo "== WV_PANEL_TILES=$pt"
  WV_PANEL_TILES=$pt python scratch/perf_large.py 512 16 1 2>&1 | grep -E "eval 2|per-class|cholesky"
done
