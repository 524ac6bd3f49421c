#!/bin/bash
for pt in 2 3 4 5 6; do for c in 148 222; do echo "== PT=$pt CTAS=$c"; WV_PANEL_TILES=$pt WV_PANEL_CTAS=$c timeout 300 python scratch/perf_large.py 512 16 1 2>&1 | grep -E "cholesky"; done; done
